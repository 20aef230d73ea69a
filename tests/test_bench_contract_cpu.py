"""bench.py's output contract, checked without a GPU: helpers import, every workload BASELINE.json names is defined, and
the committed bench lines (profiles/bench_r1.jsonl, produced by bench.py on a B200) carry every key the driver reads."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "roofline", "e2e", "clocks", "gpu_launches"]


def test_bench_helpers_and_workloads():
    import bench
    assert set(["headline", "c1", "c2", "c3", "c4", "c5", "q1"]) <= set(bench.WORKLOADS)
    n, d, dt, Q, k, metric, _ = bench.WORKLOADS["headline"]
    assert (n, d, dt, Q, k, metric) == (10_000_000, 1024, "f32", 64, 10, "cosine")   # BASELINE.json metric shape
    assert bench.WORKLOADS["c3"][:6] == (10_000_000, 1024, "f32", 1024, 100, "dot")
    assert bench.WORKLOADS["c4"][:6] == (10_000_000, 768, "f16", 1, 10, "euclidean")
    assert bench.WORKLOADS["c5"][:6] == (100_000_000, 1024, "f32", 1024, 10, "cosine")
    peak, src = bench.load_peaks()
    assert 3000 < peak < 9000 and ("measured" in src or "fallback" in src)
    tf32, _ = bench.load_tensor_peak()
    bf16, _ = bench.load_tensor_peak(bf16=True)
    assert abs(bf16 - 2 * tf32) < 1e-6
    assert bench.load_traffic("headline", "shadow") > 2.0e10 and bench.load_traffic("headline", "filter") > 4.0e10


def test_committed_bench_lines_follow_the_contract():
    lines = [json.loads(l) for l in open(os.path.join(ROOT, "profiles", "bench_r1.jsonl"))]
    ours = [d for d in lines if d.get("impl") != "reference"]
    ref = [d for d in lines if d.get("impl") == "reference"]
    assert ours and ref
    for d in ours:
        for key in REQUIRED:
            assert key in d, (d.get("run"), key)
        assert d["metric"] == "kNN queries/sec" and d["unit"] == "queries/s" and d["higher_is_better"] is True
        assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["scaling"] == "strong"
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] in ("GB/s", "TFLOP/s")
        e = d["e2e"]
        assert e["unit"] == "queries/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] > 0
        assert d["gpu_launches"] > 0 and "workload" in d["config"] and "model" not in d["config"]
        assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    default = [d for d in ours if d.get("run", "").startswith("default")][0]
    cb = default["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "queries/s" and "sample" in cb
    for d in ref:
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["cpu_baseline"]["value"] == d["value"]
