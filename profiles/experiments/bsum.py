"""Summarise bench.py output lines (stdin): one short line per JSON record; other lines are echoed."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line:
        continue
    if line.startswith("{"):
        try:
            d = json.loads(line)
        except Exception:
            print(line[:200])
            continue
        r = d.get("roofline", {})
        c = d.get("clocks", {})
        print(f"{d.get('config', {}).get('workload', '?')[:40]:40s} value={d.get('value', 0):.1f} ms/step={d.get('ms_per_step', 0):.3f} "
              f"bound={r.get('bound')} frac={r.get('frac', 0):.3f} e2e={d.get('e2e', {}).get('value', 0):.1f} "
              f"sm_mhz={c.get('sm_mhz')} W={c.get('power_w_max')} launches={d.get('gpu_launches')} path={d.get('path')}")
    else:
        print(line[:200])
