"""Pretty-print a bench.py JSON line (headline + `also` grid): python profiles/show_bench.py FILE"""
import json, sys
for f in sys.argv[1:]:
    txt = [l for l in open(f).read().splitlines() if l.startswith('{')]
    d = json.loads(txt[-1])
    print(f, {k: d.get(k) for k in ['value', 'ms_per_step', 'gpu_launches_per_step', 'path', 'n_gpus']}, 'roof', d['roofline']['bound'], round(d['roofline']['frac'], 3),
          'scan_ms', round(d['roofline']['avg_launch_ms'], 4), 'share', round(d['roofline']['scan_share_of_step'], 3), 'e2e', round(d['e2e']['value']),
          'parity', d.get('parity_check', {}).get('ok'), d.get('parity_check', {}).get('errors'), d.get('filter_retries'))
    for k, v in d.get('also', {}).items():
        if 'error' in v:
            print('  ', k, v); continue
        fr = v.get('filter_retries', {})
        print(f"   {k:20s} {v['value']:10.1f} q/s {v['ms_per_step']:8.3f} ms path={v['path']:7s} scan={v['scan_kernel_ms']:.3f}ms x{v['scan_launches_per_step']:.0f} "
              f"share={v['scan_share_of_step']:.3f} {v['bound']:6s} frac={v['roofline_frac']:.3f} l/step={v['gpu_launches_per_step']:.1f} retry={fr.get('first_stage_retry_rate')} "
              f"list={fr.get('longest_survivor_list_last_search')} bits={fr.get('overflow_bits_so_far')}", ('e2e=%.1f' % v['e2e']['value']) if 'e2e' in v else '')
    if 'cpu_baseline' in d:
        print('   cpu all-core', round(d['cpu_baseline']['value'], 3), 'cores', d['cpu_baseline']['cores'], '| single', round(d['cpu_baseline']['single_thread']['value'], 3))
