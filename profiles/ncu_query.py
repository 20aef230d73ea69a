"""Pull selected metrics out of an .ncu-rep (ncu -i ... --page raw --csv): python profiles/ncu_query.py REPORT PATTERN..."""
import csv, subprocess, sys
f = sys.argv[1]; pats = sys.argv[2:]
out = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
for r in rows[2:]:
    print('--- ', r[h.index('Kernel Name')][:60])
    for i, name in enumerate(h):
        if any(p in name for p in pats) and not any(x in name for x in ['.min', '.max', 'peak_sustained', 'per_second', 'sparsity', '.sum.p']):
            try: v = float(r[i].replace(',', ''))
            except ValueError: continue
            if v != 0: print(f'   {name:90s} {r[i]}')
