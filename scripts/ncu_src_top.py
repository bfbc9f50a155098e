"""top stall lines of an ncu report (source page): python scripts/ncu_src_top.py rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if 'Source' in r)
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
si = hdr.index('Warp Stall Sampling (All Samples)'); ins = hdr.index('Source'); ie = hdr.index('Instructions Executed')
tot = sum(int(r[si]) for r in data); tote = sum(int(r[ie]) for r in data)
print('samples', tot, 'instr', tote, 'sass lines', len(data))
# cumulative view in program order, grouped in windows of ~N lines: print every line with >=0.5%
for idx, r in enumerate(data):
  pct = 100 * int(r[si]) / tot
  if pct >= float(sys.argv[3]) if len(sys.argv) > 3 else pct >= 0.7:
    print('%5d %5.1f%% exec %9s | %s' % (idx, pct, r[ie], r[ins].strip()[:110]))
