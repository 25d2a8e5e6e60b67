"""Instructions executed and stall samples per CUDA source line (first kernel of the report).
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py [topN]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1  # 1-based index of the kernel in the report
rows = list(csv.reader(sys.stdin))
cur, hdr, agg, kernels, last_fn = None, None, [], 0, None
for r in rows:
    if not r:
        continue
    if r[0] == 'Function Name':
        if len(r) > 1 and r[1] != last_fn:
            kernels += 1
            last_fn = r[1]
    if r[0] == 'File Path':
        if kernels > which:
            break
        cur = r[1].split('/')[-1]
        continue
    if kernels != which:
        continue
    if r[0] == 'Line No':
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr and r[0].isdigit() and len(r) > 8:
        try:
            agg.append((int(r[ix['Instructions Executed']] or 0), int(r[ix['# Samples']] or 0), cur, r[0], r[1].strip()[:100]))
        except Exception:
            pass
tot = sum(a[0] for a in agg)
smp = sum(a[1] for a in agg)
print('total warp-insts', tot, 'samples', smp)
for a in sorted(agg, reverse=True)[:top]:
    print('%9d %5.1f%% smp=%5d (%4.1f%%) %s:%s  %s' % (a[0], 100 * a[0] / max(tot, 1), a[1], 100 * a[1] / max(smp, 1), a[2], a[3], a[4]))
