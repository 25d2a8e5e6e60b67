"""Summarise an ncu source page (SASS view): totals per stall reason and the hottest instructions.
usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_stalls.py [kernel-substring]"""
import csv, sys
want = sys.argv[1] if len(sys.argv) > 1 else None
rows = list(csv.reader(sys.stdin))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j]); j += 1
        i = j
        if want and want not in name: continue
        print("==", name[:120])
        idx = {h: k for k, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = {h: 0 for h in stall_cols}; samples = 0; insts = 0
        ops = {}
        for r in body:
            if len(r) < len(hdr): continue
            s = int(r[idx["# Samples"]] or 0); samples += s
            ie = int(r[idx["Instructions Executed"]] or 0); insts += ie
            for h in stall_cols: tot[h] += int(r[idx[h]] or 0)
            op = r[idx["Source"]].split()[0] if r[idx["Source"]].split() else "?"
            if op.startswith("@"): op = r[idx["Source"]].split()[1]
            op = op.split(".")[0]
            o = ops.setdefault(op, [0, 0]); o[0] += ie; o[1] += s
        print("samples", samples, "warp-insts", insts)
        print("stalls:", sorted(((v, k) for k, v in tot.items() if v), reverse=True)[:8])
        print("by opcode (samples, insts):", sorted(((v[1], v[0], k) for k, v in ops.items()), reverse=True)[:14])
        hot = sorted(body, key=lambda r: -int(r[idx["# Samples"]] or 0))[:12]
        for r in hot: print("   ", r[idx["# Samples"]], r[idx["Source"]].strip()[:90])
    else:
        i += 1
