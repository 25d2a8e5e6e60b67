"""Summarise an ncu source page (SASS view): totals per stall reason and the hottest instructions.
usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_stalls.py [kernel-substring] [--flow]
--flow adds the sample distribution in program order (bins of 32 SASS instructions, with the synchronisation / memory
opcodes each bin contains), which shows WHERE in the kernel's phases the warps sit."""
import csv, sys
flow = "--flow" in sys.argv
args = [a for a in sys.argv[1:] if a != "--flow"]
want = args[0] if args else None
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
        if flow:
            marks = ("BAR", "SYNCS", "UBLKCP", "LDGSTS", "LDG", "STG", "LDS", "STS", "MUFU", "BRA", "CALL", "RET", "EXIT", "WARPSYNC", "NANOSLEEP", "UTMASTG", "ATOM", "RED", "MEMBAR", "FENCE", "LDGDEPBAR", "DEPBAR", "ARRIVES")
            good = [r for r in body if len(r) >= len(hdr)]
            for b0 in range(0, len(good), 32):
                chunk = good[b0:b0 + 32]
                sm = sum(int(r[idx["# Samples"]] or 0) for r in chunk)
                ie = sum(int(r[idx["Instructions Executed"]] or 0) for r in chunk)
                seen = []
                for r in chunk:
                    w = r[idx["Source"]].split()
                    if not w: continue
                    op = (w[1] if w[0].startswith("@") and len(w) > 1 else w[0]).split(".")[0]
                    if op in marks and op not in seen: seen.append(op)
                st = {h: sum(int(r[idx[h]] or 0) for r in chunk) for h in stall_cols}
                top = sorted(((v, k.replace("stall_", "")) for k, v in st.items() if v), reverse=True)[:2]
                print("   [%4d] samples %5d insts %9d  %-40s %s" % (b0, sm, ie, ",".join(seen), top))
    else:
        i += 1
