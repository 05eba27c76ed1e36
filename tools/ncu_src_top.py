"""Top stall sites of an ncu source-page CSV (ncu -i x.ncu-rep --page source --csv): per kernel, the SASS instructions with
the most warp-stall samples, their dominant stall reasons and execution counts."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr): body.append(rows[j])
            j += 1
        col = {h: k for k, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[col["# Samples"]] or 0) for r in body)
        tot_inst = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
        print(f"== {name[:90]}  samples={tot} warp-insts={tot_inst}")
        agg = {h: sum(int(r[col[h]] or 0) for r in body) for h in stall_cols}
        print("   totals:", {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
        body_s = sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]
        for r in body_s:
            st = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
            print(f"   {r[col['# Samples']]:>6} {r[col['Instructions Executed']]:>8}  {r[col['Source']][:70]:<70} {st}")
        i = j
    else:
        i += 1
