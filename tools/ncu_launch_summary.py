"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`): one training step (from one random_token_mask
launch to the next), launches and serialised time per kernel family.  With dram__bytes_* metrics in the file: bytes per family."""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
H = rows[0]; body = rows[1:]
kn, mn, mv = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value")
per = collections.OrderedDict()
for r in body:
    per.setdefault(r[0], {"name": r[kn]})[r[mn]] = float(r[mv].replace(",", ""))
L = list(per.values())
idx = [i for i, x in enumerate(L) if "token_mask" in x["name"]]
if len(idx) >= 2: L = L[idx[-2]:idx[-1]]
fam = lambda n: re.sub(r"<.*", "", n.split("(")[0]).replace("void ", "").replace("mae::", "")
tot = sum(x.get("gpu__time_duration.sum", 0) for x in L) / 1e3
agg, cnt, rd, wr = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for x in L:
    k = fam(x["name"]); agg[k] += x.get("gpu__time_duration.sum", 0) / 1e3; cnt[k] += 1
    rd[k] += x.get("dram__bytes_read.sum", 0); wr[k] += x.get("dram__bytes_write.sum", 0)
print(f"launches: {len(L)}   serialised (cold-cache) time: {tot:.1f} us")
print(f"{'kernel':34s} {'launches':>8s} {'us':>9s} {'share':>7s}" + ("   DRAM read / written (MB)" if sum(rd.values()) else ""))
for k, v in agg.most_common():
    extra = f"   {rd[k] / 1e6:9.1f} / {wr[k] / 1e6:9.1f}" if sum(rd.values()) else ""
    print(f"{k:34s} {cnt[k]:8d} {v:9.1f} {100 * v / tot:6.1f}%{extra}")
