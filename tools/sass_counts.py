"""Per-kernel SASS evidence for libmae_b200.so (`cuobjdump -sass`): tcgen05 MMAs (UTCHMMA), TMA loads / stores (UTMALDG / UTMASTG),
TMEM loads (LDTM), legacy warp MMAs (HMMA), async copies (LDGSTS) and the ELECT / BRA.U.ANY waterfalls the compiler wraps
around a uniform-datapath instruction issued from a thread-dependent branch."""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ssrl-vit-mae-jepa_b200", "libmae_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "HMMA", "LDGSTS", "BRA.U.ANY", "FFMA2", "MUFU.EX2"]
cur, h = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "").replace("mae::", "")
        h[cur] = collections.Counter()
        continue
    if cur:
        for o in ops:
            if re.search(r"\b" + re.escape(o) + r"\b", line): h[cur][o] += 1
print(f"{'kernel':58s}" + "".join(f"{o:>10s}" for o in ops))
for k, c in sorted(h.items()):
    if any(c[o] for o in ops[:6]): print(f"{k[:58]:58s}" + "".join(f"{c[o]:10d}" for o in ops))
