"""HBM bandwidth probe for read/write mixes (torch elementwise kernels, device-timed): what a streaming kernel with a
given write fraction can reach on this GPU.  Used to interpret the gemm_tn roofline (write-heavy epilogues)."""
import torch
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
N = 512 * 1024 * 1024 // 4  # 512 MB fp32
a = torch.randn(N, device="cuda"); b = torch.randn(N, device="cuda"); c = torch.empty(N, device="cuda")
h = torch.empty(N // 2, device="cuda", dtype=torch.bfloat16)
GB = N * 4 / 1e9
t = timeit(lambda: c.zero_());               print(f"write only        : {GB / t:7.0f} GB/s")
t = timeit(lambda: c.copy_(a));              print(f"copy (1r:1w)      : {2 * GB / t:7.0f} GB/s")
t = timeit(lambda: torch.add(a, b, out=c));  print(f"add  (2r:1w)      : {3 * GB / t:7.0f} GB/s")
t = timeit(lambda: a.sum());                 print(f"read only (sum)   : {GB / t:7.0f} GB/s")
s = a[: N // 4]
t = timeit(lambda: torch.cat([s, s, s, s], out=c));  print(f"cat (1r:4w, L2 src): {(GB + GB / 4) / t:7.0f} GB/s")
t = timeit(lambda: h.copy_(a[: N // 2]));    print(f"cast f32->bf16 (2r:1w bytes): {(GB / 2 + GB / 4) / t:7.0f} GB/s")
