"""Microbenchmark: 3 separate TF32 GEMMs vs one K-concatenated GEMM for the error-compensated fp32 dense layers."""
import sys; sys.path.insert(0, '/root/repo')
import torch
torch.backends.cuda.matmul.allow_tf32 = True
dev = 'cuda'
def split(x):
    hi = ((x.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    return hi, x - hi
def timeit(f, n=30):
    for _ in range(5): f()
    big = torch.empty(64 << 20, device=dev)
    ts = []
    for _ in range(n):
        big.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
for rows, out in ((20000, 512), (2001, 1024), (2001, 512)):
    K = 512
    x = torch.randn(rows, K, device=dev); w = torch.randn(out, K, device=dev) / 22; b = torch.randn(out, device=dev)
    xh, xl = split(x); wh, wl = split(w)
    xc = torch.cat([xl, xh, xh], 1).contiguous(); wc = torch.cat([wh, wl, wh], 1).contiguous()
    whT, wlT, wcT = wh.t(), wl.t(), wc.t()
    ref = (x.double() @ w.double().t() + b.double())
    def f3():
        y = torch.addmm(b, xl, whT); y = torch.addmm(y, xh, wlT); return torch.addmm(y, xh, whT)
    def f1(): return torch.addmm(b, xc, wcT)
    torch.backends.cuda.matmul.allow_tf32 = False
    def fi(): return torch.addmm(b, x, w.t())
    t_i = timeit(fi); e_i = float((fi().double() - ref).abs().max())
    torch.backends.cuda.matmul.allow_tf32 = True
    t3 = timeit(f3); t1 = timeit(f1)
    e3 = float((f3().double() - ref).abs().max()); e1 = float((f1().double() - ref).abs().max())
    def ft(): return torch.addmm(b, x, w.t())
    tt = timeit(ft)
    print(f"rows {rows} out {out}: ieee {t_i:.1f} us err {e_i:.2e} | 3 gemms {t3:.1f} us err {e3:.2e} | K-concat {t1:.1f} us err {e1:.2e} | 1xtf32 {tt:.1f} us")
