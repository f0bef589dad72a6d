"""Timing experiments on the dense-layer kernel (GSD_GEMM_MODE bitmask, results are numerically wrong by design):
python tools/gemm_tc_modes.py  — run once per mode in a fresh process: for m in 0 1 2 4 8 ...; do GSD_GEMM_MODE=$m python tools/gemm_tc_modes.py; done"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gs_dynamics_b200 import gnn
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = []
for (M, N, K) in [(20021, 512, 512), (2001, 512, 512)]:
    x = torch.relu(torch.randn(M, K, device=dev)); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    ws = gnn._tc_split(w)
    ts = []
    for i in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = gnn._tc_linear(x, ws, b, relu=True); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    out.append("M=%d %.1f us" % (M, ts[len(ts) // 2]))
print("mode", os.environ.get("GSD_GEMM_MODE", "0"), "pair", os.environ.get("GSD_GEMM_PAIR", "-"), "BN", os.environ.get("GSD_GEMM_BN", "-"), " | ".join(out))
