"""Times gsd_linear_tf32x3 at the model's layer shapes against cuBLAS TF32 / fp32 (python tools/gemm_tc_prof.py [reps])."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gs_dynamics_b200 import gnn
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for (M, N, K) in [(20021, 512, 512), (2001, 512, 512), (2001, 1024, 512), (101000, 512, 512)]:
    x = torch.relu(torch.randn(M, K, device=dev)); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    ws = gnn._tc_split(w)
    def run_tc(): return gnn._tc_linear(x, ws, b, relu=True)
    def run_f32(): return torch.relu(torch.addmm(b, x, w.t()))
    for name, fn in (("tc", run_tc), ("cublas fp32", run_f32)):
        for warm in (False, True):
            ts = []
            for i in range(reps + 2):
                if not warm: flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); y = fn(); e1.record(); torch.cuda.synchronize()
                if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            print("M=%6d N=%4d K=%d %-12s %s  median %.1f us  -> %.1f TFLOP/s (real flops)" % (M, N, K, name, "warm" if warm else "cold", ts[len(ts) // 2], 2.0 * M * N * K / ts[len(ts) // 2] / 1e6))
