"""Times the photometric kernels (stats + finish, grad) alone at the bench shape; L2 flushed before each launch.
Usage: [GSD_PH_NCOL=1|2] [GSD_PH_SEG=rows] python tools/ph_bench.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gs_dynamics_b200 import tracking as TR, _lib


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    x = torch.rand(6, 480, 640, device=dev)
    y = torch.rand(6, 480, 640, device=dev)
    m = torch.zeros(3, device=dev)
    c = torch.zeros(3, device=dev)
    ws = TR._ph_workspace(x)
    st = TR.target_stats(y)
    d = TR._ph_desc(x, y, 2, 0.8, 0.2, (50.0, 200.0), ws, affine=(m, c), y_stats=st)
    d0 = TR._ph_desc(x, y, 2, 0.8, 0.2, (50.0, 200.0), ws, affine=(m, c))
    out = torch.empty(7, device=dev)
    grad = torch.empty_like(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib = _lib.lib()

    def t(fn, n=20):
        ts = []
        for i in range(n + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    print("env NCOL=%s SEG=%s" % (os.environ.get("GSD_PH_NCOL"), os.environ.get("GSD_PH_SEG")))
    print("stats<1>+finish  %.1f us" % t(lambda: lib.gsd_photometric_forward(C.byref(d), out.data_ptr(), s)))
    print("stats<0>+finish  %.1f us" % t(lambda: lib.gsd_photometric_forward(C.byref(d0), out.data_ptr(), s)))
    print("grad             %.1f us" % t(lambda: lib.gsd_photometric_backward(C.byref(d), None, grad.data_ptr(), s)))
    print("target stats<2>  %.1f us" % t(lambda: lib.gsd_photometric_target_stats(6, 480, 640, y.data_ptr(), st[0].data_ptr(), st[1].data_ptr(), s)))
    print("loss", out.tolist())


if __name__ == "__main__":
    main()
