"""gsd_linear_tf32x3 (tcgen05 3xTF32 GEMM with fused bias / residual / ReLU epilogue) and the two small plain-fp32 layer kernels
against float64 matmul.  Tolerance: max |err| <= 2e-6 * (sum_k |a||w| bound) — fp32-SIMT level; a single-pass TF32 GEMM is ~1e-3."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _split(w):
    from gs_dynamics_b200 import _lib
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().gsd_tf32_split(w.numel(), w.data_ptr(), hi.data_ptr(), lo.data_ptr(), st), "gsd_tf32_split")
    return hi, lo


def _linear(x, w_hi, w_lo, bias=None, res1=None, res2=None, relu=False):
    from gs_dynamics_b200 import _lib
    M, K = x.shape
    N = w_hi.shape[0]
    out = torch.full((M, N), float("nan"), device=x.device)
    p = lambda t: t.data_ptr() if t is not None else None
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().gsd_linear_tf32x3(M, N, K, x.data_ptr(), x.stride(0), w_hi.data_ptr(), w_lo.data_ptr(), p(bias), p(res1), p(res2),
                                            int(relu), out.data_ptr(), N, st), "gsd_linear_tf32x3")
    return out


@pytest.mark.parametrize("M,N,K,opts", [
    (128, 64, 32, ""), (130, 64, 64, "b"), (2001, 512, 512, "br"), (2001, 1024, 512, ""), (20000, 512, 512, "bR"),
    (20000, 512, 512, "b12R"), (300, 128, 96, "b1"), (24021, 256, 512, "bR"), (77, 512, 512, "b2R")])
def test_linear_tf32x3_matches_float64(M, N, K, opts):
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).cuda()
    if "R" in opts:
        x = torch.relu(x)          # activations after a ReLU, as in the model
    w = (torch.randn(N, K, generator=g) / np.sqrt(K)).cuda()
    bias = torch.randn(N, generator=g).cuda() if "b" in opts else None
    r1 = torch.randn(M, N, generator=g).cuda() if "1" in opts else None
    r2 = torch.randn(M, N, generator=g).cuda() if "2" in opts else None
    relu = "r" in opts or "R" in opts
    w_hi, w_lo = _split(w)
    assert float((w_hi + w_lo - w).abs().max()) <= 2.0 ** -21 * float(w.abs().max())
    out = _linear(x, w_hi, w_lo, bias, r1, r2, relu)
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t()
    for t in (bias, r1, r2):
        if t is not None:
            ref = ref + t.double()
    if relu:
        ref = torch.relu(ref)
    bound = (x.abs().double() @ w.abs().double().t()).max().item()
    err = float((out.double() - ref).abs().max())
    print("M=%d N=%d K=%d %-5s max err %.3g (bound scale %.3g -> rel %.3g)" % (M, N, K, opts, err, bound, err / bound))
    assert torch.isfinite(out).all()
    assert err <= 2e-6 * bound


def test_small_layers_match_float64():
    from gs_dynamics_b200 import _lib
    g = torch.Generator().manual_seed(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for M, N, K in ((2001, 512, 14), (20000, 512, 14), (501, 512, 5), (37, 128, 32)):
        x, w, b = torch.randn(M, K, generator=g).cuda(), torch.randn(N, K, generator=g).cuda(), torch.randn(N, generator=g).cuda()
        out = torch.empty(M, N, device="cuda")
        _lib.check(_lib.lib().gsd_linear_small(M, N, K, x.data_ptr(), K, w.data_ptr(), b.data_ptr(), 1, out.data_ptr(), st), "small k")
        ref = torch.relu(x.double() @ w.double().t() + b.double())
        assert float((out.double() - ref).abs().max()) < 1e-5
    for M, N, K in ((2000, 3, 512), (101, 3, 128)):
        x, w, b = torch.randn(M, K, generator=g).cuda(), torch.randn(N, K, generator=g).cuda(), torch.randn(N, generator=g).cuda()
        out = torch.empty(M, N, device="cuda")
        _lib.check(_lib.lib().gsd_linear_small(M, N, K, x.data_ptr(), K, w.data_ptr(), b.data_ptr(), 0, out.data_ptr(), st), "small n")
        ref = x.double() @ w.double().t() + b.double()
        assert float((out.double() - ref).abs().max()) < 2e-5
