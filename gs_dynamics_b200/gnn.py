"""Host side of the GNN particle-dynamics step (path B), backed by the sm_100a graph kernels of libgsd_b200.so.

Mirrors (same names, argument meaning, results):
  DynamicsPredictor(model_config, device).forward(state, attrs, Rr, Rs, p_instance, action=None, **kw)
                                             /root/reference/src/gnn/model.py:70-246  (state_dict layout identical)
  construct_edges_from_states[_batch]        /root/reference/src/data/dataset.py:88-216
  farthest_point_sampler                     dgl.geometry (call sites /root/reference/src/render/dynamics_module.py:46,65)
  fps_rad_idx_torch                          /root/reference/src/data/utils.py:50-65
Design: edges are index lists grouped by receiver (CSR), never one-hot matrices; the relation propagator
Linear([enc_e | h_r | h_s]) is split into W1 enc_e (once per step) + W2 h_r + W3 h_s (node-level GEMM per pstep), so the
per-pstep edge work is one fused gather + ReLU + segment-sum kernel.  Dense layers: inference runs every F-wide nn.Linear as ONE
hand-written tcgen05 kernel (csrc/gemm_tc.cu: error-compensated 3xTF32 with fp32-level accuracy, bias / residual / ReLU in the
epilogue, no library GEMM); training uses the same compensation through cuBLAS TF32 GEMMs (_Linear3x); matmul="ieee" keeps
plain fp32 SIMT GEMMs for bit-tight comparisons.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, _nvtx


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _cuda_f32(t, name):
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (no CPU path)" % name)
    return t.contiguous().float()


# ----------------------------------------------------------------------------------------------------
# farthest point sampling
# ----------------------------------------------------------------------------------------------------
def farthest_point_sampler(pos, npoints, start_idx=None):
    """dgl.geometry.farthest_point_sampler: pos [B,N,3] -> int64 [B,npoints]. CPU inputs (the reference passes
    `.cpu()` tensors) are staged to the GPU and the result is returned on the input's device."""
    src_dev = pos.device
    p = pos.detach()
    if not p.is_cuda:
        p = p.cuda()
    p = p.contiguous().float()
    B, N, _ = p.shape
    if start_idx is None:
        start = torch.randint(0, N, (B,), device=p.device, dtype=torch.int64)
    elif isinstance(start_idx, int):
        start = torch.full((B,), start_idx, device=p.device, dtype=torch.int64)
    else:
        start = torch.as_tensor(start_idx, device=p.device, dtype=torch.int64).reshape(B).contiguous()
    out = torch.empty((B, npoints), dtype=torch.int64, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(_lib.lib().gsd_fps(B, N, npoints, 0.0, p.data_ptr(), start.data_ptr(), out.data_ptr(), None, _stream()), "gsd_fps")
    return out.to(src_dev)


def fps_rad_idx_torch(pcd, radius, start_idx=None):
    """Radius-terminated FPS of data/utils.py:50-65: returns (pcd_fps, idx_lst). The reference draws the first index
    with np.random.randint; pass start_idx for reproducibility."""
    src_dev = pcd.device
    p = pcd.detach()
    if not p.is_cuda:
        p = p.cuda()
    p = p.contiguous().float()
    N = p.shape[0]
    if start_idx is None:
        start_idx = int(np.random.randint(N))
    start = torch.tensor([start_idx], device=p.device, dtype=torch.int64)
    out = torch.empty((1, N), dtype=torch.int64, device=p.device)
    count = torch.zeros(1, dtype=torch.int32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(_lib.lib().gsd_fps(1, N, N, float(radius), p.data_ptr(), start.data_ptr(), out.data_ptr(), count.data_ptr(),
                                      _stream()), "gsd_fps")
    n = int(count.item())  # the reference's own loop syncs once per pick
    idx = out[0, :n]
    return pcd[idx.to(src_dev)], idx.to(src_dev)


# ----------------------------------------------------------------------------------------------------
# edges
# ----------------------------------------------------------------------------------------------------
class EdgeIndex:
    """Edges grouped by receiver. receivers/senders: int32 [B,capacity] (-1 = unused slot); row_ptr: int32 [B,N+1]."""
    __slots__ = ("B", "N", "capacity", "n_tool", "row_ptr", "n_edges", "receivers", "senders", "_tr")

    def transposed(self):
        """(col_ptr int32 [B,N+1], order int32 [B,capacity]): the same edges grouped by SENDER (order = edge slots sorted by
        sender, stable; unused slots last) — what the backward kernels walk.  Built once per graph with device ops, no host sync."""
        tr = getattr(self, "_tr", None)
        if tr is None:
            N = self.N
            key = torch.where(self.senders >= 0, self.senders, torch.full_like(self.senders, N)).long()
            order = torch.argsort(key, dim=1, stable=True)
            counts = torch.zeros((self.B, N + 1), dtype=torch.int64, device=key.device)
            counts.scatter_add_(1, key, torch.ones_like(key))
            cp = torch.zeros((self.B, N + 1), dtype=torch.int32, device=key.device)
            cp[:, 1:] = torch.cumsum(counts[:, :N], 1).to(torch.int32)
            tr = (cp.contiguous(), order.to(torch.int32).contiguous())
            self._tr = tr
        return tr

    def dense(self):
        """(Rr, Rs) float32 [B, max_e, N] one-hot matrices exactly as the reference returns them (host sync)."""
        n = self.n_edges.cpu()
        E = int(n.max())
        Rr = torch.zeros((self.B, E, self.N), device=self.receivers.device)
        Rs = torch.zeros((self.B, E, self.N), device=self.receivers.device)
        for b in range(self.B):
            e = int(n[b])
            ar = torch.arange(e, device=Rr.device)
            Rr[b, ar, self.receivers[b, :e].long()] = 1
            Rs[b, ar, self.senders[b, :e].long()] = 1
        return Rr, Rs


def edge_capacity(N, n_tool, topk):
    n_obj = N - n_tool
    return n_obj * min(topk, max(n_obj, 1)) + 2 * n_obj * n_tool


def construct_edges_index(states, adj_thresh, mask, tool_mask, topk=10, connect_all=False, n_tool=None, capacity=None):
    """states [N,3] or [B,N,3]; mask/tool_mask bool; tools must be the last n_tool nodes (dataset.py:120-124).
    n_tool=None reads it from tool_mask (one host sync, as the reference's `.item()` does)."""
    st = _cuda_f32(states, "states")
    if st.dim() == 2:
        st, mask, tool_mask = st[None], mask[None], tool_mask[None]
    B, N, _ = st.shape
    dev = st.device
    if n_tool is None:
        nt = tool_mask.sum(dim=-1)
        n_tool = int(nt.max().item())
        assert n_tool == int(nt.min().item()), 'only support fixed number of tool particles'
    if capacity is None:
        capacity = edge_capacity(N, n_tool, topk)
    m8 = mask.to(torch.uint8).contiguous()
    t8 = tool_mask.to(torch.uint8).contiguous()
    g = _lib.GsdGnnEdges()
    g.B, g.N, g.n_tool, g.topk, g.connect_all, g.capacity = B, N, n_tool, int(topk), int(bool(connect_all)), int(capacity)
    g.states, g.mask, g.tool_mask = st.data_ptr(), m8.data_ptr(), t8.data_ptr()
    thr_t = None
    if isinstance(adj_thresh, torch.Tensor):
        thr_t = adj_thresh.to(dev).float().reshape(-1)
        thr_t = (thr_t.expand(B) if thr_t.numel() == 1 else thr_t).contiguous()
        g.adj_thresh = thr_t.data_ptr()
    else:
        g.adj_thresh = None
        g.adj_thresh_sq_scalar = float(adj_thresh) * float(adj_thresh)
    nbytes = C.c_size_t()
    lib = _lib.lib()
    _lib.check(lib.gsd_gnn_edges_workspace_bytes(B, N, C.byref(nbytes)), "gsd_gnn_edges_workspace_bytes")
    e = EdgeIndex()
    e.B, e.N, e.capacity, e.n_tool = B, N, int(capacity), n_tool
    with torch.cuda.device(dev):
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        e.row_ptr = torch.empty((B, N + 1), dtype=torch.int32, device=dev)
        e.n_edges = torch.empty(B, dtype=torch.int32, device=dev)
        e.receivers = torch.empty((B, capacity), dtype=torch.int32, device=dev)
        e.senders = torch.empty((B, capacity), dtype=torch.int32, device=dev)
        g.ws, g.row_ptr, g.n_edges = ws.data_ptr(), e.row_ptr.data_ptr(), e.n_edges.data_ptr()
        g.receivers, g.senders = e.receivers.data_ptr(), e.senders.data_ptr()
        _lib.check(lib.gsd_gnn_build_edges(C.byref(g), _stream()), "gsd_gnn_build_edges")
    return e


def construct_edges_from_states(states, adj_thresh, mask, tool_mask, topk=10, connect_all=False):
    """Reference contract: returns dense one-hot (Rr, Rs) of shape [n_rel, N]."""
    Rr, Rs = construct_edges_index(states, adj_thresh, mask, tool_mask, topk, connect_all).dense()
    return Rr[0], Rs[0]


def construct_edges_from_states_batch(states, adj_thresh, mask, tool_mask, topk=10, connect_all=False):
    """Reference contract: returns dense one-hot (Rr, Rs) of shape [B, max n_rel, N] (zero rows = padding)."""
    return construct_edges_index(states, adj_thresh, mask, tool_mask, topk, connect_all).dense()


def edge_index_from_dense(Rr, Rs, n_heavy=0):
    """One-hot (Rr, Rs) [B,n_rel,N] (zero rows = padding, pad_torch of dataset.py:229-238) -> EdgeIndex, no host sync."""
    B, E, N = Rr.shape
    valid = Rr.sum(-1) > 0.5
    recv = torch.where(valid, Rr.argmax(-1), torch.full_like(valid, N, dtype=torch.int64))
    send = Rs.argmax(-1)
    order = torch.argsort(recv, dim=1, stable=True)
    recv_s = torch.gather(recv, 1, order)
    send_s = torch.gather(send, 1, order)
    ok = recv_s < N
    e = EdgeIndex()
    e.B, e.N, e.capacity, e.n_tool = B, N, E, n_heavy
    e.receivers = torch.where(ok, recv_s, torch.full_like(recv_s, -1)).to(torch.int32).contiguous()
    e.senders = torch.where(ok, send_s, torch.full_like(send_s, -1)).to(torch.int32).contiguous()
    counts = torch.zeros((B, N + 1), dtype=torch.int64, device=Rr.device)
    counts.scatter_add_(1, recv_s, torch.ones_like(recv_s))
    rp = torch.zeros((B, N + 1), dtype=torch.int32, device=Rr.device)
    rp[:, 1:] = torch.cumsum(counts[:, :N], 1).to(torch.int32)
    e.row_ptr = rp.contiguous()
    e.n_edges = rp[:, N].contiguous()
    return e


# ----------------------------------------------------------------------------------------------------
# dense layers: fp32 accuracy on the tensor cores by error-compensated TF32 splitting (x = x_hi + x_lo, three TF32 GEMMs:
# hi*hi + hi*lo + lo*hi; the dropped lo*lo term is ~2^-22 relative).  The GEMMs themselves are cuBLAS (library code);
# "ieee" keeps plain fp32 SIMT GEMMs.
# Inference uses the hand-written tcgen05 kernel (gsd_linear_tf32x3, csrc/gemm_tc.cu) instead: same compensation, operands split in
# shared memory, fused epilogues.
# ----------------------------------------------------------------------------------------------------
def _tf32_pack(x, relu=False, add=None, want_full=False, weight=False):
    """Error-compensated TF32 operand of t = relu?(x + add) (gsd_tf32_pack): activations -> [rows, 3F] = [lo | hi | hi],
    weights -> [out, 3K] = [hi | lo | hi], so that one TF32 GEMM over K = 3F gives the fp32 product.  Returns (packed, t or None)."""
    x = x.contiguous()
    rows, Fd = x.shape
    out = torch.empty((rows, 3 * Fd), device=x.device, dtype=torch.float32)
    full = torch.empty_like(x) if want_full else None
    if add is not None:
        add = add.contiguous()
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().gsd_tf32_pack(rows, Fd, int(relu), int(weight), x.data_ptr(), add.data_ptr() if add is not None else None,
                                            full.data_ptr() if want_full else None, out.data_ptr(), _stream()), "gsd_tf32_pack")
    return out, full


def _tc_split(w):
    """(w_hi, w_lo): the two TF32 operands of a weight matrix for gsd_linear_tf32x3."""
    w = w.detach().contiguous()
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    with torch.cuda.device(w.device):
        _lib.check(_lib.lib().gsd_tf32_split(w.numel(), w.data_ptr(), hi.data_ptr(), lo.data_ptr(), _stream()), "gsd_tf32_split")
    return hi, lo


def _tc_linear(x, wsplit, bias=None, res1=None, res2=None, relu=False):
    """act(x W^T + bias + res1 + res2) on the tcgen05 tensor cores with fp32-level accuracy (gsd_linear_tf32x3: hand-written
    sm_100a kernel, error-compensated 3xTF32, the activation split happens in shared memory).  x: [M, K] with unit column stride."""
    w_hi, w_lo = wsplit
    M, K = x.shape
    N = w_hi.shape[0]
    if x.stride(1) != 1 or x.stride(0) % 4 != 0:
        x = x.contiguous()
    out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    p = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().gsd_linear_tf32x3(M, N, K, x.data_ptr(), x.stride(0), w_hi.data_ptr(), w_lo.data_ptr(), p(bias), p(res1), p(res2),
                                                int(relu), out.data_ptr(), N, _stream()), "gsd_linear_tf32x3")
    return out


def _small_linear(x, W, bias, relu=False):
    """The two layer shapes that are not tensor-core work (K <= 32 or N <= 8), plain fp32 (gsd_linear_small)."""
    M, K = x.shape
    N = W.shape[0]
    if x.stride(1) != 1:
        x = x.contiguous()
    out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().gsd_linear_small(M, N, K, x.data_ptr(), x.stride(0), W.data_ptr(), bias.data_ptr() if bias is not None else None,
                                               int(relu), out.data_ptr(), _stream()), "gsd_linear_small")
    return out


class _Linear3x(torch.autograd.Function):
    """y = x W^T + b with fp32 accuracy on the TF32 tensor cores, forward AND backward (GNN training):
    forward / grad-input are one GEMM over the K-concatenated error-compensated operands of gsd_tf32_pack; grad-weight
    contracts over the rows, so it is three TF32 GEMMs on the hi / lo halves (strided views of the packed operands, small
    products first).  cuBLAS does the GEMMs (library code); the split/pack is this repo's kernel."""

    @staticmethod
    def forward(ctx, x, W, b):
        xp, _ = _tf32_pack(x)                                        # [rows, 3K] = [lo | hi | hi]
        Wp, _ = _tf32_pack(W.detach(), weight=True)                  # [out, 3K]  = [hi | lo | hi]
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            y = torch.addmm(b, xp, Wp.t()) if b is not None else xp @ Wp.t()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        ctx.save_for_backward(xp, W)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        xp, W = ctx.saved_tensors
        K = W.shape[1]
        gy = gy.contiguous()
        gp, _ = _tf32_pack(gy)                                       # [rows, 3*out] = [lo | hi | hi]
        O = gy.shape[1]
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            gx = gW = None
            if ctx.needs_input_grad[0]:
                Wtp, _ = _tf32_pack(W.detach().t().contiguous(), weight=True)   # [K, 3*out]
                gx = gp @ Wtp.t()
            if ctx.needs_input_grad[1]:
                g_lo, g_hi = gp[:, :O], gp[:, O:2 * O]
                x_lo, x_hi = xp[:, :K], xp[:, K:2 * K]
                gW = g_lo.t() @ x_hi
                gW.addmm_(g_hi.t(), x_lo)
                gW.addmm_(g_hi.t(), x_hi)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        gb = gy.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gx, gW, gb


# hi / lo splits of the weights inside ONE training iteration (forward + backward of the unroll: the weights are constant until
# the optimiser step, and every unroll step / the backward would split them again).  Only valid inside split_scope(): CUDA-graph
# replays change the weights without touching their version counters, so nothing may outlive the iteration.  Entries keep their
# source tensor alive (its address cannot be recycled while cached).
_SPLIT_CACHE = None


class split_scope:
    def __enter__(self):
        global _SPLIT_CACHE
        self.prev, _SPLIT_CACHE = _SPLIT_CACHE, {}
        return self

    def __exit__(self, *exc):
        global _SPLIT_CACHE
        _SPLIT_CACHE = self.prev
        return False


def _train_split(W, transposed):
    key = (W.data_ptr(), tuple(W.shape), tuple(W.stride()), transposed)
    if _SPLIT_CACHE is not None and key in _SPLIT_CACHE:
        return _SPLIT_CACHE[key][1]
    ws = _tc_split(W.detach().t().contiguous() if transposed else W)
    if _SPLIT_CACHE is not None:
        _SPLIT_CACHE[key] = (W, ws)
    return ws


class _LinearTC(torch.autograd.Function):
    """y = x W^T + b for GNN training with the hand-written tcgen05 kernel (gsd_linear_tf32x3: error-compensated 3xTF32, operands
    split on chip) in the forward AND the grad-input product (gx = gy W = gsd_linear_tf32x3(gy, split(W^T))): no packed copy of
    the activations in front of either.  grad-weight contracts over the rows (gW = gy^T x), which the kernel's K-major operand
    layout does not cover: it stays three TF32 library GEMMs on the hi / lo halves of the packed operands, as in _Linear3x."""

    @staticmethod
    def forward(ctx, x, W, b):
        y = _tc_linear(x, _train_split(W, False), b)
        ctx.save_for_backward(x, W)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        K, O = W.shape[1], W.shape[0]
        gy = gy.contiguous()
        gx = gW = None
        if ctx.needs_input_grad[0]:
            gx = _tc_linear(gy, _train_split(W, True))
        if ctx.needs_input_grad[1]:
            xp, _ = _tf32_pack(x)                                        # [rows, 3K] = [lo | hi | hi]
            gp, _ = _tf32_pack(gy)                                       # [rows, 3O] = [lo | hi | hi]
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = True
            try:
                g_lo, g_hi = gp[:, :O], gp[:, O:2 * O]
                x_lo, x_hi = xp[:, :K], xp[:, K:2 * K]
                gW = g_lo.t() @ x_hi
                gW.addmm_(g_hi.t(), x_lo)
                gW.addmm_(g_hi.t(), x_hi)
            finally:
                torch.backends.cuda.matmul.allow_tf32 = prev
        gb = gy.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gx, gW, gb


# training GEMMs: "cublas" = library TF32 GEMMs over packed operands (_Linear3x), "tc" = the hand-written kernel for forward +
# grad-input (_LinearTC).  Measured on B200 (bench.py's gnn_train iteration, batch 64 x n_future 5, one CUDA graph): cublas
# 27.5 ms, tc 28.5 ms (eager: 35.9 / 34.6 ms) — the library's single K = 3 x 512 GEMM per product still beats the hand-written
# kernel's ~60 % of the TF32 peak by more than its pack kernel costs, so the library path stays the default; GSD_TRAIN_GEMM=tc
# (or gnn.TRAIN_GEMM = "tc") selects the hand-written one.  Both are covered by the gradient-parity tests.
TRAIN_GEMM = os.environ.get("GSD_TRAIN_GEMM", "cublas")


def _linear(x, W, b, fast):
    """F.linear, or its 3xTF32 tensor-core twin when `fast` and the shape suits the kernels."""
    if fast and x.is_cuda and x.dim() == 2 and x.shape[0] >= 256 and W.shape[1] % 4 == 0 and W.shape[1] >= 128 and W.shape[0] % 4 == 0 \
            and W.shape[0] >= 128:
        if TRAIN_GEMM == "tc" and W.shape[1] % 32 == 0 and W.shape[0] % 64 == 0 and W.shape[1] % 64 == 0:
            return _LinearTC.apply(x.contiguous(), W, b)
        return _Linear3x.apply(x, W, b)
    return F.linear(x, W, b)


# ----------------------------------------------------------------------------------------------------
# fused kernels as autograd functions
# ----------------------------------------------------------------------------------------------------
class _EdgeInputs(torch.autograd.Function):
    """rel_inputs of model.py:164-199 (gsd_gnn_edge_inputs); differentiable w.r.t. state (the position differences)."""

    @staticmethod
    def forward(ctx, state, attrs, p_instance, edges):
        B, n_his, N, _ = state.shape
        attr_dim = attrs.shape[2]
        n_p, n_inst = p_instance.shape[1], p_instance.shape[2]
        width = 2 * attr_dim + 1 + 3 * n_his
        state = state.contiguous()
        out = torch.empty((B, edges.capacity, width), dtype=torch.float32, device=state.device)
        with torch.cuda.device(state.device):
            _lib.check(_lib.lib().gsd_gnn_edge_inputs(B, N, edges.capacity, n_his, attr_dim, n_inst, n_p, state.data_ptr(),
                                                      attrs.data_ptr(), p_instance.data_ptr(), edges.receivers.data_ptr(),
                                                      edges.senders.data_ptr(), out.data_ptr(), _stream()), "gsd_gnn_edge_inputs")
        ctx.edges = edges
        ctx.dims = (B, n_his, N, width, 2 * attr_dim + 1)
        return out

    @staticmethod
    def backward(ctx, g):
        e = ctx.edges
        B, n_his, N, width, off = ctx.dims
        g = g.contiguous().float()
        col_ptr, order = e.transposed()
        gs = torch.empty((B, n_his, N, 3), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().gsd_gnn_edge_inputs_bwd(B, N, e.capacity, n_his, width, off, e.row_ptr.data_ptr(), col_ptr.data_ptr(),
                                                          order.data_ptr(), g.data_ptr(), gs.data_ptr(), _stream()),
                       "gsd_gnn_edge_inputs_bwd")
        return gs, None, None, None


def edge_inputs(state, attrs, p_instance, edges, attr_dim_used=True, group=True):
    """rel_inputs [B, capacity, 2*attr_dim + 1 + 3*n_his] (model.py:164-199)."""
    if state.shape[3] != 3:
        raise ValueError("state must be [B, n_his, N, 3]")
    return _EdgeInputs.apply(state, attrs.contiguous(), p_instance.contiguous(), edges)


_AGG_WS = {}


class _Aggregate(torch.autograd.Function):
    """agg[node] = sum_e ReLU(A[e] + P[node,:F] + P[send(e),F:]).  Backward (GNN training): gsd_gnn_aggregate_bwd recomputes
    the pre-activation sign, so the forward saves nothing but its inputs."""

    @staticmethod
    def forward(ctx, A, P, edges):
        B, N, cap = edges.B, edges.N, edges.capacity
        Fd = A.shape[-1]
        A = A.contiguous()
        P = P.contiguous()
        lib = _lib.lib()
        nbytes = C.c_size_t()
        _lib.check(lib.gsd_gnn_aggregate_workspace_bytes(B, edges.n_tool, Fd, C.byref(nbytes)), "gsd_gnn_aggregate_workspace_bytes")
        with torch.cuda.device(A.device):
            # zero-filled once, self-cleaning afterwards (include/gsd.h); the calls sharing it are issued in stream order
            key = (A.device, B, edges.n_tool, Fd)
            ws = _AGG_WS.get(key)
            if ws is None:
                ws = _AGG_WS[key] = torch.zeros(max(int(nbytes.value), 16), dtype=torch.uint8, device=A.device)
            agg = torch.empty((B * N, Fd), dtype=torch.float32, device=A.device)
            _lib.check(lib.gsd_gnn_aggregate(B, N, cap, Fd, edges.n_tool, edges.row_ptr.data_ptr(), edges.senders.data_ptr(),
                                             A.data_ptr(), P.data_ptr(), ws.data_ptr(), agg.data_ptr(), _stream()), "gsd_gnn_aggregate")
        ctx.save_for_backward(A, P)
        ctx.edges = edges
        return agg

    @staticmethod
    def backward(ctx, g):
        A, P = ctx.saved_tensors
        e = ctx.edges
        B, N, cap = e.B, e.N, e.capacity
        Fd = A.shape[-1]
        g = g.contiguous().float()
        col_ptr, order = e.transposed()
        lib = _lib.lib()
        nbytes = C.c_size_t()
        _lib.check(lib.gsd_gnn_aggregate_bwd_workspace_bytes(B, e.n_tool, Fd, C.byref(nbytes)), "gsd_gnn_aggregate_bwd_workspace_bytes")
        with torch.cuda.device(A.device):
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=A.device)
            gA = torch.zeros_like(A)       # unused edge slots keep a zero gradient
            gP = torch.empty_like(P)
            _lib.check(lib.gsd_gnn_aggregate_bwd(B, N, cap, Fd, e.n_tool, e.row_ptr.data_ptr(), e.senders.data_ptr(), col_ptr.data_ptr(),
                                                 order.data_ptr(), A.data_ptr(), P.data_ptr(), g.data_ptr(), ws.data_ptr(),
                                                 gA.data_ptr(), gP.data_ptr(), _stream()), "gsd_gnn_aggregate_bwd")
        return gA, gP, None


# ----------------------------------------------------------------------------------------------------
# model (state_dict identical to the reference: particle_encoder.model.{0,2,4}, relation_encoder.model.{0,2,4},
# particle_propagator.linear, relation_propagator.linear, non_rigid_predictor.linear_{0,1,2})
# ----------------------------------------------------------------------------------------------------
class Encoder(nn.Module):
    def __init__(self, input_size, hidden_size, output_size):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(input_size, hidden_size), nn.ReLU(), nn.Linear(hidden_size, hidden_size), nn.ReLU(),
                                   nn.Linear(hidden_size, output_size), nn.ReLU())
        self.output_size = output_size

    def forward(self, x):
        s = x.size()
        return self.model(x.reshape(-1, s[-1])).view(list(s[:-1]) + [self.output_size])


class Propagator(nn.Module):
    def __init__(self, input_size, output_size):
        super().__init__()
        self.linear = nn.Linear(input_size, output_size)
        self.relu = nn.ReLU()
        self.output_size = output_size


class ParticlePredictor(nn.Module):
    def __init__(self, input_size, hidden_size, output_size):
        super().__init__()
        self.linear_0 = nn.Linear(input_size, hidden_size)
        self.linear_1 = nn.Linear(hidden_size, hidden_size)
        self.linear_2 = nn.Linear(hidden_size, output_size)
        self.relu = nn.ReLU()
        self.output_size = output_size

    def forward(self, x):
        s = x.size()
        x = x.reshape(-1, s[-1])
        x = self.relu(self.linear_0(x))
        x = self.relu(self.linear_1(x))
        return self.linear_2(x).view(list(s[:-1]) + [self.output_size])


class DynamicsPredictor(nn.Module):
    def __init__(self, model_config, device, matmul="tc"):
        super().__init__()
        self.model_config = model_config
        self.device = device
        # "tc": hand-written tcgen05 3xTF32 GEMMs with fused epilogues (inference) / _Linear3x (training);
        # "3xtf32": round 1's cuBLAS path over K-concatenated packed operands (kept for A/B measurements); "ieee": fp32 SIMT
        self.matmul = matmul
        self.nf_particle = model_config['nf_particle']
        self.nf_relation = model_config['nf_relation']
        self.nf_effect = model_config['nf_effect']
        self.motion_clamp = 100.0
        self.motion_dim = model_config['motion_dim'] if 'motion_dim' in model_config else 0
        input_dim = model_config['n_his'] * model_config['state_dim'] + (model_config['n_his'] - 1) * self.motion_dim + \
            model_config['attr_dim'] + model_config['action_dim']
        self.particle_encoder = Encoder(input_dim, self.nf_particle, self.nf_effect)
        rel_input_dim = model_config['rel_attr_dim'] * 2 + model_config['rel_group_dim'] + \
            model_config['rel_distance_dim'] * model_config['n_his']
        self.relation_encoder = Encoder(rel_input_dim, self.nf_relation, self.nf_effect)
        self.particle_propagator = Propagator(self.nf_effect * 2, self.nf_effect)
        self.relation_propagator = Propagator(self.nf_effect * 3, self.nf_effect)
        self.non_rigid_predictor = ParticlePredictor(self.nf_effect, self.nf_effect, 3)
        if model_config.get('verbose', False):
            print("DynamicsPredictor initialized")
            print("particle input dim: {}, relation input dim: {}".format(input_dim, rel_input_dim))

    def _particle_inputs(self, state, attrs, action):
        cfg = self.model_config
        B, N = attrs.size(0), attrs.size(1)
        n_his, state_dim = cfg['n_his'], state.size(3)
        state_t = state.transpose(1, 2).contiguous().view(B, N, n_his * state_dim)
        p_inputs = attrs
        if cfg['state_dim'] > 0:
            if cfg['state_dim'] == 3:
                p_inputs = torch.cat([p_inputs, state_t], 2)
            elif cfg['state_dim'] == 1:
                assert state_dim == 3
                p_inputs = torch.cat([attrs, state_t.view(B, N, n_his, state_dim)[:, :, :, 2]], 2)
        if self.motion_dim > 0:
            assert self.motion_dim == 3
            xyz = state_t.view(B, N, n_his, state_dim)
            p_inputs = torch.cat([p_inputs, (xyz[:, :, 1:] - xyz[:, :, :-1]).reshape(B, N, (n_his - 1) * 3)], 2)
        if cfg['action_dim'] > 0:
            assert action is not None
            p_inputs = torch.cat([p_inputs, action], 2)
        return p_inputs

    def _forward_ieee(self, p_inputs, rel_inputs, edges, B, N, n_p, fast=False):
        """Differentiable path (training and the bit-tight parity tests).  fast=False: fp32 SIMT GEMMs; fast=True: the F-wide
        layers run as error-compensated TF32 tensor-core GEMMs in forward and backward (_Linear3x), same results to ~1e-6."""
        cfg, Fd = self.model_config, self.nf_effect
        pe, re, nr = self.particle_encoder.model, self.relation_encoder.model, self.non_rigid_predictor
        mlp3 = lambda m, x: torch.relu(_linear(torch.relu(_linear(torch.relu(F.linear(x, m[0].weight, m[0].bias)), m[2].weight,
                                                                      m[2].bias, fast)), m[4].weight, m[4].bias, fast))
        particle_encode = mlp3(pe, p_inputs.reshape(B * N, -1))
        relation_encode = mlp3(re, rel_inputs.reshape(B * edges.capacity, -1))
        Wr, br = self.relation_propagator.linear.weight, self.relation_propagator.linear.bias
        Wp, bp = self.particle_propagator.linear.weight, self.particle_propagator.linear.bias
        A = _linear(relation_encode, Wr[:, :Fd].contiguous() if fast else Wr[:, :Fd], br, fast)   # pstep-invariant edge term
        C0 = _linear(particle_encode, Wp[:, :Fd].contiguous() if fast else Wp[:, :Fd], bp, fast)  # pstep-invariant node term
        W23 = torch.cat([Wr[:, Fd:2 * Fd], Wr[:, 2 * Fd:]], 0)        # [2F, F]: receiver | sender projections
        Wp2 = Wp[:, Fd:].contiguous() if fast else Wp[:, Fd:]
        h = particle_encode
        for _ in range(cfg['pstep']):
            P = _linear(h, W23, None, fast)                           # [B*N, 2F]
            agg = _Aggregate.apply(A, P, edges)
            h = torch.relu(C0 + h + _linear(agg, Wp2, None, fast))
        x = h.view(B, N, Fd)[:, :n_p].reshape(B * n_p, Fd)
        x = torch.relu(_linear(x, nr.linear_0.weight, nr.linear_0.bias, fast))
        x = torch.relu(_linear(x, nr.linear_1.weight, nr.linear_1.bias, fast))
        return F.linear(x, nr.linear_2.weight, nr.linear_2.bias).view(B, n_p, 3)

    def _packed_weights(self):
        """[w_hi | w_lo | w_hi] operands of every F-wide layer, rebuilt when a parameter changes (version counters)."""
        ps = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if getattr(self, "_wkey", None) != key:
            Fd = self.nf_effect
            Wr, Wp = self.relation_propagator.linear.weight.detach(), self.particle_propagator.linear.weight.detach()
            pk = lambda w: _tf32_pack(w.detach(), weight=True)[0].t()
            pe, re, nr = self.particle_encoder.model, self.relation_encoder.model, self.non_rigid_predictor
            self._wpk = dict(pe2=pk(pe[2].weight), pe4=pk(pe[4].weight), re2=pk(re[2].weight), re4=pk(re[4].weight),
                             A=pk(Wr[:, :Fd]), C0=pk(Wp[:, :Fd]), W23=pk(torch.cat([Wr[:, Fd:2 * Fd], Wr[:, 2 * Fd:]], 0)),
                             Wp2=pk(Wp[:, Fd:]), nr0=pk(nr.linear_0.weight), nr1=pk(nr.linear_1.weight))
            self._wkey = key
        return self._wpk

    def _split_weights(self):
        """(w_hi, w_lo) of every F-wide layer for the tcgen05 path, rebuilt when a parameter changes (version counters)."""
        ps = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if getattr(self, "_tckey", None) != key:
            Fd = self.nf_effect
            Wr, Wp = self.relation_propagator.linear.weight.detach(), self.particle_propagator.linear.weight.detach()
            pe, re, nr = self.particle_encoder.model, self.relation_encoder.model, self.non_rigid_predictor
            sp = _tc_split
            self._tcw = dict(pe2=sp(pe[2].weight), pe4=sp(pe[4].weight), re2=sp(re[2].weight), re4=sp(re[4].weight),
                             A=sp(Wr[:, :Fd]), C0=sp(Wp[:, :Fd]), W23=sp(torch.cat([Wr[:, Fd:2 * Fd], Wr[:, 2 * Fd:]], 0)),
                             Wp2=sp(Wp[:, Fd:]), nr0=sp(nr.linear_0.weight), nr1=sp(nr.linear_1.weight))
            self._tckey = key
        return self._tcw

    def _node_chain_tc(self, p_inputs, B, N):
        """Particle encoder + the pstep-invariant node term + the first receiver | sender projection, launched on a SIDE stream:
        they depend only on the particle inputs, so they run beside the edge builder and the edge-row layers (whose persistent
        kernels leave SMs idle in their last wave) and are joined before the first aggregation.  Returns (h, C0, P, side stream)."""
        W = self._split_weights()
        pe, bp = self.particle_encoder.model, self.particle_propagator.linear.bias
        main = torch.cuda.current_stream()
        side = getattr(self, "_side", None)
        if side is None or side.device != p_inputs.device:
            side = self._side = torch.cuda.Stream(device=p_inputs.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            y = _small_linear(p_inputs.reshape(B * N, -1), pe[0].weight, pe[0].bias, relu=True)
            y = _tc_linear(y, W["pe2"], pe[2].bias, relu=True)
            h = _tc_linear(y, W["pe4"], pe[4].bias, relu=True)                   # particle_encode
            C0 = _tc_linear(h, W["C0"], bp)                                      # pstep-invariant node term
            P = _tc_linear(h, W["W23"])                                          # [B*N, 2F]: receiver | sender projections
        if not torch.cuda.is_current_stream_capturing():      # eager: tell the caching allocator about the cross-stream uses
            p_inputs.record_stream(side)
            for t in (h, C0, P):
                t.record_stream(main)
        return h, C0, P, side

    def _forward_tc(self, p_inputs, rel_inputs, edges, B, N, n_p, node=None):
        """Inference path on the hand-written tensor-core kernels: every F-wide layer is one gsd_linear_tf32x3 launch whose
        epilogue carries the bias, the residual adds and the ReLU (model.py:202-241 with the relation propagator's weight split
        of DESIGN.md §4); the K = 5 / 14 input layers and the 512 -> 3 head are plain fp32 kernels.  No library GEMM, no pack
        launches, no elementwise launches between layers.  The node-level chain runs on a side stream (`node`: already
        launched by the caller, e.g. before the edge builder)."""
        cfg, Fd = self.model_config, self.nf_effect
        W = self._split_weights()
        re, nr = self.relation_encoder.model, self.non_rigid_predictor
        br = self.relation_propagator.linear.bias
        if node is None:
            node = self._node_chain_tc(p_inputs, B, N)
        h, C0, P, side = node
        e = _small_linear(rel_inputs.reshape(B * edges.capacity, -1), re[0].weight, re[0].bias, relu=True)
        e = _tc_linear(e, W["re2"], re[2].bias, relu=True)
        e = _tc_linear(e, W["re4"], re[4].bias, relu=True)                       # relation_encode
        A = _tc_linear(e, W["A"], br)                                            # pstep-invariant edge term
        torch.cuda.current_stream().wait_stream(side)
        for i in range(cfg['pstep']):
            if i > 0:
                P = _tc_linear(h, W["W23"])
            agg = _Aggregate.apply(A, P, edges)
            h = _tc_linear(agg, W["Wp2"], None, res1=C0, res2=h, relu=True)      # relu(C0 + h + agg Wp2^T)
        x = h if n_p == N else h.view(B, N, Fd)[:, :n_p].reshape(B * n_p, Fd)
        x = _tc_linear(x, W["nr0"], nr.linear_0.bias, relu=True)
        x = _tc_linear(x, W["nr1"], nr.linear_1.bias, relu=True)
        return _small_linear(x, nr.linear_2.weight, nr.linear_2.bias).view(B, n_p, 3)

    def _forward_3xtf32(self, p_inputs, rel_inputs, edges, B, N, n_p):
        """Inference path: every F-wide layer is one TF32 tensor-core GEMM over the error-compensated K = 3F operands
        (gsd_tf32_pack), which also carries the ReLU / residual add between layers.  Same results as the fp32 path to ~1e-6."""
        cfg, Fd = self.model_config, self.nf_effect
        W = self._packed_weights()
        pe, re, nr = self.particle_encoder.model, self.relation_encoder.model, self.non_rigid_predictor
        br, bp = self.relation_propagator.linear.bias, self.particle_propagator.linear.bias
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        y = F.linear(p_inputs.reshape(B * N, -1), pe[0].weight, pe[0].bias)        # K = 5..14: fp32
        e = F.linear(rel_inputs.reshape(B * edges.capacity, -1), re[0].weight, re[0].bias)
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            y = torch.addmm(pe[2].bias, _tf32_pack(y, relu=True)[0], W["pe2"])
            y = torch.addmm(pe[4].bias, _tf32_pack(y, relu=True)[0], W["pe4"])
            hp, h = _tf32_pack(y, relu=True, want_full=True)                        # particle_encode
            C0 = torch.addmm(bp, hp, W["C0"])
            e = torch.addmm(re[2].bias, _tf32_pack(e, relu=True)[0], W["re2"])
            e = torch.addmm(re[4].bias, _tf32_pack(e, relu=True)[0], W["re4"])
            A = torch.addmm(br, _tf32_pack(e, relu=True)[0], W["A"])
            for _ in range(cfg['pstep']):
                P = hp @ W["W23"]
                agg = _Aggregate.apply(A, P, edges)
                y = torch.addmm(C0, _tf32_pack(agg)[0], W["Wp2"])
                hp, h = _tf32_pack(y, relu=True, add=h, want_full=True)
            if n_p != N:
                hp = hp.view(B, N, 3 * Fd)[:, :n_p].reshape(B * n_p, 3 * Fd)
            y = torch.addmm(nr.linear_0.bias, hp, W["nr0"])
            y = torch.addmm(nr.linear_1.bias, _tf32_pack(y, relu=True)[0], W["nr1"])
            torch.backends.cuda.matmul.allow_tf32 = False
            return F.linear(torch.relu(y), nr.linear_2.weight, nr.linear_2.bias).view(B, n_p, 3)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    def forward(self, state, attrs, Rr, Rs, p_instance, action=None, **kwargs):
        """Rr may be an EdgeIndex (fast path, Rs ignored) or the reference's dense one-hot matrices."""
        cfg = self.model_config
        if cfg['rel_attr_dim'] != attrs.size(2) or cfg['rel_group_dim'] != 1 or cfg['rel_distance_dim'] != 3:
            raise NotImplementedError("only the reference's relation layout (attr, group diff, 3-d distance) is implemented")
        state = _cuda_f32(state, "state")
        attrs = _cuda_f32(attrs, "attrs")
        p_instance = _cuda_f32(p_instance, "p_instance")
        edges = Rr if isinstance(Rr, EdgeIndex) else edge_index_from_dense(Rr, Rs)
        B, N = attrs.size(0), attrs.size(1)
        n_p = p_instance.size(1)
        Fd = self.nf_effect
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            p_inputs = self._particle_inputs(state, attrs, action)
            rel_inputs = edge_inputs(state, attrs, p_instance, edges)
            if self.matmul == "tc" and not torch.is_grad_enabled() and Fd % 64 == 0:
                pred_motion = self._forward_tc(p_inputs, rel_inputs, edges, B, N, n_p)
            elif self.matmul == "3xtf32" and not torch.is_grad_enabled() and B * N >= 256 and Fd % 4 == 0:
                pred_motion = self._forward_3xtf32(p_inputs, rel_inputs, edges, B, N, n_p)
            else:
                pred_motion = self._forward_ieee(p_inputs, rel_inputs, edges, B, N, n_p,
                                                 fast=self.matmul in ("tc", "3xtf32") and torch.is_grad_enabled() and Fd % 4 == 0)
            pred_pos = state[:, -1, :n_p] + torch.clamp(pred_motion, max=self.motion_clamp, min=-self.motion_clamp)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev_tf32
        return pred_pos, pred_motion


def downsample_vertices(xyz, max_nobj, fps_radius, start_idx=None):
    """DynamicsModule.downsample_vertices (/root/reference/src/render/dynamics_module.py:44-51): FPS to max_nobj points
    (start index 0), then radius-terminated FPS. Returns (points, indices into xyz)."""
    idx1 = farthest_point_sampler(xyz[None], min(max_nobj, xyz.shape[0]), start_idx=0)[0]
    sub = xyz[idx1]
    _, idx2 = fps_rad_idx_torch(sub, fps_radius, start_idx=start_idx)
    idx = idx1[idx2]
    return xyz[idx], idx


# ----------------------------------------------------------------------------------------------------
# rollout step (the loop body of DynamicsModule.rollout, /root/reference/src/render/dynamics_module.py:85-170,
# without the Gaussian skinning which is a "next" row): edges -> model -> history shift
# ----------------------------------------------------------------------------------------------------
class GnnRollout:
    """Autoregressive rollout with static shapes (CUDA-graph capturable): nobj object particles + 1 tool particle.
    batch > 1 rolls `batch` action samples out of the same initial state at once (the MPPI planner's inner loop,
    /root/reference/src/real_world/plan.py:25-154: bsz perturbed action sequences, graph rebuilt every step)."""

    def __init__(self, model, particle_pos, eef_pos, adj_thresh, topk, connect_all, use_graph=True, batch=1):
        self.model = model
        dev = particle_pos.device
        n_his = model.model_config['n_his']
        self.nobj = particle_pos.shape[-2]
        self.B = B = int(batch)
        N = self.nobj + 1
        self.states = torch.zeros((B, n_his, N, 3), device=dev)
        self.states[:, :, :self.nobj] = particle_pos          # [nobj,3] or a history [n_his,nobj,3]
        self.states[:, :, self.nobj:] = eef_pos.reshape(-1, 1, 3) if eef_pos.dim() > 1 else eef_pos
        self.action = torch.zeros((B, N, 3), device=dev)
        self.attrs = torch.zeros((B, N, 2), device=dev)
        self.attrs[:, :self.nobj, 0] = 1.
        self.attrs[:, self.nobj:, 1] = 1.
        self.p_instance = torch.ones((B, self.nobj, 1), device=dev)
        self.state_mask = torch.ones((B, N), dtype=torch.bool, device=dev)
        self.eef_mask = torch.zeros((B, N), dtype=torch.bool, device=dev)
        self.eef_mask[:, self.nobj] = True
        self.adj_thresh, self.topk, self.connect_all = adj_thresh, topk, connect_all
        self.eef_delta = torch.zeros(3, device=dev) if B == 1 else torch.zeros((B, 3), device=dev)
        self.use_graph, self.graph, self.pred = use_graph, None, None
        self.gs_xyz = self.gs_quat = None

    def attach_gaussians(self, xyz, quat):
        """Skin these Gaussians (positions [n,3], rotations [n,4] wxyz) from the particles after every step
        (interpolate_motions, dynamics_module.py:150-156); `gs_xyz` / `gs_quat` hold the current values.  Call before the first step."""
        if self.graph is not None:
            raise RuntimeError("attach_gaussians must be called before the first step")
        if self.B != 1:
            raise ValueError("Gaussian skinning follows a single rollout (batch == 1)")
        self.gs_xyz, self.gs_quat = xyz.detach().clone().float().contiguous(), quat.detach().clone().float().contiguous()

    def _fused_glue_ok(self):
        """The two glue kernels cover the reference's rollout configurations (state_dim 0 / 1 / 3, motion_dim 0 / 3, action_dim
        0 / 3, one tool node) on the tensor-core inference path; anything else takes the torch-op path below (same results)."""
        m, cfg = self.model, self.model.model_config
        return (m.matmul == "tc" and m.nf_effect % 64 == 0 and cfg['state_dim'] in (0, 1, 3) and m.motion_dim in (0, 3)
                and cfg['action_dim'] in (0, 3) and cfg['rel_attr_dim'] == self.attrs.size(2) and cfg['rel_group_dim'] == 1
                and cfg['rel_distance_dim'] == 3 and os.environ.get("GSD_ROLLOUT_GLUE", "1") != "0")

    @torch.no_grad()
    def _step_fused(self):
        """One step with the tensor plumbing in two kernels (gsd_gnn_rollout_pre / _post) instead of ~14 elementwise / cat /
        copy launches: pre -> edges -> relation inputs -> dense layers + aggregation -> post."""
        m, cfg = self.model, self.model.model_config
        B, N, nobj, n_his = self.B, self.nobj + 1, self.nobj, cfg['n_his']
        dev = self.states.device
        has_action = int(cfg['action_dim'] > 0)
        width = self.attrs.size(2) + cfg['state_dim'] * n_his + (3 * (n_his - 1) if m.motion_dim > 0 else 0) + 3 * has_action
        lib = _lib.lib()
        if getattr(self, "_m8", None) is None:
            self._m8, self._t8 = self.state_mask.to(torch.uint8).contiguous(), self.eef_mask.to(torch.uint8).contiguous()
        with torch.cuda.device(dev):
            p_inputs = torch.empty((B * N, width), device=dev)
            cur = torch.empty((B, N, 3), device=dev)
            pred = torch.empty((B, nobj, 3), device=dev)
            _lib.check(lib.gsd_gnn_rollout_pre(B, N, nobj, n_his, self.attrs.size(2), int(cfg['state_dim']), int(m.motion_dim > 0), has_action,
                                               self.states.data_ptr(), self.attrs.data_ptr(), self.action.data_ptr(),
                                               self.eef_delta.data_ptr(), 0 if self.eef_delta.dim() == 1 else 3,
                                               p_inputs.data_ptr(), cur.data_ptr(), _stream()), "gsd_gnn_rollout_pre")
            node = m._node_chain_tc(p_inputs, B, N)        # side stream: beside the edge builder and the edge-row layers
            edges = construct_edges_index(cur, self.adj_thresh, self._m8, self._t8, topk=self.topk, connect_all=self.connect_all, n_tool=1)
            rel_inputs = edge_inputs(self.states, self.attrs, self.p_instance, edges)
            motion = m._forward_tc(p_inputs, rel_inputs, edges, B, N, nobj, node=node)
            _lib.check(lib.gsd_gnn_rollout_post(B, N, nobj, n_his, self.states.data_ptr(), motion.data_ptr(), self.eef_delta.data_ptr(),
                                                0 if self.eef_delta.dim() == 1 else 3, float(m.motion_clamp), pred.data_ptr(), _stream()),
                       "gsd_gnn_rollout_post")
        if self.gs_xyz is not None:
            from .skinning import interpolate_motions
            bones = cur[0, :nobj]                                          # the positions before this step
            x, q, _ = interpolate_motions(bones, pred[0] - bones, edges, self.gs_xyz, quat=self.gs_quat, return_weights=False)
            self.gs_xyz.copy_(x)
            self.gs_quat.copy_(q)
        return pred

    @torch.no_grad()
    def _step_impl(self):
        if self._fused_glue_ok():
            return self._step_fused()
        # tool moves by eef_delta; history shift of the tool row; action row of the tool
        new_eef = self.states[:, -1, self.nobj] + self.eef_delta          # [B,3]
        self.action[:, self.nobj] = self.eef_delta
        edges = construct_edges_index(self.states[:, -1], self.adj_thresh, self.state_mask, self.eef_mask, topk=self.topk,
                                      connect_all=self.connect_all, n_tool=1)
        pred, _ = self.model(self.states, self.attrs, edges, None, self.p_instance, action=self.action)
        if self.gs_xyz is not None:
            from .skinning import interpolate_motions
            bones = self.states[0, -1, :self.nobj]
            x, q, _ = interpolate_motions(bones, pred[0] - bones, edges, self.gs_xyz, quat=self.gs_quat, return_weights=False)
            self.gs_xyz.copy_(x)
            self.gs_quat.copy_(q)
        nxt = torch.cat([pred, new_eef[:, None]], 1)                      # [B,N,3]
        self.states.copy_(torch.cat([self.states[:, 1:], nxt[:, None]], 1))
        return pred

    def step(self, eef_delta=None):
        with _nvtx.range("gsd.gnn_rollout_step"):
            return self._step(eef_delta)

    def _step(self, eef_delta=None):
        if eef_delta is not None:
            self.eef_delta.copy_(eef_delta)
        if not self.use_graph:
            return self._step_impl()
        if self.graph is None:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            keep = self.states.clone()
            keep_gs = (self.gs_xyz.clone(), self.gs_quat.clone()) if self.gs_xyz is not None else None

            def restore():
                self.states.copy_(keep)
                if keep_gs is not None:
                    self.gs_xyz.copy_(keep_gs[0])
                    self.gs_quat.copy_(keep_gs[1])
            with torch.cuda.stream(s):
                self._step_impl()
            torch.cuda.current_stream().wait_stream(s)
            restore()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.pred = self._step_impl()
            restore()
        self.graph.replay()
        return self.pred
