"""GNN training step on the fused kernels — mirror of the inner loop of /root/reference/src/train.py:105-261.

The reference unrolls ``n_future`` model calls on a fixed graph (``data['Rr'], data['Rs']`` are not rebuilt inside an
iteration, train.py:183-211), feeds each prediction back as the newest history frame, sums ``weight * loss`` over the steps
and back-propagates through the whole unroll; ``length_loss`` / ``local_rigid_loss`` gather edge endpoints with one-hot bmm
(train.py:66-102).  Here the graph is an :class:`EdgeIndex` (index lists), the model's gather / segment-reduce and their
backward run in the CUDA kernels of ``csrc/gnn.cu`` / ``csrc/gnn_bwd.cu``, the edge losses index instead of multiplying by
one-hots, and data-parallel training averages one flat fp32 gradient bucket with a single all-reduce (NCCL over NVLink).
Same function names, argument meaning and results as the reference's loss functions.
"""
import torch
import torch.nn.functional as F

from . import dist as gdist
from .gnn import EdgeIndex, edge_index_from_dense, split_scope as gnn_split_scope


def mse_loss(pred, gt, **kwargs):
    """train.py:59-60"""
    return F.mse_loss(pred, gt)


def l1_loss(pred, gt, **kwargs):
    """train.py:62-63"""
    return F.l1_loss(pred, gt)


def _edges_of(kwargs):
    e = kwargs.get('edges')
    if e is None:
        Rr = kwargs['Rr']
        e = Rr if isinstance(Rr, EdgeIndex) else edge_index_from_dense(Rr, kwargs['Rs'])
    return e


def _endpoints(x, edges, n_p):
    """(x[recv], x[send], weight) per edge slot, restricted to the first n_p nodes exactly like ``Rr[:, :, :n_p].bmm(x)``:
    an endpoint outside [0, n_p) (tool node) contributes a zero row; unused slots (the reference's zero padding rows of
    Rr/Rs) are rows of zeros too and DO count in the reference's mean — they are kept, with both endpoints zero."""
    r, s = edges.receivers.long(), edges.senders.long()
    ok_r = (r >= 0) & (r < n_p)
    ok_s = (s >= 0) & (s < n_p)
    idx_r = r.clamp(0, n_p - 1)[..., None].expand(-1, -1, x.shape[-1])
    idx_s = s.clamp(0, n_p - 1)[..., None].expand(-1, -1, x.shape[-1])
    xr = torch.gather(x, 1, idx_r) * ok_r[..., None]
    xs = torch.gather(x, 1, idx_s) * ok_s[..., None]
    return xr, xs


def _safe_norm(d):
    # torch.norm's subgradient at 0 is 0 (padding rows / self edges); sqrt of the clamped sum reproduces it
    sq = (d * d).sum(-1)
    pos = sq > 0
    return torch.where(pos, torch.sqrt(torch.where(pos, sq, torch.ones_like(sq))), torch.zeros_like(sq))


def _edge_mean(sq, edges, kwargs):
    """Mean over the relation rows the way train.py:66-102 takes it: F.mse_loss over [B, n_rel_padded] where the dataset pads every
    sample's one-hot matrices to `max_nR` rows (data/dataset.py:344,491-492) and the padding rows contribute 0.  Unused edge slots
    contribute 0 here too, so only the DENOMINATOR depends on the padding: pass max_nR=<the reference config's value> in the batch
    dict to get the reference's loss scale with a native EdgeIndex (default: the index's own capacity — identical when the edges
    come from the reference's dense matrices through edge_index_from_dense)."""
    denom = int(kwargs['max_nR']) if kwargs.get('max_nR') else sq.shape[1]
    return sq.sum() / float(sq.shape[0] * denom)


def length_loss(pred, gt, **kwargs):
    """MSE between the edge lengths of the prediction and of the oldest history frame (train.py:66-83)."""
    n_p = pred.shape[1]
    pos = kwargs['state'][:, 0, :n_p].detach()
    edges = _edges_of(kwargs)
    pos_r, pos_s = _endpoints(pos, edges, n_p)
    pred_r, pred_s = _endpoints(pred, edges, n_p)
    return _edge_mean((_safe_norm(pred_r - pred_s) - _safe_norm(pos_r - pos_s)) ** 2, edges, kwargs)


def local_rigid_loss(pred, gt, **kwargs):
    """train.py:85-102"""
    n_p = pred.shape[1]
    pos = kwargs['state'][:, 0, :n_p].detach()
    edges = _edges_of(kwargs)
    pos_r, pos_s = _endpoints(pos, edges, n_p)
    pred_r, pred_s = _endpoints(pred, edges, n_p)
    return _edge_mean((_safe_norm(pred_r - pos_r) - _safe_norm(pred_s - pos_s)) ** 2, edges, kwargs)


def umeyama_algorithm(X, Y, mask, fixed_scale=True):
    """Least-squares similarity transform Y ~ c R X + t over the masked points (gnn/utils.py:7-40; Umeyama 1991).
    X, Y [B,N,3], mask [B,N] -> c [B], R [B,3,3], t [B,1,3]."""
    w = mask.to(X.dtype)
    n = w.sum(1)
    mean = lambda P: torch.einsum('bn,bnk->bk', w, P)[:, None, :] / n[:, None, None]
    mu_x, mu_y = mean(X), mean(Y)
    Xc, Yc = (X - mu_x) * w[..., None], (Y - mu_y) * w[..., None]
    cov = torch.einsum('bni,bnj->bij', Yc, Xc) / n[:, None, None]
    U, D, Vh = torch.linalg.svd(cov)
    flip = (torch.linalg.det(U) * torch.linalg.det(Vh) < 0).to(X.dtype)
    sign = torch.ones_like(D)
    sign[:, -1] = 1.0 - 2.0 * flip                                   # reflect the last singular direction when det < 0
    U = U * sign[:, None, :]
    if fixed_scale:
        c = torch.ones_like(n)
    else:
        var_x = torch.einsum('bn,bn->b', w, ((X - mu_x) ** 2).sum(-1)) / n
        c = (D * sign).sum(1) / var_x
    R = U @ Vh
    t = mu_y - c[:, None, None] * (mu_x @ R.transpose(1, 2))
    return c, R, t


def rigid_loss(pred, gt, **kwargs):
    """MSE between the prediction and the best rigid motion of the oldest history frame (train.py:30-38); the aligned
    target is detached.  Masked mean instead of boolean indexing (no host sync); same value."""
    n_p = pred.shape[1]
    orig = kwargs['state'][:, 0, :n_p]
    mask = kwargs['obj_mask']
    _, R, t = umeyama_algorithm(orig, pred, mask, fixed_scale=True)
    target = (orig @ R.transpose(1, 2) + t).detach()
    w = mask.to(pred.dtype)[..., None]
    return (((pred - target) ** 2) * w).sum() / (w.sum() * pred.shape[-1])


def default_loss_funcs(train_config):
    """train.py:141-155"""
    if train_config.get('rigid_loss'):
        return [(mse_loss, 1.0), (length_loss, 0.05), (rigid_loss, 0.05)]
    funcs = [(mse_loss, train_config['mse_loss'] if train_config.get('mse_loss', 0) > 0 else 1.0)]
    funcs.append((length_loss, train_config['length_loss'] if train_config.get('length_loss', 0) > 0 else 0.01))
    return funcs


def unrolled_loss(model, data, n_future, loss_funcs):
    """The n_future-step unroll of train.py:183-211 on one batch.  ``data``: dict with state [B,n_his,N,3], attrs, Rr (EdgeIndex
    or dense) / Rs, p_instance, action, state_future [B,n_future,n_p,3], tool_future / action_future [B,n_future-1,N,3].
    Returns (loss_sum, [per-function weighted losses of every step]); ``data`` is not modified."""
    data = dict(data)
    if not isinstance(data['Rr'], EdgeIndex):
        data['Rr'] = edge_index_from_dense(data['Rr'], data['Rs'])
    data['edges'] = data['Rr']
    loss_sum = 0
    parts = []
    for fi in range(n_future):
        gt_state = data['state_future'][:, fi]
        pred_state, _ = model(**data)
        pred_state_p = pred_state[:, :gt_state.shape[1], :3]
        loss = [w * f(pred_state_p, gt_state, **data) for f, w in loss_funcs]
        loss_sum = loss_sum + sum(loss)
        parts.append(loss)
        if fi < n_future - 1:
            next_state = data['tool_future'][:, fi].clone().unsqueeze(1)
            next_state[:, -1, :pred_state_p.shape[1]] = pred_state_p
            data['state'] = torch.cat([data['state'][:, 1:], next_state], dim=1)
            data['action'] = data['action_future'][:, fi]
    return loss_sum, parts


class GradientBucket:
    """One flat fp32 buffer aliasing every parameter gradient: data-parallel training all-reduces it once per step
    (2.9 M parameters = 11.6 MB at nf = 512) instead of once per tensor."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()
        for p, g in zip(self.params, self._views()):
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g

    def _views(self):
        o = 0
        for p in self.params:
            yield self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def all_reduce_mean(self):
        """One collective over the flat buffer.  NCCL averages inside the collective (ReduceOp.AVG, no extra pass over the
        buffer); gloo (the CPU tests) has no AVG, so sum + scale there."""
        if gdist.world_size() > 1:
            if torch.distributed.get_backend() == "nccl":
                torch.distributed.all_reduce(self.flat, op=torch.distributed.ReduceOp.AVG)
            else:
                torch.distributed.all_reduce(self.flat, op=torch.distributed.ReduceOp.SUM)
                self.flat.mul_(1.0 / gdist.world_size())


def train_iteration(model, optimizer, data, n_future, loss_funcs, bucket=None):
    """optimizer.zero_grad(); unroll; backward; (DP: average the gradient bucket); optimizer.step() — train.py:176-218."""
    if bucket is not None:
        bucket.zero()
    else:
        optimizer.zero_grad()
    with gnn_split_scope():      # the weights' hi / lo splits are shared by the unroll steps and the backward of this iteration
        loss_sum, parts = unrolled_loss(model, data, n_future, loss_funcs)
        loss_sum.backward()
    if bucket is not None:
        bucket.all_reduce_mean()
    optimizer.step()
    return loss_sum.detach(), parts


class GraphedTrainStep:
    """train_iteration captured in ONE CUDA graph: bucket.zero, the n_future unroll, backward, the gradient all-reduce (NCCL
    collectives are capturable) and a capturable Adam step.  The eager iteration is CPU-launch-bound (hundreds of small kernels of
    the unroll); replaying a graph removes the launch cost and, across ranks, the skew it causes in front of the all-reduce
    (round 1 measured 31.1 ms -> 34.0 ms per iteration from 1 to 8 GPUs for an 11.6 MB collective that takes ~30 us on NVLink).
    `data` holds the static batch buffers: copy the next batch into them (load_batch) before each step."""

    def __init__(self, model, optimizer, data, n_future, loss_funcs, bucket, warmup=3):
        self.model, self.optimizer, self.data, self.n_future, self.funcs, self.bucket = model, optimizer, data, n_future, loss_funcs, bucket
        for g in optimizer.param_groups:
            if not g.get('capturable', False):
                raise ValueError("GraphedTrainStep needs torch.optim.Adam(..., capturable=True)")
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        self.warmup_losses = []
        with torch.cuda.stream(s):
            for _ in range(warmup):   # real iterations (allocator + NCCL communicator warm-up); the caller counts them
                self.warmup_losses.append(train_iteration(model, optimizer, data, n_future, loss_funcs, bucket)[0])
        torch.cuda.current_stream().wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, _ = train_iteration(model, optimizer, data, n_future, loss_funcs, bucket)

    def load_batch(self, batch):
        for k, v in batch.items():
            if isinstance(v, torch.Tensor) and isinstance(self.data.get(k), torch.Tensor):
                self.data[k].copy_(v, non_blocking=True)

    def step(self):
        self.graph.replay()
        return self.loss
