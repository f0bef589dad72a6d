"""Linear-blend skinning of the Gaussians from the GNN particles — host-side mirror of the reference's
`render/utils.py` (/root/reference/src/render/utils.py:52-243): `interpolate_motions`, `relations_to_matrix`, `mat2quat`,
`quat2mat`, same names / argument meaning / results, on the CUDA kernels of csrc/skinning.cu through the C ABI.

No CPU path: tensors must be CUDA tensors (the reference's `device` argument is accepted and must be a CUDA device)."""
import ctypes as C

import torch

from . import _lib
from .gnn import EdgeIndex

_BONE_TF = 20  # floats per bone record, include/gsd.h: gsd_skin_bone_transforms


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, name):
    if not torch.is_tensor(t) or not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (no CPU path)" % name)
    return t.detach().contiguous().float()


def quat2mat(q):
    """render/utils.py:52-66: (w, x, y, z) -> rotation matrices [n,3,3]; q is normalised first."""
    q = q / torch.sqrt((q * q).sum(-1, keepdim=True))
    r, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).view(-1, 3, 3)


def mat2quat(rot):
    """render/utils.py:68-109: rotation matrices [n,3,3] -> (w, x, y, z), the reference's four branches, without its masked
    writes (branch-free selects)."""
    r = rot
    t = torch.clamp(r[:, 0, 0] + r[:, 1, 1] + r[:, 2, 2], min=-1)
    m0 = t > -1
    m1 = ~m0 & (r[:, 0, 0] >= r[:, 1, 1]) & (r[:, 0, 0] >= r[:, 2, 2])
    m2 = ~m0 & (r[:, 1, 1] >= r[:, 2, 2]) & (r[:, 1, 1] > r[:, 0, 0])
    one = torch.ones_like(t)

    def safe_sqrt(v, m):
        return torch.sqrt(torch.where(m, v, one))
    s0 = safe_sqrt(t + 1, m0)
    q_0 = torch.stack([0.5 * s0, (r[:, 2, 1] - r[:, 1, 2]) * (0.5 / s0), (r[:, 0, 2] - r[:, 2, 0]) * (0.5 / s0),
                       (r[:, 1, 0] - r[:, 0, 1]) * (0.5 / s0)], -1)
    s1 = 0.5 / safe_sqrt(1 + r[:, 0, 0] - r[:, 1, 1] - r[:, 2, 2], m1)
    q_1 = torch.stack([(r[:, 2, 1] - r[:, 1, 2]) * s1, 0.5 * s1, (r[:, 1, 0] + r[:, 0, 1]) * s1, (r[:, 2, 0] + r[:, 0, 2]) * s1], -1)
    s2 = 0.5 / safe_sqrt(1 + r[:, 1, 1] - r[:, 0, 0] - r[:, 2, 2], m2)
    q_2 = torch.stack([(r[:, 0, 2] - r[:, 2, 0]) * s2, (r[:, 2, 1] + r[:, 1, 2]) * s2, 0.5 * s2, (r[:, 0, 1] + r[:, 1, 0]) * s2], -1)
    m3 = ~(m0 | m1 | m2)
    s3 = 0.5 / safe_sqrt(1 + r[:, 2, 2] - r[:, 0, 0] - r[:, 1, 1], m3)
    q_3 = torch.stack([(r[:, 1, 0] - r[:, 0, 1]) * s3, (r[:, 0, 2] + r[:, 2, 0]) * s3, (r[:, 1, 2] + r[:, 2, 1]) * s3, 0.5 * s3], -1)
    return torch.where(m0[:, None], q_0, torch.where(m1[:, None], q_1, torch.where(m2[:, None], q_2, q_3)))


def relations_to_matrix(Rr, Rs):
    """render/utils.py:128-134: one-hot edge matrices [1,E,N] -> dense adjacency [N,N] (receiver row, sender column); the
    reference's Python loop with two `.item()` per edge becomes one scatter.  Rows that are not one-hot are rejected like the
    reference's asserts."""
    if not bool(((Rr[0].sum(-1) == 1) & (Rs[0].sum(-1) == 1)).all()):
        raise AssertionError("Rr / Rs rows must be one-hot")
    rel = torch.zeros((Rr.shape[-1], Rs.shape[-1]), dtype=torch.int64, device=Rr.device)
    rel[Rr[0].argmax(-1), Rs[0].argmax(-1)] = 1
    return rel


def _csr(relations, n_bones, device):
    """(row_ptr int32 [n_bones+1], cols int32 [nnz]) of the bone graph; columns >= n_bones are dropped by the kernel."""
    if isinstance(relations, EdgeIndex):
        if relations.B != 1:
            raise ValueError("interpolate_motions takes the edge index of one graph (B = 1)")
        return relations.row_ptr[0], relations.senders[0]
    if isinstance(relations, (tuple, list)):
        rp, cols = relations
        return rp.to(device=device, dtype=torch.int32).contiguous(), cols.to(device=device, dtype=torch.int32).contiguous()
    rel = relations.to(device)
    if rel.shape[0] < n_bones or rel.shape[1] < n_bones:
        raise ValueError("relations must be at least [n_bones, n_bones]")
    rel = rel[:n_bones, :n_bones] != 0
    row_ptr = torch.zeros(n_bones + 1, dtype=torch.int32, device=device)
    row_ptr[1:] = rel.sum(1).cumsum(0).to(torch.int32)
    cols = rel.nonzero()[:, 1].to(torch.int32).contiguous()
    return row_ptr, cols


def bone_transforms(bones, motions, relations):
    """Per-bone records [n_bones, 20] (R | c | q | b, include/gsd.h) and rotations [n_bones,3,3]."""
    bones, motions = _f32(bones, "bones"), _f32(motions, "motions")
    n_bones = bones.shape[0]
    row_ptr, cols = _csr(relations, n_bones, bones.device)
    if cols.numel() == 0:
        cols = torch.zeros(1, dtype=torch.int32, device=bones.device)
    tf = torch.empty((n_bones, _BONE_TF), device=bones.device, dtype=torch.float32)
    rot = torch.empty((n_bones, 3, 3), device=bones.device, dtype=torch.float32)
    with torch.cuda.device(bones.device):
        _lib.check(_lib.lib().gsd_skin_bone_transforms(n_bones, bones.data_ptr(), motions.data_ptr(), row_ptr.data_ptr(), cols.data_ptr(),
                                                       tf.data_ptr(), rot.data_ptr(), _stream()), "gsd_skin_bone_transforms")
    return tf, rot


def interpolate_motions(bones, motions, relations, xyz, rot=None, quat=None, weights=None, device='cuda', return_weights=True):
    """render/utils.py:137-243.  bones, motions [n_bones,3]; relations: dense [n_bones,n_bones] adjacency (reference), an
    `EdgeIndex` of the rollout graph (columns >= n_bones, i.e. the tool node, are ignored like `relations[:nobj, :nobj]`), or a
    CSR pair (row_ptr, cols); xyz [n_particles,3]; quat [n_particles,4] (w,x,y,z) or None; weights [n_particles,n_bones] or None
    (inverse-distance weights are computed).  Returns (xyz_transformed, rot, weights) like the reference: `rot` is the blended
    quaternion when `quat` is given, else the `rot` argument unchanged.  `return_weights=False` skips materialising the dense
    [n_particles, n_bones] weight matrix (the rollout discards it, dynamics_module.py:150) and returns None in its place."""
    if not torch.device(device).type == 'cuda':
        raise ValueError("interpolate_motions runs on CUDA only (no CPU path)")
    xyz = _f32(xyz, "xyz")
    n_particles, n_bones = xyz.shape[0], bones.shape[0]
    tf, _ = bone_transforms(bones, motions, relations)
    q_in = _f32(quat, "quat") if quat is not None else None
    w_in = _f32(weights, "weights") if weights is not None else None
    if w_in is not None and tuple(w_in.shape) != (n_particles, n_bones):
        raise ValueError("weights must be [n_particles, n_bones]")
    xyz_out = torch.empty_like(xyz)
    q_out = torch.empty_like(q_in) if q_in is not None else None
    w_out = torch.empty((n_particles, n_bones), device=xyz.device, dtype=torch.float32) if (return_weights and w_in is None) else None
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.lib().gsd_skin_apply(n_particles, n_bones, xyz.data_ptr(), q_in.data_ptr() if q_in is not None else None,
                                             tf.data_ptr(), w_in.data_ptr() if w_in is not None else None, xyz_out.data_ptr(),
                                             q_out.data_ptr() if q_out is not None else None,
                                             w_out.data_ptr() if w_out is not None else None, _stream()), "gsd_skin_apply")
    if w_in is not None:
        w_ret = weights
    else:
        w_ret = w_out
    return xyz_out, (q_out if q_in is not None else rot), w_ret
