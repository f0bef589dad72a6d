"""CPU: host-side logic of the product (no kernels): workloads, camera conventions, graph bookkeeping, sharding."""
import numpy as np
import torch

from gs_dynamics_b200 import dist as gdist
from gs_dynamics_b200 import gnn, scenes, workloads


def test_camera_matrices_follow_reference_conventions():
    w, h, cams = scenes.demo_cameras()
    assert (w, h) == (640, 480) and len(cams) == 4
    k, w2c = cams[0]
    m = scenes.camera_matrices(w, h, k, w2c, near=1.0, far=100.0)
    assert m["viewmatrix"].shape == (1, 4, 4) and m["projmatrix"].shape == (1, 4, 4)
    np.testing.assert_allclose(m["viewmatrix"][0].numpy(), np.asarray(w2c, np.float32).T, atol=1e-6)   # viewmatrix = w2c^T
    assert abs(m["tanfovx"] - w / (2 * k[0][0])) < 1e-9
    # campos = camera centre in world coordinates
    np.testing.assert_allclose(m["campos"].numpy(), np.linalg.inv(w2c)[:3, 3], atol=1e-5)
    # a world point in front of the camera projects inside NDC [-1, 1]
    p = np.linalg.inv(w2c) @ np.array([0, 0, 0.7, 1.0])
    hom = torch.tensor(p, dtype=torch.float32) @ m["projmatrix"][0]
    assert abs(hom[0] / hom[3]) < 1 and abs(hom[1] / hom[3]) < 1


def test_tracking_workload_is_seeded_and_consistent():
    a = workloads.tracking_problem(500, seed=3, num_knn=5)
    b = workloads.tracking_problem(500, seed=3, num_knn=5)
    assert torch.equal(a["params"]["means3D"], b["params"]["means3D"])
    v = a["variables"]
    assert v["neighbor_indices"].shape == (500, 5) and v["prev_offset"].shape == (500, 5, 3)
    assert int((v["neighbor_indices"] == torch.arange(500)[:, None]).sum()) == 0   # self excluded (o3d_knn drops index 0)
    np.testing.assert_allclose(v["neighbor_weight"].numpy(), np.exp(-2000 * v["neighbor_dist"].numpy() ** 2), rtol=1e-4)
    m = torch.rand(4, 6)
    s = workloads.seg_target_from_mask(m)
    assert s.shape == (3, 4, 6) and torch.equal(s[2], 1 - m) and float(s[1].abs().max()) == 0


def test_edge_index_from_dense_handles_padding_and_order():
    N = 6
    recv = torch.tensor([0, 0, 2, 5, 5, 5])
    send = torch.tensor([1, 2, 0, 0, 1, 3])
    Rr = torch.zeros(1, 9, N); Rs = torch.zeros(1, 9, N)
    perm = torch.tensor([3, 0, 5, 1, 4, 2])                 # unsorted input rows + 3 zero padding rows
    for slot, e in enumerate(perm):
        Rr[0, slot, recv[e]] = 1; Rs[0, slot, send[e]] = 1
    e = gnn.edge_index_from_dense(Rr, Rs)
    assert e.capacity == 9 and int(e.n_edges[0]) == 6
    assert e.row_ptr[0].tolist() == [0, 2, 2, 3, 3, 3, 6]
    got = sorted(zip(e.receivers[0, :6].tolist(), e.senders[0, :6].tolist()))
    assert got == sorted(zip(recv.tolist(), send.tolist()))
    assert e.receivers[0, 6:].tolist() == [-1, -1, -1]
    assert gnn.edge_capacity(2001, 1, 8) == 2000 * 8 + 2 * 2000


def test_shard_partitions_units_evenly():
    for n, w in ((8, 8), (8, 3), (5, 2), (3, 4)):
        parts = [gdist.shard(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_gnn_training_losses_match_reference_formulation_cpu():
    """gnn_train.length_loss / local_rigid_loss (index gathers on an EdgeIndex built from the dataset's dense one-hots, pure torch:
    runs on the CPU) vs the one-hot bmm formulation of train.py:66-102 restated in the oracle — values and gradients, including
    zero padding rows and a tool node outside the first n_p columns."""
    import torch
    from gs_dynamics_b200 import gnn_train
    from oracle import gnn_oracle as GO
    batch = GO.make_training_batch(3, 30, 5, "sloth", 2)
    Rr, Rs = GO.batch_edges(batch, 0.075, 4, True)
    Rr = torch.cat([Rr, torch.zeros(3, 6, Rr.shape[2])], 1)      # padded rows as the dataset pads to max_nR
    Rs = torch.cat([Rs, torch.zeros(3, 6, Rs.shape[2])], 1)
    n_p = batch["p_instance"].shape[1]
    g = torch.Generator().manual_seed(0)
    pred0 = batch["state"][:, -1, :n_p] + 0.01 * torch.randn(3, n_p, 3, generator=g)
    for ours, ref in ((gnn_train.length_loss, GO.length_loss), (gnn_train.local_rigid_loss, GO.local_rigid_loss)):
        pa = pred0.clone().requires_grad_(True)
        pb = pred0.clone().requires_grad_(True)
        la = ours(pa, None, state=batch["state"], Rr=Rr, Rs=Rs)
        lb = ref(pb, batch["state"], Rr, Rs)
        la.backward()
        lb.backward()
        assert abs(la.item() - lb.item()) <= 1e-6 * max(1.0, abs(lb.item())) + 1e-12
        assert torch.allclose(pa.grad, pb.grad, rtol=1e-4, atol=1e-9)
        assert torch.isfinite(pa.grad).all()


def test_rigid_loss_and_umeyama_match_reference_fixture():
    """gnn_train.rigid_loss / umeyama_algorithm vs outputs of the reference's train.rigid_loss and gnn.utils.umeyama_algorithm
    (tests/golden/gnn_rigid_golden.npz, tools/make_golden.py)."""
    import os
    import numpy as np
    import torch
    from gs_dynamics_b200 import gnn_train
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "gnn_rigid_golden.npz"))
    X, mask = torch.tensor(G["rig_X"]), torch.tensor(G["rig_mask"])
    Y = torch.tensor(G["rig_Y"]).requires_grad_(True)
    loss = gnn_train.rigid_loss(Y, None, state=X[:, None].repeat(1, 3, 1, 1), obj_mask=mask)
    loss.backward()
    assert abs(loss.item() - float(G["rig_loss"])) <= 1e-6 * max(1.0, float(G["rig_loss"])) + 1e-10
    # the gradient is 2 (pred - target) / n with pred - target ~ 1e-2 of the coordinates: fp32 cancellation -> 1e-4 of max
    assert np.abs(Y.grad.numpy() - G["rig_grad"]).max() <= 1e-4 * np.abs(G["rig_grad"]).max() + 1e-12
    c, R, t = gnn_train.umeyama_algorithm(X, Y.detach(), mask.float(), fixed_scale=False)
    np.testing.assert_allclose(c.numpy(), G["ume_c"], rtol=1e-5)
    np.testing.assert_allclose(R.numpy(), G["ume_R"], atol=1e-5)
    np.testing.assert_allclose(t.numpy(), G["ume_t"], atol=1e-5)
    assert gnn_train.default_loss_funcs({"rigid_loss": True})[2][0] is gnn_train.rigid_loss


def test_first_frame_step_refuses_graph_capture():
    """TrackingStep(is_initial_timestep=True, use_graph=True) must fail loudly in Python, not as a CUDA capture error."""
    import pytest
    from gs_dynamics_b200 import tracking as TR
    with pytest.raises(ValueError, match="use_graph=False"):
        TR.TrackingStep({}, {}, None, [], is_initial_timestep=True, use_graph=True)
    TR.TrackingStep({}, {}, None, [], is_initial_timestep=True, use_graph=False)


def test_morton_order_is_a_locality_preserving_permutation():
    """Host side of the priors' packed tables (tracking._morton_order): a permutation whose consecutive points are close."""
    from gs_dynamics_b200 import tracking as TR
    g = torch.Generator().manual_seed(0)
    pts = torch.rand(20000, 3, generator=g) * torch.tensor([0.5, 0.5, 0.1])
    perm = TR._morton_order(pts)
    assert sorted(perm.tolist()) == list(range(20000))
    d_curve = (pts[perm][1:] - pts[perm][:-1]).norm(dim=1).mean()
    d_rand = (pts[1:] - pts[:-1]).norm(dim=1).mean()
    assert d_curve < 0.2 * d_rand
    # degenerate clouds (all points equal / a single point) must not divide by zero
    assert TR._morton_order(torch.zeros(5, 3)).tolist() == [0, 1, 2, 3, 4]
    assert TR._morton_order(torch.ones(1, 3)).tolist() == [0]


def test_edge_id_division_by_multiplication_is_exact():
    """track_losses.cu replaces e / K by (e * magic) >> shift with magic = ceil(2^shift / K), shift = 31 + ceil(log2 K): exact for
    every 31-bit edge id (same arithmetic restated with Python integers)."""
    rng = np.random.default_rng(0)
    for K in (1, 2, 3, 5, 7, 16, 17, 20, 31, 32, 33, 100, 1000, 65535):
        lg = 0
        while (1 << lg) < K:
            lg += 1
        shift = 31 + lg
        magic = ((1 << shift) + K - 1) // K
        assert magic < (1 << 32)
        es = np.concatenate([rng.integers(0, 2 ** 31, size=2000), np.arange(0, 4 * K), 2 ** 31 - 1 - np.arange(0, 4 * K),
                             np.arange(1, 50) * K, np.arange(1, 50) * K - 1])
        for e in es.tolist():
            assert (e * magic) >> shift == e // K, (K, e)
