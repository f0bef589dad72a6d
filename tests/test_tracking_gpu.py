"""GPU parity tests of the tracking-iteration kernels (photometric, priors, Adam, get_loss) through the C ABI.

Tolerances: photometric loss 1e-5 abs, its gradient 1e-3 of max; prior losses 1e-4 rel, gradients 2e-3 of max (fp32 vs the
float64 autograd oracle; the sqrt(.+1e-20) terms amplify rounding near zero residuals); Adam 1e-6 rel vs torch.optim.Adam.
"""
import os

import numpy as np
import pytest
import torch

from oracle import tracking_oracle as T
from tests.helpers import make_camera, make_scene, rel_err

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "tracking_golden.npz"))


def _variables_from_case(c, dev="cuda"):
    from gs_dynamics_b200 import tracking as TR
    is_fg = c["is_fg"]
    v = {}
    v["neighbor_indices"] = c["neighbor_indices"].to(dev)
    v["neighbor_indices_i32"] = c["neighbor_indices"].to(torch.int32).to(dev).contiguous()
    v["neighbor_weight"] = c["neighbor_weight"].float().to(dev)
    v["neighbor_dist"] = c["neighbor_dist"].float().to(dev)
    v["prev_offset"] = c["prev_offset"].float().to(dev)
    v["prev_inv_rot_fg"] = c["prev_inv_rot_fg"].float().to(dev)
    v["in_ptr"], v["in_edge"] = TR.build_in_edges(v["neighbor_indices_i32"])
    v["fg_index"] = None if bool(is_fg.all()) else torch.nonzero(is_fg).reshape(-1).to(torch.int32).to(dev)
    v["bg_index"] = torch.nonzero(~is_fg).reshape(-1).to(torch.int32).to(dev)
    v["init_bg_pts"] = c["init_bg_pts"].float().to(dev)
    v["init_bg_rot"] = c["init_bg_rot"].float().to(dev)
    return v


def test_photometric_golden_fixture():
    from gs_dynamics_b200 import tracking as TR
    x = torch.tensor(GOLD["ph_x"]).cuda().requires_grad_(True)
    y = torch.tensor(GOLD["ph_y"]).cuda()
    loss, parts = TR.photometric_loss(x, y, return_parts=True)
    loss.backward()
    assert abs(loss.item() - float(GOLD["ph_loss"])) < 1e-5
    assert abs(parts[1].item() - float(GOLD["ph_l1"])) < 1e-5 and abs(parts[2].item() - float(GOLD["ph_ssim"])) < 1e-5
    assert rel_err(x.grad.cpu(), GOLD["ph_grad"]) < 1e-3


@pytest.mark.parametrize("C,H,W", [(3, 480, 640), (3, 33, 47), (1, 16, 16), (6, 100, 70)])
def test_photometric_vs_oracle(C, H, W):
    from gs_dynamics_b200 import tracking as TR
    g = torch.Generator().manual_seed(C * H)
    x = torch.rand(C, H, W, generator=g)
    y = (x + 0.1 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    xd = x.double().requires_grad_(True)
    lo = T.photometric(xd, y.double())
    lo.backward()
    xc = x.cuda().requires_grad_(True)
    loss = TR.photometric_loss(xc, y.cuda())
    (3.0 * loss).backward()  # non-unit upstream gradient
    assert abs(loss.item() - lo.item()) < 1e-5
    assert rel_err(xc.grad.cpu(), 3.0 * xd.grad) < 1e-3


@pytest.mark.parametrize("tag", ["a", "b"])
def test_priors_golden_fixture(tag):
    from gs_dynamics_b200 import tracking as TR
    G, K, seed, fb = GOLD[f"pr_{tag}_cfg"]
    c = T.make_prior_case(int(G), int(K), int(seed), float(fb))
    v = _variables_from_case(c)
    x = c["means3D"].cuda().requires_grad_(True)
    q = c["rotations"].cuda().requires_grad_(True)
    total, parts = TR.track_prior_losses(x, q, v, 200.0, 4.0, 1000.0, 200.0)
    total.backward()
    for i, k in enumerate(("rigid", "rot", "iso", "floor", "bg")):
        ref = float(GOLD[f"pr_{tag}_{k}"])
        assert abs(parts[i].item() - ref) <= 1e-4 * max(1e-3, abs(ref)), k
    assert abs(total.item() - float(GOLD[f"pr_{tag}_total"])) <= 1e-4 * abs(float(GOLD[f"pr_{tag}_total"]))
    assert rel_err(x.grad.cpu(), GOLD[f"pr_{tag}_gx"]) < 2e-3
    assert rel_err(q.grad.cpu(), GOLD[f"pr_{tag}_gq"]) < 2e-3


@pytest.mark.parametrize("G,K,fb", [(5000, 20, 0.0), (3000, 20, 0.3), (257, 3, 0.5)])
def test_priors_vs_float64_oracle(G, K, fb):
    from gs_dynamics_b200 import tracking as TR
    c = T.make_prior_case(G, K, 7, fb, dtype=torch.float64)
    x64 = c["means3D"].clone().requires_grad_(True)
    q64 = c["rotations"].clone().requires_grad_(True)
    L = T.prior_losses(x64, q64, c["is_fg"], c["prev_inv_rot_fg"], c["neighbor_indices"], c["neighbor_weight"],
                       c["neighbor_dist"], c["prev_offset"], c["init_bg_pts"], c["init_bg_rot"])
    w = dict(rigid=200.0, rot=4.0, iso=1000.0, floor=2.0, bg=200.0)
    tot = sum(w[k] * L[k] for k in w)
    tot.backward()
    v = _variables_from_case(c)
    x = c["means3D"].float().cuda().requires_grad_(True)
    q = c["rotations"].float().cuda().requires_grad_(True)
    total, parts = TR.track_prior_losses(x, q, v, 200.0, 4.0, 1000.0, 200.0)
    (0.5 * total).backward()
    for i, k in enumerate(("rigid", "rot", "iso", "floor", "bg")):
        assert abs(parts[i].item() - L[k].item()) <= 2e-4 * max(1e-3, abs(L[k].item())), k
    assert rel_err(x.grad.cpu(), 0.5 * x64.grad) < 2e-3
    assert rel_err(q.grad.cpu(), 0.5 * q64.grad) < 2e-3


def test_priors_packed_edge_records_match_unpacked_tables():
    from gs_dynamics_b200 import tracking as TR
    c = T.make_prior_case(4000, 20, 11, 0.2)
    v = _variables_from_case(c)
    x = c["means3D"].cuda().requires_grad_(True)
    q = c["rotations"].cuda().requires_grad_(True)
    t0, p0 = TR.track_prior_losses(x, q, v, 200.0, 4.0, 1000.0, 200.0)
    t0.backward()
    gx0, gq0 = x.grad.clone(), q.grad.clone()
    x.grad = None; q.grad = None
    # packed tables in the caller's order, then re-numbered along the Morton curve of the means (what the episode loop uses)
    for positions in (None, x.detach()):
        v.pop("priors_layout", None)
        TR.pack_edge_records(v, positions=positions)
        lay = v["edge_records"]["layout"]
        ident = bool((lay["perm"] == torch.arange(lay["perm"].numel(), device="cuda")).all())
        assert ident == (positions is None)
        x.grad = None; q.grad = None
        t1, p1 = TR.track_prior_losses(x, q, v, 200.0, 4.0, 1000.0, 200.0)
        t1.backward()
        assert float((p1 - p0).abs().max()) <= 2e-6 * float(p0.abs().max())
        assert rel_err(x.grad.cpu(), gx0.cpu()) < 1e-5 and rel_err(q.grad.cpu(), gq0.cpu()) < 1e-5
    # a stale pack (prev_offset modified afterwards) must not be used
    v["prev_offset"].mul_(1.0)
    assert v["edge_records"]["ver"] != v["prev_offset"]._version


def test_fused_adam_matches_torch_adam():
    from gs_dynamics_b200 import tracking as TR
    g = torch.Generator().manual_seed(0)
    shapes = [(1000, 3), (1000, 4), (1000, 1), (50, 3), (7,)]
    lrs = [1.6e-4, 1e-3, 0.05, 1e-4, 0.0]
    p_ref = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    p_gpu = [torch.nn.Parameter(p.detach().clone().cuda()) for p in p_ref]
    opt_ref = torch.optim.Adam([{'params': [p], 'lr': lr} for p, lr in zip(p_ref, lrs)], lr=0.0, eps=1e-15)
    opt = TR.FusedAdam([{'params': [p], 'lr': lr, 'name': str(i)} for i, (p, lr) in enumerate(zip(p_gpu, lrs))], lr=0.0, eps=1e-15)
    for it in range(5):
        for pr, pg in zip(p_ref, p_gpu):
            gr = torch.randn(pr.shape, generator=g) * (10.0 ** (it - 2))
            pr.grad = gr.clone()
            pg.grad = gr.cuda()
        opt_ref.step()
        opt.step()
    for pr, pg in zip(p_ref, p_gpu):
        assert rel_err(pg.detach().cpu(), pr.detach()) < 1e-6
    assert float(opt.state[p_gpu[0]]['step'].item()) == 5.0


def _tracking_problem(G, cam_ids=(0, 1), w=160, h=120, boost=1.0, box=0.6, seed=0):
    """Steady-state (t > 0) tracking set-up on synthetic data: scene S(G, seed) to fit, targets rendered from a slightly
    moved copy, kNN tables from the initial points (scaled down)."""
    from gs_dynamics_b200 import tracking as TR, scenes
    from gs_dynamics_b200 import rasterizer as R
    W0, H0, cams = scenes.demo_cameras()
    sc, act = make_scene(G, seed, scale_boost=boost, box_scale=box)
    params = {k: torch.nn.Parameter(v.cuda().contiguous()) for k, v in sc.items()}
    params['cam_m'] = torch.nn.Parameter(torch.zeros(50, 3, device="cuda"))
    params['cam_c'] = torch.nn.Parameter(torch.zeros(50, 3, device="cuda"))
    params['rgb_colors'].requires_grad = False
    variables = {'max_2D_radius': torch.zeros(G, device="cuda"), 'scene_radius': 1.0,
                 'means2D_gradient_accum': torch.zeros(G, device="cuda"), 'denom': torch.zeros(G, device="cuda")}
    opt = TR.initialize_optimizer(params, variables)
    variables = TR.initialize_post_first_timestep(params, variables, opt, num_knn=8)
    dataset = []
    tgt = {k: v.clone() for k, v in act.items()}
    tgt['means3D'] = tgt['means3D'] + torch.tensor([0.002, -0.001, 0.001])
    for cid in cam_ids:
        k, w2c = cams[cid]
        k = k.copy(); k[0] *= w / W0; k[1] *= h / H0
        cam = TR.setup_camera(w, h, k, w2c, near=1.0, far=100)
        with torch.no_grad():
            tc = {kk: vv.cuda() for kk, vv in tgt.items()}
            im, _, _, _ = R.raster_forward(cam, tc['means3D'], tc['opacities'], tc['colors_precomp'], tc['scales'], tc['rotations'])
            sg, _, _, _ = R.raster_forward(cam, tc['means3D'], tc['opacities'], sc['seg_colors'].cuda(), tc['scales'], tc['rotations'])
        dataset.append({'cam': cam, 'im': im.clone(), 'seg': sg.clone(), 'id': cid})
    params, variables = TR.initialize_per_timestep(params, variables, opt)
    # move away from the previous state: at exactly zero residual the priors' sqrt(r^2 + 1e-20) terms are non-smooth and
    # their gradient is rounding noise (in the reference as well), which makes trajectory comparisons meaningless
    g = torch.Generator().manual_seed(seed + 7)
    with torch.no_grad():
        params['means3D'].add_((1e-3 * torch.randn(G, 3, generator=g)).cuda())
        params['unnorm_rotations'].add_((1e-2 * torch.randn(G, 4, generator=g)).cuda())
    return params, variables, opt, dataset


def test_get_loss_fused_equals_two_pass_and_oracle_composition():
    from gs_dynamics_b200 import tracking as TR
    params, variables, opt, dataset = _tracking_problem(3000)
    data = dataset[0]
    opt.zero_grad()
    l_f, variables = TR.get_loss(params, data, variables, False, fused=True)
    l_f.backward()
    g_f = {k: p.grad.clone() for k, p in params.items() if p.grad is not None}
    opt.zero_grad()
    l_2, variables = TR.get_loss(params, data, variables, False, fused=False)
    l_2.backward()
    assert abs(l_f.item() - l_2.item()) <= 1e-5 * abs(l_2.item())
    for k in g_f:
        assert rel_err(g_f[k].cpu(), params[k].grad.cpu()) < 1e-4, k
    # loss value against the oracle composition on the same rendered images
    rv = TR.params2rendervar(params)
    with torch.no_grad():
        from gs_dynamics_b200 import rasterizer as R
        im, radius, _, _ = R.raster_forward(data['cam'], rv['means3D'], rv['opacities'], rv['colors_precomp'], rv['scales'], rv['rotations'])
        sg, _, _, _ = R.raster_forward(data['cam'], rv['means3D'], rv['opacities'], params['seg_colors'].detach(), rv['scales'], rv['rotations'])
    ref = 50.0 * T.photometric(im.cpu().double(), data['im'].cpu().double()) + 200.0 * T.photometric(sg.cpu().double(), data['seg'].cpu().double())
    is_fg = (params['seg_colors'][:, 0] > 0.5).cpu()
    L = T.prior_losses(rv['means3D'].detach().cpu().double(), rv['rotations'].detach().cpu().double(), is_fg,
                       variables['prev_inv_rot_fg'].cpu().double(), variables['neighbor_indices'].cpu(),
                       variables['neighbor_weight'].cpu().double(), variables['neighbor_dist'].cpu().double(),
                       variables['prev_offset'].cpu().double(), variables['init_bg_pts'].cpu().double(), variables['init_bg_rot'].cpu().double())
    ref = ref + 200.0 * L['rigid'] + 4.0 * L['rot'] + 1000.0 * L['iso'] + 2.0 * L['floor'] + 200.0 * L['bg']
    assert abs(l_2.item() - ref.item()) <= 2e-4 * abs(ref.item())
    assert variables['seen'].dtype == torch.bool and bool((variables['max_2D_radius'] >= radius.float()).all())


def test_tracking_step_graph_matches_eager_and_reduces_loss():
    from gs_dynamics_b200 import tracking as TR
    pa, va, oa, da = _tracking_problem(2000)
    pb, vb, ob, db = _tracking_problem(2000)
    eager = TR.TrackingStep(pa, va, oa, da, use_graph=False)
    graph = TR.TrackingStep(pb, vb, ob, db, use_graph=True)
    eager.prepare(); graph.prepare()   # prepare() undoes its warm-up iterations: both paths start from the same state
    seq = [0, 1, 1, 0, 1, 0, 0, 1]
    le = [float(eager.step(c)) for c in seq]
    lg = [float(graph.step(c)) for c in seq]
    np.testing.assert_allclose(lg, le, rtol=2e-4)
    assert rel_err(pb['means3D'].detach().cpu(), pa['means3D'].detach().cpu()) < 1e-5
    first = float(eager.step(0))
    for _ in range(60):
        eager.step(0)
    assert float(eager.step(0)) < first


def test_photometric_two_sets_with_affine_matches_composition():
    """One launch for both renders + fused cam_m/cam_c colour correction == the reference composition (train_utils.py:182-195)."""
    import ctypes as C
    from gs_dynamics_b200 import tracking as TR, _lib
    g = torch.Generator().manual_seed(5)
    H, W = 70, 100
    x = torch.rand(6, H, W, generator=g)
    y = (x + 0.1 * torch.randn(6, H, W, generator=g)).clamp(0, 1)
    m, c = 0.1 * torch.randn(3, generator=g), 0.05 * torch.randn(3, generator=g)
    xd = x.double().requires_grad_(True)
    im = torch.exp(m.double())[:, None, None] * xd[:3] + c.double()[:, None, None]
    ref = 50.0 * T.photometric(im, y[:3].double()) + 200.0 * T.photometric(xd[3:], y[3:].double())
    ref.backward()
    xc, yc = x.cuda().contiguous(), y.cuda().contiguous()
    ws = TR._ph_workspace(xc)
    out = torch.empty(7, device="cuda")
    d = TR._ph_desc(xc, yc, 2, 0.8, 0.2, (50.0, 200.0), ws, affine=(m.cuda(), c.cuda()))
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().gsd_photometric_forward(C.byref(d), out.data_ptr(), st), "fwd")
    grad = torch.empty_like(xc)
    _lib.check(_lib.lib().gsd_photometric_backward(C.byref(d), None, grad.data_ptr(), st), "bwd")
    assert abs(out[6].item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert rel_err(grad.cpu(), xd.grad) < 1e-3


def test_fused_tracking_step_matches_autograd_path():
    """FusedTrackingStep (no PyTorch ops in the loop) follows the same loss / parameter trajectory as get_loss + backward +
    FusedAdam in the steady state (lr = 0 for every group except means3D / unnorm_rotations)."""
    from gs_dynamics_b200 import tracking as TR
    def freeze(opt):
        for g in opt.param_groups:
            if g['name'] not in ('means3D', 'unnorm_rotations'):
                g['lr'] = 0.0
    pa, va, oa, da = _tracking_problem(2500)
    pb, vb, ob, db = _tracking_problem(2500)
    freeze(oa); freeze(ob)
    eager = TR.TrackingStep(pa, va, oa, da, use_graph=False)
    fused = TR.FusedTrackingStep(pb, vb, ob, db, use_graph=False)
    eager.prepare(); fused.prepare()
    seq = [0, 1, 1, 0, 0, 1, 0, 1, 1, 0]
    x0 = pa['means3D'].detach().clone()
    le = [float(eager.step(c)) for c in seq]
    lf = [float(fused.step(c)) for c in seq]
    np.testing.assert_allclose(lf, le, rtol=5e-4)
    # parameters moved by ~10 lr-sized Adam steps; the two paths must agree on that displacement
    moved = (pa['means3D'].detach() - x0).abs().max().item()
    assert float((pb['means3D'].detach() - pa['means3D'].detach()).abs().max()) < 0.05 * moved
    assert rel_err(pb['unnorm_rotations'].detach().cpu(), pa['unnorm_rotations'].detach().cpu()) < 1e-3
    # and under CUDA-graph replay
    pc, vc, oc, dc = _tracking_problem(2500)
    freeze(oc)
    fg = TR.FusedTrackingStep(pc, vc, oc, dc, use_graph=True)
    fg.prepare()
    pd, vd, od, dd = _tracking_problem(2500)
    freeze(od)
    fe = TR.FusedTrackingStep(pd, vd, od, dd, use_graph=False)
    fe.prepare()
    lg = [float(fg.step(c)) for c in seq]
    l2 = [float(fe.step(c)) for c in seq]
    np.testing.assert_allclose(lg, l2, rtol=1e-4)


def test_fused_backward_update_kernel_equals_backward_plus_update():
    """gsd_track_backward_update (per-Gaussian rasterizer backward with the Adam update applied in registers) against the two
    launches it replaces (gsd_raster_backward + gsd_track_update): same parameters, Adam moments, step counters, radii bookkeeping."""
    from gs_dynamics_b200 import tracking as TR
    runs = []
    for fuse in (True, False):
        p, v, o, d = _tracking_problem(3000)
        for g in o.param_groups:
            if g['name'] not in ('means3D', 'unnorm_rotations'):
                g['lr'] = 0.0
        st = TR.FusedTrackingStep(p, v, o, d, use_graph=False)
        st.fuse_update = fuse
        st.prepare()
        losses = [float(st.step(c)) for c in (0, 1, 1, 0, 1)]
        runs.append((p, o, v, losses))
    (pa, oa, va, la), (pb, ob, vb, lb) = runs
    # same arithmetic, compiled in two kernels (FMA contraction may differ by an ulp per step): 1e-5 after five iterations
    np.testing.assert_allclose(la, lb, rtol=1e-5)
    for k in ('means3D', 'unnorm_rotations'):
        a, b = pa[k].detach(), pb[k].detach()
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()), k
        sa, sb = oa.state[pa[k]], ob.state[pb[k]]
        assert float(sa['step']) == float(sb['step']) == 5.0
        assert rel_err(sa['exp_avg'].cpu(), sb['exp_avg'].cpu()) < 1e-5 and rel_err(sa['exp_avg_sq'].cpu(), sb['exp_avg_sq'].cpu()) < 1e-5
    assert torch.equal(va['max_2D_radius'], vb['max_2D_radius']) and torch.equal(va['seen'], vb['seen'])


def test_set_target_u8_builds_the_reference_planes():
    """set_target_u8 (bytes up, float planes built on the device) == set_target with the planes train_utils.py:66-75 builds on the host."""
    from gs_dynamics_b200 import tracking as TR
    p, v, o, d = _tracking_problem(2000)
    for g in o.param_groups:
        if g['name'] not in ('means3D', 'unnorm_rotations'):
            g['lr'] = 0.0
    st = TR.FusedTrackingStep(p, v, o, d, use_graph=False)
    st.prepare()
    H, W = d[0]['im'].shape[1:]
    gen = torch.Generator().manual_seed(3)
    for cin in (3, 4):
        im_u8 = torch.randint(0, 256, (H, W, cin), dtype=torch.uint8, generator=gen)
        seg_u8 = torch.randint(0, 2, (H, W), dtype=torch.uint8, generator=gen)
        im = (torch.tensor(im_u8.numpy()).float().permute(2, 0, 1)[:3].contiguous() / 255)          # the reference's host code
        seg = torch.tensor(seg_u8.numpy().astype(np.float32)).float()
        seg_col = torch.stack((seg, torch.zeros_like(seg), 1 - seg))
        st.set_target(0, im.cuda(), seg_col.cuda())
        ref_t, ref_mu, ref_s = st.targets[0].clone(), st.tstats[0][0].clone(), st.tstats[0][1].clone()
        st.targets[0].zero_()
        st.set_target_u8(0, im_u8.pin_memory(), seg_u8.pin_memory())
        torch.cuda.synchronize()
        assert torch.equal(st.targets[0], ref_t)
        assert torch.equal(st.tstats[0][0], ref_mu) and torch.equal(st.tstats[0][1], ref_s)
    with pytest.raises(ValueError):
        st.set_target_u8(0, torch.zeros((H, W, 2), dtype=torch.uint8), torch.zeros((H, W), dtype=torch.uint8))


def test_t0_densification_surgery_and_short_episode():
    """A9/A12 + episode loop: densify bookkeeping (clone / split / prune, Adam state surgery) and a 2-frame episode."""
    from gs_dynamics_b200 import tracking as TR, scenes, rasterizer as R
    W0, H0, cams = scenes.demo_cameras()
    rng = np.random.default_rng(0)
    n = 1200
    pts = np.concatenate([rng.uniform([-0.1, -0.1, -0.05], [0.1, 0.1, 0.0], (n, 3)) + [0.28, 0.073, 0.0], rng.uniform(size=(n, 3)),
                          np.ones((n, 1))], 1)
    cam_centers = np.stack([np.linalg.inv(c[1])[:3, 3] for c in cams])
    params, variables = TR.initialize_params_from_point_cloud(pts, cam_centers)
    opt = TR.initialize_optimizer(params, variables)
    # targets: the same cloud rendered with slightly different colours
    datasets = []
    for t in range(2):
        ds = []
        for cid in (0, 1):
            k, w2c = cams[cid]
            k = k.copy(); k[0] *= 160 / W0; k[1] *= 120 / H0
            cam = TR.setup_camera(160, 120, k, w2c, near=1.0, far=100)
            with torch.no_grad():
                rv = TR.params2rendervar(params)
                shift = torch.tensor([0.002 * t, 0.0, 0.0], device="cuda")
                im, _, _, _ = R.raster_forward(cam, rv['means3D'] + shift, rv['opacities'], (rv['colors_precomp'] * 0.9).contiguous(),
                                               rv['scales'] * 1.2, rv['rotations'])
                sg, _, _, _ = R.raster_forward(cam, rv['means3D'] + shift, rv['opacities'], params['seg_colors'].detach(),
                                               rv['scales'] * 1.2, rv['rotations'])
            ds.append({'cam': cam, 'im': im.clone(), 'seg': sg.clone(), 'id': cid})
        datasets.append(ds)
    # a few t = 0 iterations, then force one densification round
    for i in range(6):
        loss, variables = TR.get_loss(params, datasets[0][i % 2], variables, True)
        loss.backward()
        with torch.no_grad():
            params, variables, npts = TR.densify(params, variables, opt, i, 0.005, 0.25, 0.05)
            opt.step(); opt.zero_grad(set_to_none=True)
    assert npts == n and float(variables['denom'].max()) == 6.0
    loss, variables = TR.get_loss(params, datasets[0][0], variables, True)
    loss.backward()
    g_before = (variables['means2D_gradient_accum'] + torch.norm(variables['means2D'].grad[:, :2], dim=-1) * variables['seen']) / \
               (variables['denom'] + variables['seen'])
    limit = 0.05 * variables['scene_radius']
    big = torch.exp(params['log_scales']).max(1).values > limit
    n_clone = int(((g_before >= 1e-6) & ~big).sum()); n_split = int(((g_before >= 1e-6) & big).sum())
    n_transparent = 0  # opacities start at sigmoid(0) = 0.5 > thresholds
    with torch.no_grad():
        params, variables, npts = TR.densify(params, variables, opt, 500, 0.005, 0.25, 0.05, grad_thresh=1e-6)
        opt.step(); opt.zero_grad(set_to_none=True)
    assert npts == n + n_clone + n_split - n_transparent and n_clone + n_split > 0
    for k in TR.PER_POINT_KEYS:
        p = params[k]
        st = opt.state[p]
        assert p.shape[0] == npts and st['exp_avg'].shape == p.shape and st['exp_avg_sq'].shape == p.shape
    assert variables['denom'].shape[0] == npts and float(variables['denom'].abs().max()) == 0.0
    # the episode loop: a short t = 0 and one tracked frame through the fused graph path
    l0 = float(TR.get_loss(params, datasets[0][0], variables, True)[0])
    params, variables, snaps = TR.train_frames(params, variables, opt, datasets, iters_first=25, iters_next=40, num_knn=6)
    assert len(snaps) == 2 and snaps[1]['means3D'].shape == (npts, 3)
    l1 = float(TR.get_loss(params, datasets[0][0], variables, True)[0])
    assert np.isfinite(l1) and l1 < l0
    assert "edge_records" in variables and variables['prior_losses'].shape == (6,)


@pytest.mark.parametrize("n,k", [(50000, 20), (3000, 3), (65, 64), (2, 1)])
def test_device_knn_matches_kdtree(n, k):
    """gsd_knn (exact, float64) vs scipy's KD-tree in float64 on the same float32 points: squared distances to 1e-12
    relative, indices identical wherever the distance is not tied with the next one."""
    from scipy.spatial import cKDTree
    from gs_dynamics_b200 import tracking as TR
    rng = np.random.default_rng(n)
    pts = rng.uniform([-0.25, -0.25, -0.1], [0.25, 0.25, 0.0], size=(n, 3)).astype(np.float32)
    sq, idx = TR.knn(pts, k)
    d, ref = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=k + 1)
    ref_sq, ref_idx = d[:, 1:].reshape(n, k) ** 2, ref[:, 1:].reshape(n, k)
    assert sq.dtype == np.float64 and idx.dtype == np.int64 and sq.shape == (n, k)
    assert np.allclose(sq, ref_sq, rtol=1e-12, atol=1e-18)
    untied = np.ones((n, k), bool)
    if k > 1:
        gap = np.diff(ref_sq, axis=1) > 1e-15
        untied[:, 1:] &= gap
        untied[:, :-1] &= gap
    assert np.array_equal(idx[untied], ref_idx[untied])
    assert not np.any(idx == np.arange(n)[:, None])          # self excluded
    t_sq, t_idx = TR.knn(torch.tensor(pts).cuda(), k)         # tensor in -> tensors out
    assert t_sq.is_cuda and np.array_equal(t_idx.cpu().numpy(), idx)
