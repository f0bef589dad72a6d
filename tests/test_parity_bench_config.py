"""Parity at the BENCHMARKED configuration (BASELINE.json configs[1] / [2]: 50k and 100k Gaussians, 4 cams @640x480), directly
against the oracle — not through the repo's own other instantiations:

  * the kernel instantiation bench.py times, gsd_blend_bwd_chunk_kernel<6, geometry-only, half-tile> (two colour sets in one
    pass, colours / opacities frozen), and the full 6-channel backward, vs oracle/raster_oracle.c with CH = 6;
  * ONE FusedTrackingStep iteration (the thing bench.py replays) vs ONE oracle.tracking_cpu.iteration — the reference's
    get_loss + backward + Adam composition (train_utils.py:167-246, train_gs.py:25-39) — from the same
    workloads.tracking_problem(G) state and the same target images: loss, the gradients (read from Adam's first moment:
    exp_avg = 0.1 g after the first step), the parameter step, `seen`, `max_2D_radius`.

Tolerances (fp32, stated here as the north-star asks): gradients max-abs error <= 1e-3 of the tensor max (rel_err) AND the
per-element metric |d| / (|ref| + 1e-3 max|ref|) <= 2e-2 at the 99.9th percentile (pct_rel_err; printed with -s);
loss 2e-4 relative; radii-derived statistics exact.
"""
import numpy as np
import pytest
import torch

from tests.helpers import make_camera, make_scene, oracle_backward, pct_rel_err, rel_err, settings_from

pytestmark = pytest.mark.gpu

TOL_MAX = 1e-3     # max-abs error / max-abs value
TOL_P999 = 2e-2    # 99.9th percentile of the per-element relative error


def _check(name, got, ref, tol_max=TOL_MAX, tol_p=TOL_P999):
    e_max = rel_err(got, ref)
    p999, worst = pct_rel_err(got, ref)
    print("%-28s rel_err %.3g  p99.9 %.3g  worst %.3g" % (name, e_max, p999, worst))
    assert e_max < tol_max, (name, e_max)
    assert p999 < tol_p, (name, p999)


@pytest.mark.parametrize("G", [50000])
def test_six_channel_backward_vs_oracle_at_benchmark_size(G):
    """n_sets = 2 (RGB + seg), 640x480: geometry-only (the timed instantiation) and full backward vs the C oracle (CH = 6)."""
    from gs_dynamics_b200 import rasterizer as R
    cam = make_camera(1, 640, 480)
    sc, act = make_scene(G, 0)
    seg = sc["seg_colors"] * torch.rand(G, 1, generator=torch.Generator().manual_seed(3))   # non-trivial second colour set
    col6 = torch.cat([act["colors_precomp"], seg], 1).contiguous()
    bg6 = torch.tensor([0.1, 0.2, 0.3, 0.05, 0.0, 0.4])
    dL = torch.randn(6, 480, 640, generator=torch.Generator().manual_seed(G))
    bo = oracle_backward(act, cam, bg6, dL, colors=col6)
    a = {k: v.cuda() for k, v in act.items()}
    st = settings_from(cam, bg6[:3].tolist())
    color, radii, depth, state = R.raster_forward(st, a["means3D"], a["opacities"], a["colors_precomp"], a["scales"], a["rotations"],
                                                  colors1=seg.cuda(), bg1=bg6[3:].cuda())
    geom = R.raster_backward(state, dL.cuda(), need_means2D=True, geom_only=True)
    full = R.raster_backward(state, dL.cuda(), need_means2D=True)
    torch.cuda.synchronize()
    assert geom["colors0"] is None and geom["opacities"] is None
    for k in ("means3D", "means2D", "scales", "rotations"):
        _check("geom-only %s" % k, geom[k].cpu(), bo[k])
        _check("full      %s" % k, full[k].cpu(), bo[k])
    _check("full      colors0", full["colors0"].cpu(), bo["colors"][:, :3])
    _check("full      colors1", full["colors1"].cpu(), bo["colors"][:, 3:])
    _check("full      opacities", full["opacities"].cpu().reshape(-1), bo["opacities"])


def _perturbed_problem(G):
    """workloads.tracking_problem(G) with the current state moved 0.5 mm / 0.01 away from the points the kNN distances were
    measured on.  At the unperturbed state the isometry residual |x_j - x_i| - d_ij is EXACTLY zero up to rounding, where the
    prior sqrt(w r^2 + 1e-20) is non-smooth: its gradient is sign(rounding noise) * sqrt(w) — O(1e-3) per Gaussian of noise in
    the reference as much as here (measured: the reference composition on the CPU, its eager CUDA version and this repo differ
    from one another by that much there).  One Adam step later (lr 1.6e-4) every real run has left that point."""
    from gs_dynamics_b200 import workloads
    prob = workloads.tracking_problem(G, 0)
    g = torch.Generator().manual_seed(G + 17)
    prob["params"]["means3D"] = (prob["params"]["means3D"] + 5e-4 * torch.randn(G, 3, generator=g)).contiguous()
    prob["params"]["unnorm_rotations"] = (prob["params"]["unnorm_rotations"] + 1e-2 * torch.randn(G, 4, generator=g)).contiguous()
    return prob


def _one_iteration_gpu_vs_cpu(G, use_graph):
    from gs_dynamics_b200 import tracking as TR, workloads
    from oracle import tracking_cpu
    prob = _perturbed_problem(G)
    params, variables, opt, dataset, _ = workloads.tracking_problem_gpu(G, 0, torch.device("cuda"), prob=prob)
    cam_id = 2
    # the oracle side: same state, same target images (the device-rendered ones), the reference's composition on the host
    cparams = {k: torch.nn.Parameter(v.clone().contiguous()) for k, v in prob["params"].items()}
    cparams["rgb_colors"].requires_grad = False
    cvars = dict(prob["variables"])
    cvars["max_2D_radius"] = cvars["max_2D_radius"].clone()
    copt = tracking_cpu.make_optimizer(cparams, cvars["scene_radius"])
    d = dataset[cam_id]
    cdata = {"cam": prob["cams"][cam_id]["mats"], "im": d["im"].cpu(), "seg": d["seg"].cpu(), "id": d["id"]}
    x0 = params["means3D"].detach().clone()
    q0 = params["unnorm_rotations"].detach().clone()
    step = TR.FusedTrackingStep(params, variables, opt, dataset, use_graph=use_graph)
    step.prepare()
    loss = float(step.step(cam_id))
    torch.cuda.synchronize()
    loss_ref = tracking_cpu.iteration(cparams, cdata, cvars, copt)
    print("G=%d graph=%s loss %.6f oracle %.6f" % (G, use_graph, loss, loss_ref))
    assert abs(loss - loss_ref) <= 2e-4 * abs(loss_ref)
    # gradients, read back from Adam's first moment (exp_avg = (1 - beta1) g after the first step from zero moments)
    for k in ("means3D", "unnorm_rotations"):
        m_gpu = opt.state[params[k]]["exp_avg"].cpu().numpy()
        m_ref = copt.state[cparams[k]]["exp_avg"].numpy()
        _check("%s gradient (G=%d)" % (k, G), m_gpu / 0.1, m_ref / 0.1)
        # the parameter step itself: first Adam step = -lr * g / (|g| + eps) = -lr * sign(g); compare where the gradient is
        # not rounding noise
        p0 = x0 if k == "means3D" else q0
        d_gpu = (params[k].detach() - p0).cpu().numpy()
        d_ref = cparams[k].detach().numpy() - prob["params"][k].numpy()
        big = np.abs(m_ref) > 1e-4 * np.abs(m_ref).max()
        agree = np.mean(np.sign(d_gpu[big]) == np.sign(d_ref[big]))
        print("%s step: sign agreement %.6f over %d entries, max |step| %.3g / %.3g" % (k, agree, big.sum(), np.abs(d_gpu).max(), np.abs(d_ref).max()))
        assert agree > 0.9995
        assert abs(np.abs(d_gpu).max() - np.abs(d_ref).max()) <= 1e-3 * np.abs(d_ref).max()
        assert float(opt.state[params[k]]["step"]) == 1.0
    assert np.array_equal(variables["seen"].cpu().numpy(), cvars["seen"].numpy())
    assert np.array_equal(variables["max_2D_radius"].cpu().numpy(), cvars["max_2D_radius"].numpy())


def test_fused_tracking_iteration_vs_oracle_iteration_50k():
    _one_iteration_gpu_vs_cpu(50000, use_graph=False)


def test_fused_tracking_iteration_vs_oracle_iteration_100k_graph_replay():
    """The exact thing bench.py times: the captured CUDA graph, 100k Gaussians (BASELINE configs[2] / north-star target size)."""
    _one_iteration_gpu_vs_cpu(100000, use_graph=True)
