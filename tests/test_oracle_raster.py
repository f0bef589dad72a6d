"""CPU: the C restatement (oracle/raster_oracle.c) against the independently written float64 autograd oracle
(oracle/raster_torch.py).  The reference ships no golden vectors for the rasterizer (parity unpinned);
this cross-check is what pins the oracle's hand-derived backward."""
import numpy as np
import pytest
import torch

from oracle import raster_c, raster_torch
from tests.helpers import make_camera, make_scene, oracle_backward, oracle_forward, rel_err


def _run_pair(G, w, h, seed, boost, bg, cam_id=0):
    cam = make_camera(cam_id, w, h)
    sc, act = make_scene(G, seed, scale_boost=boost, box_scale=0.5)
    bg_t = torch.tensor(bg)
    fo = oracle_forward(act, cam, bg_t, debug=True)
    ins = {k: v.clone().double().requires_grad_(True) for k, v in act.items()}
    m2d = torch.zeros(G, 3, dtype=torch.float64, requires_grad=True)
    to = raster_torch.rasterize_dense(ins["means3D"], m2d, ins["opacities"], ins["colors_precomp"], ins["scales"],
                                      ins["rotations"], cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"],
                                      cam["tanfovy"], h, w, bg_t)
    return cam, act, fo, to, ins, m2d, bg_t


@pytest.mark.parametrize("G,w,h,seed,boost,cam_id", [(300, 64, 48, 1, 1.0, 0), (500, 80, 64, 2, 0.7, 1), (200, 48, 48, 3, 1.5, 2)])
def test_forward_matches_dense_float64(G, w, h, seed, boost, cam_id):
    cam, act, fo, to, *_ = _run_pair(G, w, h, seed, boost, [0.1, 0.2, 0.3], cam_id)
    assert fo["R"] == to["R"] and fo["R"] > 0
    assert np.array_equal(fo["radii"], to["radii"].numpy())
    np.testing.assert_allclose(fo["color"], to["color"].detach().numpy(), atol=5e-6)
    np.testing.assert_allclose(fo["depth"], to["depth"].detach().numpy(), atol=5e-6)
    np.testing.assert_allclose(fo["final_T"], to["final_T"].detach().numpy(), atol=5e-6)
    assert np.array_equal(fo["n_contrib"], to["n_contrib"].numpy())


@pytest.mark.parametrize("G,w,h,seed,boost", [(300, 64, 48, 1, 1.0), (400, 64, 64, 5, 0.8)])
def test_backward_matches_autograd(G, w, h, seed, boost):
    cam, act, fo, to, ins, m2d, bg_t = _run_pair(G, w, h, seed, boost, [0.3, 0.1, 0.7])
    g = torch.Generator().manual_seed(seed)
    dL = torch.randn(3, h, w, dtype=torch.float64, generator=g)
    (to["color"] * dL).sum().backward()
    bo = oracle_backward(act, cam, bg_t, dL)
    assert rel_err(bo["means3D"], ins["means3D"].grad.numpy()) < 5e-5
    assert rel_err(bo["means2D"], m2d.grad.numpy()) < 5e-5
    assert rel_err(bo["colors"], ins["colors_precomp"].grad.numpy()) < 5e-5
    assert rel_err(bo["opacities"], ins["opacities"].grad.numpy().reshape(-1)) < 5e-5
    assert rel_err(bo["scales"], ins["scales"].grad.numpy()) < 5e-5
    assert rel_err(bo["rotations"], ins["rotations"].grad.numpy()) < 5e-5


def test_mask_invariant_ones_colour():
    """colour == 1, bg == 0  =>  image == 1 - final_T (how /root/reference/src/predict.py:116-123 makes masks)."""
    cam = make_camera(0, 96, 64)
    sc, act = make_scene(400, 7, scale_boost=0.8, box_scale=0.5)
    ones = torch.ones_like(act["colors_precomp"])
    fo = oracle_forward(act, cam, torch.zeros(3), colors=ones)
    np.testing.assert_allclose(fo["color"][0], 1.0 - fo["final_T"], atol=2e-6)


def test_empty_and_culled_inputs():
    cam = make_camera(0, 32, 32)
    sc, act = make_scene(5, 0)
    act["means3D"] = act["means3D"] + torch.tensor([0.0, 0.0, 100.0])  # far outside every frustum / behind cameras
    fo = oracle_forward(act, cam, torch.tensor([0.2, 0.4, 0.6]))
    assert fo["R"] == 0 and (fo["radii"] == 0).all()
    np.testing.assert_allclose(fo["color"][1], 0.4)


def test_known_answer_two_gaussians_on_the_optical_axis():
    """Closed-form known-answer test (tests/helpers.py::analytic_two_gaussians): both oracles against values computed by hand
    from the published formulas — an anchor that does not depend on either restatement."""
    from tests.helpers import analytic_two_gaussians
    K = analytic_two_gaussians()
    fo = oracle_forward(K["act"], K["cam"], K["bg"])
    v = K["valid"]
    assert np.array_equal(fo["radii"], K["radii"])
    assert np.abs(fo["color"] - K["color"])[:, v].max() < 2e-6
    assert np.abs(fo["depth"][0] - K["depth"])[v].max() < 2e-6
    assert np.abs(fo["final_T"] - K["final_T"])[v].max() < 2e-6
    a = K["act"]
    to = raster_torch.rasterize_dense(a["means3D"].double(), torch.zeros(2, 3, dtype=torch.float64), a["opacities"].double(),
                                      a["colors_precomp"].double(), a["scales"].double(), a["rotations"].double(),
                                      K["cam"]["viewmatrix"], K["cam"]["projmatrix"], K["cam"]["tanfovx"], K["cam"]["tanfovy"],
                                      64, 64, K["bg"])
    assert np.abs(to["color"].numpy() - K["color"])[:, v].max() < 1e-6
    assert np.abs(to["depth"].numpy()[0] - K["depth"])[v].max() < 1e-6


def test_known_answer_gradients_one_gaussian():
    """The C oracle's hand-derived backward against closed-form gradients (tests/helpers.py::analytic_one_gaussian_gradients):
    colour, opacity, mean x/y and scale gradients of one on-axis Gaussian, independent of the autograd cross-check."""
    from tests.helpers import analytic_one_gaussian_gradients
    K = analytic_one_gaussian_gradients()
    bo = oracle_backward(K["act"], K["cam"], K["bg"], K["dL"])
    g = K["grads"]
    tol = lambda ref: 2e-5 * np.abs(ref).max() + 1e-7
    assert np.abs(bo["colors"][0] - g["colors"]).max() < tol(g["colors"])
    assert abs(bo["opacities"][0] - g["opacity"]) < tol(np.array([g["opacity"]]))
    assert np.abs(bo["means3D"][0, :2] - g["mean"]).max() < tol(g["mean"])
    assert np.abs(bo["scales"][0] - g["scales"]).max() < tol(g["scales"])
