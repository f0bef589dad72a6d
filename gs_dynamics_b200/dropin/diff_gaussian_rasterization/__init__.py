"""Put ``gs_dynamics_b200/dropin`` on sys.path to make ``import diff_gaussian_rasterization`` resolve to the B200 kernels
(the two names the reference imports: /root/reference/src/tracking/helpers.py:5, src/tracking/train_utils.py:6)."""
from gs_dynamics_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians  # noqa: F401
