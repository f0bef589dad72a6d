"""No-grad render wrapper of the reference (A13): /root/reference/src/render/renderer.py:8-50 (1280x720, near 0.01, far 100,
grey background) and GSTrainer.render (/root/reference/src/real_world/gs/trainer.py:54-62)."""
import torch

from .rasterizer import GaussianRasterizer
from .tracking import setup_camera


class Renderer:
    def __init__(self, device, w=1280, h=720, near=0.01, far=100.0):
        self.device, self.w, self.h, self.near, self.far = device, w, h, near, far
        self.remove_background = False

    def setup_camera(self, k, w2c, bg):
        return setup_camera(self.w, self.h, k, w2c, near=self.near, far=self.far, bg=bg, device=self.device)

    @torch.no_grad()
    def render(self, w2c, k, timestep_data, bg=(0.7, 0.7, 0.7)):
        """timestep_data: dict(means3D, colors_precomp, rotations, opacities, scales, means2D) -> (image [3,H,W], depth [1,H,W])."""
        data = {key: v.to(self.device) for key, v in timestep_data.items()}
        im, _, depth = GaussianRasterizer(raster_settings=self.setup_camera(k, w2c, list(bg)))(**data)
        return im, depth

    @torch.no_grad()
    def render_mask(self, w2c, k, timestep_data):
        """All-ones colours on a black background (how /root/reference/src/predict.py:116-123 produces object masks)."""
        data = dict(timestep_data)
        data['colors_precomp'] = torch.ones_like(data['colors_precomp'])
        return self.render(w2c, k, data, bg=(0.0, 0.0, 0.0))[0][0]
