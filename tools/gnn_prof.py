import sys; sys.path.insert(0, '/root/repo')
import torch
from gs_dynamics_b200 import gnn, workloads as GO
cfg = GO.sloth_cfg(512); dev = torch.device('cuda')
gi = GO.make_graph_inputs(2000, 1, 'sloth')
model = gnn.DynamicsPredictor(dict(cfg), dev).to(dev).eval(); model.load_state_dict(GO.make_state_dict(cfg, 0, head_scale=1e-3))
ro = gnn.GnnRollout(model, gi['state'][0, :, :2000].to(dev), gi['state'][0, :, 2000:].to(dev), 0.075, 8, True, use_graph=False)
d = torch.tensor([0.005, 0, 0], device=dev)
for _ in range(3): ro.step(d)
torch.cuda.synchronize()
