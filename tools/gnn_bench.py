import sys, time; sys.path.insert(0, '/root/repo')
import torch
from gs_dynamics_b200 import gnn, workloads as GO
cfg = GO.sloth_cfg(512); dev = torch.device('cuda')
gi = GO.make_graph_inputs(2000, 1, 'sloth')
for mode in ('ieee', '3xtf32', 'tc'):
    model = gnn.DynamicsPredictor(dict(cfg), dev, matmul=mode).to(dev).eval(); model.load_state_dict(GO.make_state_dict(cfg, 0, head_scale=1e-3))
    ro = gnn.GnnRollout(model, gi['state'][0, :, :2000].to(dev), gi['state'][0, :, 2000:].to(dev), 0.075, 8, True, use_graph=True)
    d = torch.tensor([0.005, 0, 0], device=dev)
    for _ in range(5): ro.step(d)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): p = ro.step(d)
    e1.record(); torch.cuda.synchronize()
    print(mode, 'ms/step', e0.elapsed_time(e1) / 50, 'pred checksum', float(p.double().sum()))
# single-step agreement of the two GEMM modes against the fp64 oracle-free reference (ieee path in float64 is not available; compare modes)
ms = {m: gnn.DynamicsPredictor(dict(cfg), dev, matmul=m).to(dev).eval() for m in ('ieee', '3xtf32', 'tc')}
for m in ms.values(): m.load_state_dict(GO.make_state_dict(cfg, 0))
g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in gi.items()}
e = gnn.construct_edges_index(g['state'][0, -1], 0.075, g['state_mask'], g['eef_mask'], topk=8, connect_all=True)
with torch.no_grad():
    a = ms['ieee'](g['state'], g['attrs'], e, None, g['p_instance'], action=g['action'])[1]
    b = ms['3xtf32'](g['state'], g['attrs'], e, None, g['p_instance'], action=g['action'])[1]
print('single step |ieee - 3xtf32| max', float((a - b).abs().max()), 'motion max', float(a.abs().max()))
with torch.no_grad():
    c = ms['tc'](g['state'], g['attrs'], e, None, g['p_instance'], action=g['action'])[1]
print('single step |ieee - tc| max', float((a - c).abs().max()))
