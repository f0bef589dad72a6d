import sys; sys.path.insert(0,'/root/repo')
import ctypes as C
import torch, numpy as np
from tests.test_tracking_gpu import _tracking_problem
from tests.helpers import rel_err
from gs_dynamics_b200 import tracking as TR, rasterizer as R, _lib
pa,va,oa,da=_tracking_problem(2500)
data=da[0]; cap=None
# ---- eager pieces via autograd
rv=TR.params2rendervar(pa)
out,radius,_=TR.render_two_sets(data['cam'],rv,pa['seg_colors'])
out.retain_grad()
im=torch.exp(pa['cam_m'][0])[:,None,None]*out[:3]+pa['cam_c'][0][:,None,None]
L=50.0*TR.photometric_loss(im,data['im'])+200.0*TR.photometric_loss(out[3:],data['seg'])
rot=rv['rotations']; rot.retain_grad()
prior,parts=TR.track_prior_losses(rv['means3D'],rot,va,200.0,4.0,1000.0,200.0)
(L).backward(retain_graph=True)
dL_e=out.grad.clone(); grot_raster_e=rot.grad.clone(); gx_raster_e=pa['means3D'].grad.clone()
rot.grad=None; pa['means3D'].grad=None
prior.backward()
grot_prior_e=rot.grad.clone(); gx_prior_e=pa['means3D'].grad.clone()
# ---- fused pieces
with torch.no_grad():
    x=pa['means3D'].detach(); uq=pa['unnorm_rotations'].detach()
    rotf=torch.empty_like(uq)
    st=C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().gsd_track_normalize_rotations(x.shape[0],uq.data_ptr(),rotf.data_ptr(),st),'n')
    print('rot normalize maxdiff', float((rotf-rot.detach()).abs().max()))
    color,radii,_,state=R.raster_forward(data['cam'],x,torch.sigmoid(pa['logit_opacities']).reshape(-1).contiguous(),pa['rgb_colors'].detach(),torch.exp(pa['log_scales']).contiguous(),rotf,colors1=pa['seg_colors'].detach())
    print('color maxdiff', float((color-out.detach()).abs().max()))
    tgt=torch.cat([data['im'],data['seg']],0).contiguous()
    ws=TR._ph_workspace(color); ph=torch.empty(7,device='cuda')
    d=TR._ph_desc(color,tgt,2,0.8,0.2,(50.0,200.0),ws,affine=(pa['cam_m'][0],pa['cam_c'][0]))
    _lib.check(_lib.lib().gsd_photometric_forward(C.byref(d),ph.data_ptr(),st),'f')
    dL=torch.empty_like(color)
    _lib.check(_lib.lib().gsd_photometric_backward(C.byref(d),None,dL.data_ptr(),st),'b')
    print('photometric total', float(ph[6]), float(L), 'dL rel', rel_err(dL.cpu(),dL_e.cpu()), 'per set', rel_err(dL[:3].cpu(),dL_e[:3].cpu()), rel_err(dL[3:].cpu(),dL_e[3:].cpu()))
    g=R.raster_backward(state,dL,need_means2D=False)
    print('raster gx rel', rel_err(g['means3D'].cpu(),gx_raster_e.cpu()), 'grot rel', rel_err(g['rotations'].cpu(),grot_raster_e.cpu()))
    g2=R.raster_backward(state,dL_e.contiguous(),need_means2D=False)
    print('raster (eager dL) gx rel', rel_err(g2['means3D'].cpu(),gx_raster_e.cpu()), 'grot rel', rel_err(g2['rotations'].cpu(),grot_raster_e.cpu()))
    pr,pp=TR._TrackPriors.forward(TR._NullCtx(),x,rotf,va,(200.0,4.0,1000.0,2.0,200.0))
    gxp,gqp=TR._NullCtx.saved
    print('prior', float(pr), float(prior), 'gx rel', rel_err(gxp.cpu(),gx_prior_e.cpu()), 'gq rel', rel_err(gqp.cpu(),grot_prior_e.cpu()))
    print('magnitudes: raster grot', float(grot_raster_e.abs().max()), 'prior grot', float(grot_prior_e.abs().max()), 'raster gx', float(gx_raster_e.abs().max()), 'prior gx', float(gx_prior_e.abs().max()))
