import sys; sys.path.insert(0,'/root/repo')
import torch
from gs_dynamics_b200 import gnn
dev=torch.device('cuda')
import os
M,N,K=int(os.environ.get("GM","20021")),512,512
x=torch.relu(torch.randn(M,K,device=dev)); w=torch.randn(N,K,device=dev)/K**0.5
ws=gnn._tc_split(w)
for i in range(3): y=gnn._tc_linear(x,ws,relu=True)
torch.cuda.synchronize()
