import sys; sys.path.insert(0,'/root/repo')
import torch
from oracle import gat_oracle as O
from pygat_b200.functional import gat_v2_layer
from pygat_b200.graph import Graph
from pygat_b200.synth import power_law_csr
from tests.golden_io import rel_err
DEV='cuda'
H,D,f_in,skip,concat=4,64,40,False,True
n=4000
rowptr,col=power_law_csr(n,12.0,seed=21,exponent=0.7)
g=torch.Generator().manual_seed(13)
x=torch.randn(n,f_in,generator=g)
Ws=[torch.randn(2*f_in,D,generator=g)*O.xavier_std(2*f_in,D) for _ in range(H)]
As=[torch.randn(1,D,generator=g)*O.xavier_std(1,D) for _ in range(H)]
gout=torch.randn(n,H*D,generator=g)
edge=O.PatternAdj(rowptr,col).nonzero().t()
xo=x.double().requires_grad_(True)
Wo=[w.double().requires_grad_(True) for w in Ws]
Ao=[a.double().requires_grad_(True) for a in As]
yo=torch.cat([O.sparse_head_v2(xo,w,a,edge,0.2,concat,None,faithful=False) for w,a in zip(Wo,Ao)],1)
yo.backward(gout.double())
graph=Graph.from_csr(rowptr.to(DEV),col.to(DEV))
xd=x.to(DEV).requires_grad_(True)
Wd=[w.to(DEV).requires_grad_(True) for w in Ws]
Ad=[a.to(DEV).requires_grad_(True) for a in As]
y=gat_v2_layer(xd,graph,Wd,[a.reshape(-1) for a in Ad],None,0.2,concat)
y.backward(gout.to(DEV))
print('y',rel_err(y,yo),'dx',rel_err(xd.grad,xo.grad),'dW',[rel_err(a.grad,b.grad) for a,b in zip(Wd,Wo)],'da',[rel_err(a.grad,b.grad) for a,b in zip(Ad,Ao)])
df=(xd.grad.cpu().double()-xo.grad).abs()
i=df.max(1).values.topk(5).indices
print('worst rows',i.tolist(),'deg',[(rowptr[k+1]-rowptr[k]).item() for k in i.tolist()], df.max(1).values[i].tolist(), xo.grad.abs().max().item())
print(xd.grad[i[0]].cpu()[:8], xo.grad[i[0]][:8])
