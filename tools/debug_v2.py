import sys; sys.path.insert(0,'/root/repo')
import torch
from pygat_b200.functional import GatV2AttnFunction, EngineMatmul, _gemm
from pygat_b200.graph import Graph
from pygat_b200.synth import power_law_csr
DEV='cuda'
def rel(a,b): return ((a.double()-b.double()).abs().max()/b.double().abs().max()).item()
for (H,D,f_in) in [(4,64,40),(8,8,100),(1,16,12)]:
    n=4000
    rowptr,col=power_law_csr(n,12.0,seed=21,exponent=0.7,device=DEV)
    graph=Graph.from_csr(rowptr,col)
    g=torch.Generator(device=DEV).manual_seed(1)
    HD=H*D
    z=torch.randn(n,2*HD,generator=g,device=DEV)*0.5
    a=torch.randn(H,D,generator=g,device=DEV)*0.3
    gout=torch.randn(n,HD,generator=g,device=DEV)
    z1=z.clone().requires_grad_(True); a1=a.clone().requires_grad_(True)
    out=GatV2AttnFunction.apply(z1,a1,graph,H,D,False,0.2,True,None,1.0)
    out.backward(gout)
    # torch fp64 reference
    zd=z.double().requires_grad_(True); ad=a.double().requires_grad_(True)
    dst,src=graph.edge_index()
    whi=zd[:,:HD].view(n,H,D); whj=zd[:,HD:].view(n,H,D)
    u=whi[dst]+whj[src]
    s=(torch.where(u>0,u,0.2*u)*ad).sum(-1)   # E,H
    top=torch.full((n,H),-1e30,dtype=torch.float64,device=DEV).scatter_reduce(0,dst[:,None].expand(-1,H),s.detach(),reduce='amax')
    ex=torch.exp(s-top[dst])
    den=torch.zeros(n,H,dtype=torch.float64,device=DEV).index_add(0,dst,ex)
    agg=torch.zeros(n,H,D,dtype=torch.float64,device=DEV).index_add(0,dst,ex[:,:,None]*whi[src])
    ref=torch.nn.functional.elu((agg/den[:,:,None]).reshape(n,HD))
    ref.backward(gout.double())
    print(H,D,'out',rel(out,ref),'dz',rel(z1.grad,zd.grad),'dz_hi',rel(z1.grad[:,:HD],zd.grad[:,:HD]),'dz_hj',rel(z1.grad[:,HD:],zd.grad[:,HD:]),'da',rel(a1.grad,ad.grad))
    df=(z1.grad.double()-zd.grad).abs().max(1).values
    print('   worst rows', df.topk(3).indices.tolist(), 'deg', [(rowptr[i+1]-rowptr[i]).item() for i in df.topk(3).indices.tolist()])
    # GEMM check
    w=torch.randn(f_in,2*HD,generator=g,device=DEV)*0.1
    dx=torch.empty(n,f_in,device=DEV)
    dzz=z1.grad.contiguous()
    _gemm(0,1,n,f_in,2*HD,dzz,2*HD,w,2*HD,dx,f_in)
    print('   gemm dx', rel(dx, dzz.double()@w.double().t()))
