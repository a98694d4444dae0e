import sys; sys.path.insert(0,'/root/repo')
import torch
from pygat_b200 import _mem
from pygat_b200.functional import gat_layer, random_masks, seeded_masks, padded_width
from pygat_b200.graph import Graph
from pygat_b200.synth import power_law_csr
DEV='cuda'
def rel(a,b): return ((a.double()-b.double()).abs().max()/b.double().abs().max()).item()
H,D,f_in=8,8,500
n,p,seed=3000,0.6,123456789
rowptr,col=power_law_csr(n,9.0,seed=5,exponent=0.7,device=DEV)
graph=Graph.from_csr(rowptr,col,seg_len=100000)
g=torch.Generator().manual_seed(2)
x=torch.randn(n,f_in,generator=g).to(DEV)
Ws=[(torch.randn(f_in,D,generator=g)*0.2).to(DEV) for _ in range(H)]
As=[(torch.randn(2*D,generator=g)*0.3).to(DEV) for _ in range(H)]
gout=torch.randn(n,H*D,generator=g).to(DEV)
Dp=padded_width(D)
mm=random_masks(n,f_in,H,Dp,graph.nnz,p,DEV,seed=seed)
out={}
for mode in ("seeded","materialised"):
    masks=seeded_masks(n,f_in,H,Dp,graph.nnz,seed=seed) if mode=="seeded" else mm
    Wd=[w.clone().requires_grad_(True) for w in Ws]; Ad=[a.clone().requires_grad_(True) for a in As]
    _mem.trace=[]
    y=gat_layer(x,graph,Wd,[a[:D] for a in Ad],[a[D:] for a in Ad],None,0.2,True,p=p,training=True,masks=masks)
    y.backward(gout); torch.cuda.synchronize()
    tr=_mem.trace; _mem.trace=None
    cand=[t for t in tr if tuple(t.shape)==(n,H*Dp) and t.dtype==torch.float32]
    print(mode, 'n cand', len(cand))
    out[mode]=(y.detach(), [w.grad for w in Wd], cand)
ys,yms=out['seeded'][0],out['materialised'][0]
print('y', rel(ys,yms))
for i,(a,b) in enumerate(zip(out['seeded'][2], out['materialised'][2])):
    df=(a-b).abs(); print('buf',i, rel(a,b), 'rows>1e-4:', int((df.max(1).values>1e-4*b.abs().max()).sum()))
# fp64 dW from each mode's dz (last candidate = dz_rows presumably)
keep=mm.keep_in.double()  # H,n,F
for mode in ("seeded","materialised"):
    for ci,dz in enumerate(out[mode][2]):
        ref=torch.stack([ (x.double()*keep[h]/(1-p)).t() @ dz[:, h*Dp:(h+1)*Dp].double() for h in range(H)],0)  # H,F,Dp
        got=torch.stack(out[mode][1],0)
        print(mode,'dW vs fp64 using buf',ci, rel(got, ref))
