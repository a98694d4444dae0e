"""Which CPU op of the oracle changes between processes on the GPU box's host?  (Round-1 finding: smoke() parity
'moved' under ncu; the engine's outputs were bit-identical, the CPU oracle's were not.)  Prints, per process, the
error of each intermediate against an fp64 evaluation plus operand alignments."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import gat_oracle as O
from tests.golden_io import dense_adj, load

d = load("sp_head_hub")
adj = dense_adj(d)
edge = O.edge_list(adj)
x, W, a, S = d["x"], d["W"], d["a"], d["skip"]
n, dd = x.shape[0], W.shape[1]


def err(t, ref):
    df = (t.double() - ref).abs()
    rows = (df.reshape(df.shape[0], -1).max(1).values > 4e-6 * ref.abs().max()).nonzero().flatten()
    return f"{(df.max() / ref.abs().max()).item():.2e} rows[{rows.numel()}]{rows[:1].tolist()}..{rows[-1:].tolist()}"


x64, W64, a64, S64 = x.double(), W.double(), a.double(), S.double()
wh = x.mm(W)
wh64 = x64.mm(W64)
sk = x.mm(S)
cat = torch.cat((wh[edge[0]], wh[edge[1]]), dim=1).t()
lg = a.reshape(1, 2 * dd).mm(cat).squeeze(0)
lg64 = a64.reshape(1, 2 * dd).mm(torch.cat((wh.double()[edge[0]], wh.double()[edge[1]]), dim=1).t()).squeeze(0)
ex = torch.exp(lg - lg.max())
sp = torch.sparse_coo_tensor(edge, ex, (n, n), check_invariants=False)
agg = torch.matmul(sp, wh)
agg64 = torch.matmul(torch.sparse_coo_tensor(edge, ex.double(), (n, n), check_invariants=False), wh.double())
yo = O.sparse_head(x, W, a, edge, d["alpha"], True, S, 0.0)
print(f"threads={torch.get_num_threads()} x%64={x.data_ptr() % 64} W%64={W.data_ptr() % 64} | wh {err(wh, wh64)} | skip {err(sk, x64.mm(S64))} | "
      f"logit {err(lg[:, None], lg64[:, None])} | spmm {err(agg, agg64)} | y-vs-golden {err(yo, d['y'].double())}", flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "info":
    print(torch.__config__.show())
    print(torch.__config__.parallel_info())
    os.system("lscpu | head -30")
