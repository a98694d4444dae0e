"""Debugging aid for the round-1 finding "smoke() parity moves under ncu": run the sp_head_hub golden through the
engine with every workspace traced, save all buffers, and diff two such dumps.
  python tools/diag_nondet.py run <tag>          -> gpurun_out/diag_<tag>.pt
  python tools/diag_nondet.py cmp <tagA> <tagB>
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch


def run(tag, case="sp_head_hub"):
    import layers
    from pygat_b200 import _mem
    from tests.golden_io import dense_adj, load, rel_err
    from oracle import gat_oracle as O
    dev = torch.device("cuda:0")
    d = load(case)
    adj = dense_adj(d)
    xo = d["x"].clone().requires_grad_(True)
    Wo, ao, so = (d[k].clone().requires_grad_(True) for k in ("W", "a", "skip"))
    yo = O.sparse_head(xo, Wo, ao, O.edge_list(adj), d["alpha"], True, so, 0.0)
    yo.backward(d["gout"])
    f_in, dd = d["W"].shape
    head = layers.SpGraphAttentionLayer(f_in, dd, dropout=0.0, alpha=d["alpha"], concat=True, skip_connection=True)
    with torch.no_grad():
        head.W.copy_(d["W"]); head.a.copy_(d["a"]); head.skip_projection.copy_(d["skip"])
    head = head.to(dev)
    x = d["x"].to(dev).requires_grad_(True)
    _mem.trace = []
    y = head(x, adj.to(dev))
    y.backward(d["gout"].to(dev))
    torch.cuda.synchronize()
    errs = {"y": rel_err(y, yo), "dx": rel_err(x.grad, xo.grad), "dW": rel_err(head.W.grad, Wo.grad),
            "da": rel_err(head.a.grad, ao.grad), "dskip": rel_err(head.skip_projection.grad, so.grad)}
    print(tag, "rel errors vs oracle:", {k: f"{v:.2e}" for k, v in errs.items()}, flush=True)
    out = {"y": y.detach().cpu(), "dx": x.grad.cpu(), "dW": head.W.grad.cpu(), "da": head.a.grad.cpu(),
           "yo": yo.detach(), "trace": [t.detach().cpu() for t in _mem.trace]}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    torch.save(out, os.path.join(ROOT, "gpurun_out", f"diag_{tag}.pt"))


def cmp(a, b):
    A = torch.load(os.path.join(ROOT, "gpurun_out", f"diag_{a}.pt"))
    B = torch.load(os.path.join(ROOT, "gpurun_out", f"diag_{b}.pt"))
    for k in ("y", "dx", "dW", "da"):
        df = (A[k] - B[k]).abs()
        print(k, "max diff", df.max().item(), "rows differing", (df.reshape(df.shape[0], -1).max(1).values > 0).sum().item())
    print(len(A["trace"]), len(B["trace"]))
    for i, (ta, tb) in enumerate(zip(A["trace"], B["trace"])):
        if ta.shape != tb.shape or ta.dtype != tb.dtype:
            print(i, "shape/dtype mismatch", ta.shape, tb.shape)
            continue
        if ta.is_floating_point():
            same = torch.equal(torch.nan_to_num(ta, nan=12345.0), torch.nan_to_num(tb, nan=12345.0))
            df = (ta - tb).abs()
            df = torch.nan_to_num(df, nan=0.0)
            print(i, tuple(ta.shape), ta.dtype, "bit-equal" if same else f"DIFF max {df.max().item():.3e} n={int((df > 0).sum())}",
                  "nan:", int(torch.isnan(ta).sum()), int(torch.isnan(tb).sum()))
        else:
            print(i, tuple(ta.shape), ta.dtype, "equal" if torch.equal(ta, tb) else "DIFF")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2], *(sys.argv[3:4]))
    else:
        cmp(sys.argv[2], sys.argv[3])
