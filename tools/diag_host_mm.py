"""Host-side check behind the round-1 'parity moves under ncu' finding: repeat the oracle's small CPU products in one
process and count calls whose result differs from the first call's (a deterministic library on healthy hardware gives
0).  Prints the row range and size of every deviation class seen."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from tests.golden_io import load

d = load("sp_head_hub")
x, S, W = d["x"], d["skip"], d["W"]
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
print("threads", torch.get_num_threads(), flush=True)
for name, fn in (("sgemm x@skip", lambda: x.mm(S)), ("sgemm x@W", lambda: x.mm(W)),
                 ("dgemm x@skip", lambda: x.double().mm(S.double())), ("exp", lambda: torch.exp(x)),
                 ("sgemm 1 thread", None)):
    if fn is None:
        torch.set_num_threads(1)
        fn = lambda: x.mm(S)
    ref = fn()
    t0, calls, bad, seen = time.time(), 0, 0, {}
    while time.time() - t0 < secs:
        r = fn()
        calls += 1
        if not torch.equal(r, ref):
            bad += 1
            df = (r.double() - ref.double()).abs()
            rows = (df.reshape(df.shape[0], -1).max(1).values > 0).nonzero().flatten()
            key = (int(rows[0]), int(rows[-1]))
            if key not in seen:
                seen[key] = (df.max() / ref.abs().max()).item()
    print(f"{name}: calls {calls} deviating {bad} classes {seen}", flush=True)
