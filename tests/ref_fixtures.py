"""Synthetic data in the reference's ON-DISK formats, and run directories in which the reference's own, unmodified
training scripts (train.py, train_ppi.py) execute against this repository's `layers` / `models`.

The reference ships no features for Pubmed / PPI (.MISSING_LARGE_BLOBS) and nothing of /root/reference exists on
the GPU box, so: (1) `__graft_entry__.build()` copies the reference's six .py files into the git-ignored
`baseline/_ref/` (it travels with the snapshot); (2) the writers below produce small datasets with the shapes the
scripts hard-code (Pubmed: 500 features / 3 classes, train.py:73-81; PPI: 50 features / 121 labels,
train_ppi.py:43-51) in the formats utils.load_data (utils.py:36-45) and load_data_ppi (load_data_ppi.py:124-137)
read; (3) `make_run_dir` assembles a directory holding the reference's scripts and loaders next to EITHER this
repository's layers.py / models.py (engine=True) or the reference's own (engine=False, CPU oracle runs)."""
import json
import os
import shutil

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
REF_FILES = ("train.py", "train_ppi.py", "utils.py", "load_data_ppi.py", "layers.py", "models.py")


def reference_available() -> bool:
    return all(os.path.exists(os.path.join(REF, f)) for f in REF_FILES)


def _sym_edges(n, avg_deg, rng):
    m = int(n * avg_deg / 2)
    r = rng.integers(0, n, m)
    c = rng.integers(0, n, m)
    keep = r != c
    return r[keep], c[keep]


def write_pubmed_fixture(run_dir, n=1500, seed=0):
    """pubmed_dgl/{features,labels,idx_train,idx_val,idx_test}.pt + adj_sparse.npz as get_pubmed.ipynb exports them."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    d = os.path.join(run_dir, "pubmed_dgl")
    os.makedirs(d, exist_ok=True)
    feats = (rng.random((n, 500)) < 0.1) * rng.random((n, 500))
    feats[:, 0] += 0.01  # no all-zero rows (normalize_features would divide by zero)
    torch.save(torch.tensor(feats, dtype=torch.float32), os.path.join(d, "features.pt"))
    torch.save(torch.tensor(rng.integers(0, 3, n), dtype=torch.long), os.path.join(d, "labels.pt"))
    perm = rng.permutation(n)
    torch.save(torch.tensor(perm[:60], dtype=torch.long), os.path.join(d, "idx_train.pt"))
    torch.save(torch.tensor(perm[60:360], dtype=torch.long), os.path.join(d, "idx_val.pt"))
    torch.save(torch.tensor(perm[360:860], dtype=torch.long), os.path.join(d, "idx_test.pt"))
    r, c = _sym_edges(n, 4.5, rng)
    adj = sp.coo_matrix((np.ones(r.size, dtype=np.float32), (r, c)), shape=(n, n)).tocsr()
    adj.data[:] = 1.0
    adj = adj.maximum(adj.T) + sp.eye(n, dtype=np.float32, format="csr")  # the real file stores self-loops too
    sp.save_npz(os.path.join(d, "adj_sparse.npz"), adj.tocsr())
    return n


def write_ppi_fixture(run_dir, train_sizes=(180, 140, 210, 120), valid_sizes=(110, 90), test_sizes=(100, 130), seed=0):
    """data/ppi/{split}_{feats,labels,graph_id}.npy + {split}_graph.json (node-link JSON carrying both the "links" key
    DGL writes and the "edges" key networkx >= 3.4 reads)."""
    rng = np.random.default_rng(seed)
    d = os.path.join(run_dir, "data", "ppi")
    os.makedirs(d, exist_ok=True)
    gid = 1
    for split, sizes in (("train", train_sizes), ("valid", valid_sizes), ("test", test_sizes)):
        n = int(sum(sizes))
        np.save(os.path.join(d, f"{split}_feats.npy"), rng.standard_normal((n, 50)).astype(np.float32))
        np.save(os.path.join(d, f"{split}_labels.npy"), (rng.random((n, 121)) < 0.3).astype(np.int64))
        ids, links, off = [], [], 0
        for s in sizes:
            ids += [gid] * s
            r, c = _sym_edges(s, 8.0, rng)
            links += [{"source": int(off + a), "target": int(off + b)} for a, b in zip(r, c)]
            links += [{"source": int(off + b), "target": int(off + a)} for a, b in zip(r, c)]
            off += s
            gid += 1
        np.save(os.path.join(d, f"{split}_graph_id.npy"), np.array(ids, dtype=np.int64))
        graph = {"directed": False, "multigraph": False, "graph": {}, "nodes": [{"id": i} for i in range(n)],
                 "links": links, "edges": links}
        with open(os.path.join(d, f"{split}_graph.json"), "w") as fh:
            json.dump(graph, fh)


def make_run_dir(run_dir, engine: bool):
    """Copy the reference's scripts + loaders into run_dir, next to this repository's operator modules (engine=True)
    or the reference's own (engine=False).  Returns the PYTHONPATH the scripts need."""
    os.makedirs(run_dir, exist_ok=True)
    for f in ("train.py", "train_ppi.py", "utils.py", "load_data_ppi.py"):
        shutil.copy(os.path.join(REF, f), os.path.join(run_dir, f))
    for f in ("layers.py", "models.py"):
        shutil.copy(os.path.join(ROOT if engine else REF, f), os.path.join(run_dir, f))
    paths = [os.path.join(ROOT, "tests", "shims_ref")]
    paths.append(ROOT if engine else os.path.join(ROOT, "tests", "golden", "shims"))  # pygat_b200 / the torch_scatter shim
    return os.pathsep.join(paths)


RUNNER = ("import runpy, sys; sys.argv = sys.argv[1:]; "
          "exec('try:\\n    runpy.run_path(sys.argv[0], run_name=\"__main__\")\\nfinally:\\n"
          "    m = sys.modules.get(\"pygat_b200._lib\")\\n    print(\"ENGINE_CALLS\", m.call_count if m else 0, flush=True)')")
