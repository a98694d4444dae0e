"""pygat_b200.checkpoint.BestCheckpoint against the cadence of the reference's loop (train.py:196-233), replayed on a
loss sequence: same best epoch, same early stop, same surviving file, same restored weights."""
import glob
import os

import torch

from pygat_b200.checkpoint import BestCheckpoint


def reference_loop(model, losses, patience, directory, dataset):
    """train.py:196-233 verbatim in structure (save every epoch, prune older than best, stop on patience, prune newer,
    reload best); `losses[e]` stands in for train(epoch) and the weights are nudged every epoch like a training step."""
    best, best_epoch, bad = len(losses) + 1, 0, 0
    for epoch in range(len(losses)):
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
        torch.save(model.state_dict(), os.path.join(directory, "{}_{}.pkl".format(epoch, dataset)))
        if losses[epoch] < best:
            best, best_epoch, bad = losses[epoch], epoch, 0
        else:
            bad += 1
        if bad == patience:
            break
        for f in glob.glob(os.path.join(directory, "*.pkl")):
            if int(os.path.basename(f).split("_")[0]) < best_epoch:
                os.remove(f)
    for f in glob.glob(os.path.join(directory, "*.pkl")):
        if int(os.path.basename(f).split("_")[0]) > best_epoch:
            os.remove(f)
    model.load_state_dict(torch.load(os.path.join(directory, "{}_{}.pkl".format(best_epoch, dataset))))
    return best_epoch, epoch


def test_best_checkpoint_follows_the_reference_cadence(tmp_path):
    losses = [1.0, 0.9, 0.95, 0.8, 0.85, 0.81, 0.79, 0.9, 0.91, 0.92, 0.93, 0.5]
    for patience, lag in ((100, 0), (4, 0), (4, 3), (2, 1)):
        d_ref, d_new = tmp_path / f"ref{patience}{lag}", tmp_path / f"new{patience}{lag}"
        d_ref.mkdir(); d_new.mkdir()
        torch.manual_seed(0)
        m_ref = torch.nn.Linear(3, 2)
        m_new = torch.nn.Linear(3, 2)
        m_new.load_state_dict(m_ref.state_dict())
        best_ref, last_ref = reference_loop(m_ref, losses, patience, str(d_ref), "toy")
        ck = BestCheckpoint(m_new, "toy", str(d_new), patience=patience, lag=lag)
        for epoch, lv in enumerate(losses):
            with torch.no_grad():
                for p in m_new.parameters():
                    p.add_(1.0)
            # device-style lazy losses when lag > 0
            if ck.update(epoch, torch.tensor(lv) if lag else lv):
                break
        best_new = ck.finish()
        assert best_new == best_ref
        for a, b in zip(m_new.parameters(), m_ref.parameters()):
            assert torch.equal(a, b)
        files = sorted(os.path.basename(f) for f in glob.glob(str(d_new / "*.pkl")))
        assert files == [f"{best_ref}_toy.pkl"] == sorted(os.path.basename(f) for f in glob.glob(str(d_ref / "*.pkl")))
        sd = torch.load(str(d_new / files[0]))
        assert set(sd) == set(m_ref.state_dict())
