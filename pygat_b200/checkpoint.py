"""Checkpoint cadence of the reference's training loop without its per-epoch stall.

train.py:201-224 does, EVERY epoch: `torch.save(model.state_dict(), f'{epoch}_{dataset}.pkl')` (a synchronous
device->host copy of every parameter plus a pickle + file write on the training thread), then `glob('*.pkl')` and
deletes every file older than the best epoch; after the loop it deletes the files newer than the best epoch and
reloads `f'{best_epoch}_{dataset}.pkl'` (train.py:217-233).  Net effect: at any time the best epoch's file (and the
files after it) exist, and at the end exactly the best one is loaded.

`BestCheckpoint` keeps that contract -- same file name, same `state_dict` format (reference checkpoints and these are
interchangeable), best-so-far-by-validation-loss semantics incl. patience -- but (1) only snapshots when the epoch
is the new best (the files the reference writes for non-best epochs are never read), (2) copies the parameters to
pinned host memory on a side stream, so the training stream never waits, and (3) pickles / writes / prunes on a
background thread.  The validation loss may stay on the device: `update` accepts a tensor and resolves the
comparison lazily, `k` epochs later, so the training loop needs no per-epoch `.item()` either."""
from __future__ import annotations

import glob
import os
import queue
import threading
from typing import Dict, Optional

import torch


class BestCheckpoint:
    def __init__(self, model: torch.nn.Module, dataset: str, directory: str = ".", patience: int = 100, lag: int = 0):
        """lag: how many epochs the host may run behind the device when `update` is given device tensors (0 = resolve
        at once, like the reference's `.item()`)."""
        self.model, self.dataset, self.dir, self.patience, self.lag = model, dataset, directory, int(patience), int(lag)
        self.best = float("inf")
        self.best_epoch = -1
        self.bad = 0
        self.stop = False
        self._pending = []            # (epoch, loss tensor or float, snapshot or None)
        self._q: "queue.Queue" = queue.Queue()
        self._worker = threading.Thread(target=self._run, daemon=True)
        self._worker.start()
        self._side = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._errors = []

    def path(self, epoch: int) -> str:
        return os.path.join(self.dir, "{}_{}.pkl".format(epoch, self.dataset))   # train.py:201

    # ------------------------------------------------------------------ training-thread side
    def _snapshot(self) -> Dict[str, torch.Tensor]:
        """state_dict -> pinned host copies, asynchronously on a side stream (ordered after the work already queued on
        the training stream, i.e. after this epoch's optimizer step)."""
        sd = self.model.state_dict()
        if self._side is None or not any(v.is_cuda for v in sd.values()):
            return {k: v.detach().clone() for k, v in sd.items()}, None
        self._side.wait_stream(torch.cuda.current_stream())
        out = {}
        with torch.cuda.stream(self._side):
            for k, v in sd.items():
                buf = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                buf.copy_(v.detach(), non_blocking=True)
                v.record_stream(self._side)
                out[k] = buf
            done = torch.cuda.Event()
            done.record(self._side)
        return out, done

    def update(self, epoch: int, loss_val) -> bool:
        """Call once per epoch with the validation loss (float or 0-d tensor).  Returns True when training should stop
        (patience exhausted, train.py:209-210)."""
        # the snapshot must be taken NOW (the next step overwrites the parameters); whether it is kept is decided
        # when the loss is known
        self._pending.append((epoch, loss_val, self._snapshot()))
        while self._pending and (len(self._pending) > self.lag or not torch.is_tensor(self._pending[0][1])):
            self._resolve(*self._pending.pop(0))
        return self.stop

    def _resolve(self, epoch, loss_val, snap):
        loss = float(loss_val.item()) if torch.is_tensor(loss_val) else float(loss_val)
        if self.stop:
            return
        if loss < self.best:                      # train.py:202-205
            self.best, self.best_epoch, self.bad = loss, epoch, 0
            self._q.put(("save", epoch, snap))
        else:
            self.bad += 1                         # train.py:206-207
        if self.bad == self.patience:             # train.py:209-210
            self.stop = True

    def finish(self) -> int:
        """Resolve what is pending, wait for the writer, prune to the best file (train.py:219-224) and load it into the
        model (train.py:232-233).  Returns the best epoch."""
        while self._pending:
            self._resolve(*self._pending.pop(0))
        self._q.put(("stop", None, None))
        self._worker.join()
        if self._errors:
            raise self._errors[0]
        if self.best_epoch >= 0:
            self.model.load_state_dict(torch.load(self.path(self.best_epoch)))
        return self.best_epoch

    # ------------------------------------------------------------------ writer thread
    def _run(self):
        while True:
            op, epoch, snap = self._q.get()
            if op == "stop":
                return
            try:
                tensors, done = snap
                if done is not None:
                    done.synchronize()
                torch.save({k: v.clone() for k, v in tensors.items()}, self.path(epoch))
                for f in glob.glob(os.path.join(self.dir, "*_{}.pkl".format(self.dataset))):   # train.py:213-217
                    try:
                        nb = int(os.path.basename(f).split("_")[0])
                    except ValueError:
                        continue
                    if nb < epoch:
                        os.remove(f)
            except Exception as exc:  # surfaced by finish()
                self._errors.append(exc)
