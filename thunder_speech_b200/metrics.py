"""Character / word error rate accumulators for ``CTCModule.validation_step`` -- the host-side counterpart of the
``torchmetrics`` objects the reference keeps (``CharErrorRate`` / ``WordErrorRate``, src/thunder/module.py:17-18,67-68,
157-162): errors = Levenshtein distance between prediction and target (characters, or whitespace-separated words),
rate = total errors / total target length, accumulated over ``update`` calls until ``reset``."""
from __future__ import annotations

from typing import List, Sequence, Union

import numpy as np


def edit_distance(a: Sequence, b: Sequence) -> int:
    """Levenshtein distance (unit costs).  One numpy pass per element of `a`: deletions and substitutions are elementwise,
    the insertion recurrence ``d[j] = min(d[j], d[j-1] + 1)`` is a running minimum of ``d[j] - j``."""
    n, m = len(a), len(b)
    if n == 0 or m == 0:
        return n + m
    table = {}
    bi = np.fromiter((table.setdefault(t, len(table)) for t in b), dtype=np.int64, count=m)
    ar = np.arange(m + 1, dtype=np.int64)
    prev = ar.copy()
    for i, t in enumerate(a, 1):
        code = table.get(t, -1)
        cur = np.empty(m + 1, dtype=np.int64)
        cur[0] = i
        np.minimum(prev[1:] + 1, prev[:-1] + (bi != code), out=cur[1:])
        cur = np.minimum.accumulate(cur - ar) + ar
        prev = cur
    return int(prev[m])


class ErrorRate:
    """``kind="char"``: CharErrorRate, ``kind="word"``: WordErrorRate (same call protocol as torchmetrics: calling the
    object updates the running totals and returns the rate of that batch; ``compute()`` gives the accumulated rate)."""

    def __init__(self, kind: str = "char"):
        if kind not in ("char", "word"):
            raise ValueError("kind must be 'char' or 'word'")
        self.kind = kind
        self.reset()

    def reset(self) -> None:
        self.errors = 0
        self.total = 0

    def _tokens(self, s: str):
        return s.split() if self.kind == "word" else s

    def update(self, preds: Union[str, List[str]], targets: Union[str, List[str]]) -> float:
        if isinstance(preds, str):
            preds = [preds]
        if isinstance(targets, str):
            targets = [targets]
        if len(preds) != len(targets):
            raise ValueError("ErrorRate: predictions and targets differ in length")
        e = t = 0
        for p, g in zip(preds, targets):
            pt, gt = self._tokens(p), self._tokens(g)
            e += edit_distance(pt, gt)
            t += len(gt)
        self.errors += e
        self.total += t
        return e / t if t else float("nan")

    __call__ = update

    def compute(self) -> float:
        return self.errors / self.total if self.total else float("nan")


def CharErrorRate() -> ErrorRate:
    return ErrorRate("char")


def WordErrorRate() -> ErrorRate:
    return ErrorRate("word")
