"""Evaluation step and metrics - the reference's ``evaluate(loader)``
(src/scripts/gcn.py:342-362) and its ROC post-processing (:389-398), with everything up to the
final curve kept on the device.

``evaluate`` runs ``model(inputs, training=False)`` (BatchNorm on moving statistics, folded into
scale/shift; gcs_model_forward) over one pass of the loader, evaluates loss and accuracy with the
fused softmax/cross-entropy kernel (gcs_softmax_xent) and returns the batch-size-weighted averages
plus the per-batch predictions, like the reference.  ``roc_curve`` / ``auc`` follow scikit-learn's
definitions (the reference calls ``sklearn.metrics.roc_curve(labels, probas)`` and ``auc``); they
are sort + prefix sums and run on whatever device the scores live on.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .losses import CategoricalCrossentropy, categorical_accuracy


def evaluate(model, loader, loss_fn=None):
    """One pass over ``loader`` -> (np.array([loss, acc]) weighted by batch size, [pred per batch]).
    Mirrors gcn.py:342-362; the only host transfers are the two scalars per batch."""
    torch = _lib.require_cuda()
    loss_fn = loss_fn or CategoricalCrossentropy()
    stats, preds = [], []
    step = 0
    while step < loader.steps_per_epoch:
        step += 1
        inputs, target = next(loader)
        pred = model(inputs, training=False)
        preds.append(pred)
        stats.append(torch.stack([loss_fn(target, pred).reshape(()), categorical_accuracy(target, pred).mean().reshape(()),
                                  torch.tensor(float(target.shape[0]), device=pred.device)]))
    out = torch.stack(stats).cpu().numpy().astype(np.float64)
    return np.average(out[:, :-1], 0, weights=out[:, -1]), preds


def positive_scores(preds, column: int = 1):
    """gcn.py:390-391: stack the per-batch softmax outputs and keep the positive-class column."""
    torch = _lib.require_cuda()
    return torch.cat([p[:, column] for p in preds])


def _as_torch(a, like=None):
    import torch
    if isinstance(a, torch.Tensor):
        return a
    t = torch.as_tensor(np.asarray(a))
    return t.to(like.device) if like is not None else t


def roc_curve(y_true, y_score, drop_intermediate: bool = True):
    """``sklearn.metrics.roc_curve`` for binary labels in {0, 1} (positive = 1): returns
    (fpr, tpr, thresholds) as NumPy arrays.  Sort and prefix sums run on the scores' device."""
    import torch
    s = _as_torch(y_score).reshape(-1).to(torch.float64)
    y = _as_torch(y_true, like=s).reshape(-1).to(s.device).to(torch.float64)
    if y.shape[0] != s.shape[0]:
        raise ValueError(f"y_true has {y.shape[0]} entries, y_score {s.shape[0]}")
    if y.shape[0] == 0:
        raise ValueError("roc_curve needs at least one sample")
    order = torch.argsort(s, descending=True, stable=True)
    s, y = s[order], y[order]
    n = s.shape[0]
    distinct = torch.nonzero(s[1:] != s[:-1]).reshape(-1)
    idx = torch.cat([distinct, torch.tensor([n - 1], device=s.device)])
    tps = torch.cumsum(y, 0)[idx]
    fps = (1 + idx).to(torch.float64) - tps
    thr = s[idx]
    if drop_intermediate and idx.shape[0] > 2:
        d2f, d2t = torch.diff(fps, n=2), torch.diff(tps, n=2)
        keep = torch.cat([torch.tensor([True], device=s.device), (d2f != 0) | (d2t != 0),
                          torch.tensor([True], device=s.device)])
        fps, tps, thr = fps[keep], tps[keep], thr[keep]
    zero = torch.zeros(1, dtype=torch.float64, device=s.device)
    tps, fps = torch.cat([zero, tps]), torch.cat([zero, fps])
    thr = torch.cat([torch.full((1,), float("inf"), dtype=torch.float64, device=s.device), thr])
    fpr = fps / fps[-1] if float(fps[-1]) > 0 else torch.full_like(fps, float("nan"))
    tpr = tps / tps[-1] if float(tps[-1]) > 0 else torch.full_like(tps, float("nan"))
    return fpr.cpu().numpy(), tpr.cpu().numpy(), thr.cpu().numpy()


def auc(x, y) -> float:
    """``sklearn.metrics.auc``: trapezoidal area of y over a monotonic x (sign follows the direction)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    if x.shape[0] != y.shape[0]:
        raise ValueError("x and y must have the same length")
    if x.shape[0] < 2:
        raise ValueError(f"At least 2 points are needed to compute area under curve, but x.shape = {x.shape[0]}")
    dx = np.diff(x)
    direction = 1.0
    if np.any(dx < 0):
        if np.all(dx <= 0):
            direction = -1.0
        else:
            raise ValueError("x is neither increasing nor decreasing : {}.".format(x))
    return float(direction * np.trapezoid(y, x))


def roc_auc(y_true, y_score) -> float:
    """Area under the ROC curve (fpr on the x axis)."""
    fpr, tpr, _ = roc_curve(y_true, y_score, drop_intermediate=False)
    return auc(fpr, tpr)
