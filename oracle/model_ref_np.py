"""O1 — NumPy float64 restatement of GeneralGNN forward, backward and optimizers
(TEST INFRASTRUCTURE; the numerical truth tolerances are measured against).

Reference call sites: /root/reference/src/scripts/gcn.py:320 (model), :326,:335 (loss),
:321-325,:338 (SGD + PiecewiseConstantDecay), :334,:351 (training / inference calls).
Upstream algorithms restated (parity unpinned, see oracle/__init__):

  GeneralGNN.call   out = pre(x); for conv: z = conv([out, a]); out = concat([z, out]);
                    out = segment_sum(out, i); out = post(out)
  MLP layer         Dense -> BatchNormalization -> Dropout(0) -> PReLU | final activation
  GeneralConv.call  x.W + b -> BatchNormalization -> Dropout(0) -> PReLU -> propagate:
                    out[t] = sum_{(t,s) in a.indices} x[s]   (values of ``a`` never read)
  BatchNormalization(momentum=.99, eps=1e-3), training: biased two-pass moments,
                    y = h*inv + (beta - mean*inv), inv = gamma*rsqrt(var+eps);
                    moving -= (moving - batch) * (1 - momentum); inference: moving stats
  PReLU             f(z) = relu(z) - alpha*relu(-z), alpha per channel; d/dz at z == 0 is 0
  CategoricalCrossentropy on softmax output (Keras recovers the logits):
                    loss = mean_b( logsumexp(z_b) * sum(y_b) - y_b . z_b )
  categorical_accuracy   mean(argmax(y) == argmax(p))
"""
from __future__ import annotations

import numpy as np


class Block:
    """Parameters of one dense block as float64 arrays."""

    def __init__(self, spec, w, s):
        f = np.float64
        self.spec = spec
        o, n = spec.kernel
        self.W = w[o:o + n].astype(f).reshape(spec.k_in, spec.m_out)
        o, n = spec.bias
        self.b = w[o:o + n].astype(f)
        o, n = spec.gamma
        self.gamma = w[o:o + n].astype(f)
        o, n = spec.beta
        self.beta = w[o:o + n].astype(f)
        o, n = spec.alpha
        self.alpha = w[o:o + n].astype(f) if n else None
        o, n = spec.moving_mean
        self.mm = s[o:o + n].astype(f)
        o, n = spec.moving_variance
        self.mv = s[o:o + n].astype(f)


def _block_fwd(blk, x, training, eps):
    h = x @ blk.W + blk.b
    if training:
        mean = h.mean(0)
        var = ((h - mean) ** 2).mean(0)
    else:
        mean, var = blk.mm, blk.mv
    rstd = 1.0 / np.sqrt(var + eps)
    inv = blk.gamma * rstd
    z = h * inv + (blk.beta - mean * inv)
    a = np.where(z > 0, z, blk.alpha * z) if blk.alpha is not None else z
    return a, dict(x=x, h=h, mean=mean, var=var, rstd=rstd, z=z)


def _block_bwd(blk, c, da, grads, need_dx=True, branch=None):
    """Training-mode backward of one block; writes into the flat ``grads``.  ``branch`` (int8 [rows, width]: 1 / -1 / 0)
    overrides the side of PReLU's kink every element is differentiated on - see loss_and_grads."""
    sp_ = blk.spec
    z = c["z"]
    if blk.alpha is not None:
        side = np.sign(z) if branch is None else branch
        slope = np.where(side > 0, 1.0, np.where(side < 0, blk.alpha, 0.0))
        dz = da * slope
        o, n = sp_.alpha
        grads[o:o + n] = (da * np.minimum(z, 0.0)).sum(0)
    else:
        dz = da
    xhat = (c["h"] - c["mean"]) * c["rstd"]
    dgamma = (dz * xhat).sum(0)
    dbeta = dz.sum(0)
    n_rows = z.shape[0]
    dh = blk.gamma * c["rstd"] * (dz - dbeta / n_rows - xhat * (dgamma / n_rows))
    o, n = sp_.gamma
    grads[o:o + n] = dgamma
    o, n = sp_.beta
    grads[o:o + n] = dbeta
    o, n = sp_.kernel
    grads[o:o + n] = (c["x"].T @ dh).reshape(-1)
    o, n = sp_.bias
    grads[o:o + n] = dh.sum(0)
    return dh @ blk.W.T if need_dx else None


def spmm_sum(rows, cols, x, n):
    """out[t] = sum over entries (t, s) of x[s] — tf.gather + unsorted_segment_sum."""
    out = np.zeros((n, x.shape[1]), dtype=x.dtype)
    np.add.at(out, rows, x[cols])
    return out


def aggregate(rows, cols, x, n, agg="sum", w=None):
    """spektral.layers.ops scatter_sum / scatter_mean / scatter_max over messages w_ij * x[j]:
    tf.math.unsorted_segment_{sum,mean,max}(messages, targets, n).  An empty row gives 0 (sum, mean) or the lowest
    float (max).  Returns (out, ctx) with what the gradient needs."""
    msg = x[cols] if w is None else x[cols] * w[:, None]
    if agg == "sum":
        out = np.zeros((n, x.shape[1]), dtype=x.dtype)
        np.add.at(out, rows, msg)
        return out, None
    cnt = np.bincount(rows, minlength=n).astype(x.dtype)
    if agg == "mean":
        out = np.zeros((n, x.shape[1]), dtype=x.dtype)
        np.add.at(out, rows, msg)
        return out / np.maximum(cnt, 1.0)[:, None], cnt
    assert agg == "max"
    out = np.full((n, x.shape[1]), np.finfo(np.float32).min, dtype=x.dtype)
    np.maximum.at(out, rows, msg)
    return out, msg


def aggregate_bwd(rows, cols, dz, n, agg="sum", w=None, ctx=None, z=None):
    """Gradient of ``aggregate`` with respect to x.  max: tf's _UnsortedSegmentMinOrMaxGrad - the entries that attain a
    row's maximum share its gradient equally."""
    g = dz[rows]
    if agg == "mean":
        g = g / np.maximum(ctx, 1.0)[rows][:, None]
    if agg == "max":
        sel = (ctx == z[rows]).astype(dz.dtype)
        num = np.zeros_like(dz)
        np.add.at(num, rows, sel)
        g = sel * g / np.maximum(num, 1.0)[rows]
    if w is not None:
        g = g * w[:, None]
    out = np.zeros((n, dz.shape[1]), dtype=dz.dtype)
    np.add.at(out, cols, g)
    return out


def segment_sum(x, seg, n_seg):
    out = np.zeros((n_seg, x.shape[1]), dtype=x.dtype)
    np.add.at(out, seg, x)
    return out


def forward(cfg, specs, w, s, x, rows, cols, seg, n_graphs, training=False, edge_weight=None):
    """Returns (output [B,C] (probabilities if activation == 'softmax'), cache).  ``edge_weight``: optional per-entry
    weights (the reference's unfinished ``use_edge_data`` switch); GeneralConv itself ignores adjacency values."""
    blocks = [Block(sp_, w, s) for sp_ in specs]
    P, L = cfg.pre_process, cfg.message_passing
    n = x.shape[0]
    caches = []
    out = x.astype(np.float64)
    for blk in blocks[:P]:
        out, c = _block_fwd(blk, out, training, cfg.bn_epsilon)
        caches.append(c)
    for blk in blocks[P:P + L]:
        a, c = _block_fwd(blk, out, training, cfg.bn_epsilon)
        caches.append(c)
        ew = None if edge_weight is None else np.asarray(edge_weight, dtype=np.float64)
        z, c["agg_ctx"] = aggregate(rows, cols, a, n, getattr(cfg, "aggregate", "sum"), ew)
        c["agg_out"] = z
        # Concatenate()([z, out]) | Add()([z, out]) | no skip connection (GeneralGNN.call)
        out = np.concatenate([z, out], axis=1) if cfg.connectivity == "cat" else (z + out if cfg.connectivity == "sum" else z)
    node_out = out
    if cfg.pool == "sum":
        out = segment_sum(out, seg, n_graphs)
    for blk in blocks[P + L:]:
        out, c = _block_fwd(blk, out, training, cfg.bn_epsilon)
        caches.append(c)
    logits = out
    if cfg.activation == "softmax":
        e = np.exp(logits - logits.max(1, keepdims=True))
        out = e / e.sum(1, keepdims=True)
    return out, dict(blocks=blocks, caches=caches, logits=logits, node_out=node_out)


def xent_from_logits(logits, y):
    m = logits.max(1, keepdims=True)
    lse = (m + np.log(np.exp(logits - m).sum(1, keepdims=True)))[:, 0]
    per = lse * y.sum(1) - (y * logits).sum(1)
    return per.mean(), per


def accuracy(probs, y):
    return float((probs.argmax(1) == y.argmax(1)).mean())


def loss_and_grads(cfg, specs, w, s, x, rows, cols, seg, y, n_graphs, prelu_branch=None, edge_weight=None):
    """One training-mode forward + backward.  Returns dict(loss, acc, probs, grads (flat
    float64, same layout as w), new_state (flat float64 moving statistics)).

    ``prelu_branch``: optional list (one entry per block, None for blocks without PReLU) of int8 arrays giving the
    branch of PReLU (1: z > 0, -1: z < 0, 0: z == 0) another implementation differentiated each element on.  PReLU's
    derivative jumps at 0, so an input within float32 rounding of 0 may land on either side in a float32
    implementation; with the branches pinned the comparison measures arithmetic error only.  ``ctx['prelu_flips']``
    counts the elements whose pinned branch differs from the float64 sign."""
    assert cfg.activation == "softmax" and cfg.pool == "sum" and cfg.connectivity in ("cat", "sum", None)
    y = y.astype(np.float64)
    probs, ctx = forward(cfg, specs, w, s, x, rows, cols, seg, n_graphs, training=True, edge_weight=edge_weight)
    blocks, caches, logits = ctx["blocks"], ctx["caches"], ctx["logits"]
    loss, _ = xent_from_logits(logits, y)
    B = logits.shape[0]
    grads = np.zeros(w.shape[0], dtype=np.float64)
    d = (probs * y.sum(1, keepdims=True) - y) / B
    P, L, H = cfg.pre_process, cfg.message_passing, cfg.hidden
    br = prelu_branch if prelu_branch is not None else [None] * len(blocks)
    flips = 0
    for bi, (blk, c) in enumerate(zip(blocks, caches)):
        if br[bi] is not None and blk.alpha is not None:
            flips += int((np.sign(c["z"]) != br[bi]).sum())
    ctx["prelu_flips"] = flips
    for bi in range(len(blocks) - 1, P + L - 1, -1):
        d = _block_bwd(blocks[bi], caches[bi], d, grads, branch=br[bi])
    dout = d[seg]                                       # grad of segment_sum
    for k in range(L - 1, -1, -1):
        bi = P + k
        if cfg.connectivity == "cat":
            dz, dprev = dout[:, :H], dout[:, H:]
        else:                                           # Add: the gradient reaches both operands; None: only z
            dz, dprev = dout, (dout if cfg.connectivity == "sum" else 0.0)
        ew = None if edge_weight is None else np.asarray(edge_weight, dtype=np.float64)
        da = aggregate_bwd(rows, cols, dz, x.shape[0], getattr(cfg, "aggregate", "sum"), ew, caches[bi]["agg_ctx"],
                           caches[bi]["agg_out"])      # (sum, no weights: pattern(A)^T . dz)
        dout = dprev + _block_bwd(blocks[bi], caches[bi], da, grads, branch=br[bi])
    for bi in range(P - 1, -1, -1):
        dout = _block_bwd(blocks[bi], caches[bi], dout, grads, need_dx=bi > 0, branch=br[bi])
    new_state = s.astype(np.float64).copy()
    mom = cfg.bn_momentum
    for blk, c in zip(blocks, caches):
        o, n = blk.spec.moving_mean
        new_state[o:o + n] = blk.mm - (blk.mm - c["mean"]) * (1.0 - mom)
        o, n = blk.spec.moving_variance
        new_state[o:o + n] = blk.mv - (blk.mv - c["var"]) * (1.0 - mom)
    return dict(loss=float(loss), acc=accuracy(probs, y), probs=probs, grads=grads,
                new_state=new_state, ctx=ctx)


def piecewise_constant(step, boundaries, values):
    """tf.keras.optimizers.schedules.PiecewiseConstantDecay: values[0] for step <=
    boundaries[0], values[k] for boundaries[k-1] < step <= boundaries[k], else values[-1]."""
    for b, v in zip(boundaries, values):
        if step <= b:
            return v
    return values[-1]


def sgd_step(w, g, lr):
    """Keras SGD without momentum (gcn.py:325): w <- w - lr * g."""
    return w - lr * g


def adam_step(w, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras Adam (epsilon outside the bias-corrected root), t = 1-based step."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    return w - lr_t * m / (np.sqrt(v) + eps), m, v
