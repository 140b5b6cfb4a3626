"""O2 — PyTorch-CPU float32 restatement of the reference's op sequence with autograd
(TEST INFRASTRUCTURE; independent gradient check and the timed CPU baseline).

It executes what TensorFlow executes for /root/reference/src/scripts/gcn.py:328-340, op
by op and unfused: host scipy collate per step -> Dense / BN moments /
PReLU -> materialised ``gather`` of [nnz, H] messages -> segment sum (``index_add_``) ->
``Concatenate`` copies -> ``segment_sum`` pool -> post-MLP -> softmax cross-entropy ->
reverse-mode backward -> SGD apply.  It is a restatement ("port"), NOT TensorFlow; parity
is unpinned (oracle/__init__.py).
"""
from __future__ import annotations

import time

import numpy as np
import torch


def unpack(specs, w, s, requires_grad=True):
    """Flat numpy buffers -> list of per-block dicts of leaf tensors (float32)."""
    out = []
    for sp_ in specs:
        def leaf(rng, shape=None, buf=w):
            o, n = rng
            if n == 0:
                return None
            t = torch.tensor(np.asarray(buf[o:o + n], dtype=np.float32))
            if shape:
                t = t.reshape(shape)
            return t
        d = dict(W=leaf(sp_.kernel, (sp_.k_in, sp_.m_out)), b=leaf(sp_.bias), gamma=leaf(sp_.gamma),
                 beta=leaf(sp_.beta), alpha=leaf(sp_.alpha),
                 mm=leaf(sp_.moving_mean, buf=s), mv=leaf(sp_.moving_variance, buf=s), spec=sp_)
        if requires_grad:
            for k in ("W", "b", "gamma", "beta", "alpha"):
                if d[k] is not None:
                    d[k].requires_grad_(True)
        out.append(d)
    return out


def _block(p, x, training, eps, stats=None, branch=None):
    h = x @ p["W"] + p["b"]
    if training:
        mean = h.mean(0)
        var = ((h - mean.detach()) ** 2).mean(0)           # tf.nn.moments
        if stats is not None:
            stats.append((mean.detach(), var.detach()))
    else:
        mean, var = p["mm"], p["mv"]
    inv = torch.rsqrt(var + eps) * p["gamma"]              # tf.nn.batch_normalization
    z = h * inv + (p["beta"] - mean * inv)
    if p["alpha"] is not None:
        if branch is not None:
            # PReLU with the side of the kink pinned per element (see model_ref_np.loss_and_grads): identical to the
            # line below wherever sign(z) agrees with the pinned branch, i.e. everywhere but within rounding of 0
            br = torch.as_tensor(branch)
            return z * (br > 0).to(z.dtype) + p["alpha"] * z * (br < 0).to(z.dtype)
        return torch.relu(z) - p["alpha"] * torch.relu(-z)  # Keras PReLU
    return z


def forward(cfg, params, x, rows, cols, seg, n_graphs, training, stats=None, prelu_branch=None):
    P, L = cfg.pre_process, cfg.message_passing
    br = prelu_branch if prelu_branch is not None else [None] * len(params)
    out = x
    for k, p in enumerate(params[:P]):
        out = _block(p, out, training, cfg.bn_epsilon, stats, br[k])
    n = x.shape[0]
    for k, p in enumerate(params[P:P + L]):
        a = _block(p, out, training, cfg.bn_epsilon, stats, br[P + k])
        msgs = a[cols]                                      # tf.gather -> [nnz, H]
        z = torch.zeros(n, a.shape[1], dtype=a.dtype).index_add_(0, rows, msgs)
        out = torch.cat([z, out], dim=1) if cfg.connectivity == "cat" else (z + out if cfg.connectivity == "sum" else z)
    if cfg.pool == "sum":
        out = torch.zeros(n_graphs, out.shape[1], dtype=out.dtype).index_add_(0, seg, out)
    for k, p in enumerate(params[P + L:]):
        out = _block(p, out, training, cfg.bn_epsilon, stats, br[P + L + k])
    return out                                              # logits (pre-softmax)


def loss_and_grads(cfg, specs, w, s, x, rows, cols, seg, y, n_graphs, prelu_branch=None):
    params = unpack(specs, w, s)
    xt = torch.tensor(np.asarray(x, dtype=np.float32))
    rt, ct, st = (torch.tensor(np.asarray(v, dtype=np.int64)) for v in (rows, cols, seg))
    yt = torch.tensor(np.asarray(y, dtype=np.float32))
    stats = []
    logits = forward(cfg, params, xt, rt, ct, st, n_graphs, True, stats, prelu_branch)
    logp = torch.log_softmax(logits, dim=1)
    loss = -(yt * logp).sum(1).mean()
    loss.backward()
    grads = np.zeros(w.shape[0], dtype=np.float32)
    for p in params:
        sp_ = p["spec"]
        for key, rng in (("W", sp_.kernel), ("b", sp_.bias), ("gamma", sp_.gamma),
                         ("beta", sp_.beta), ("alpha", sp_.alpha)):
            o, n = rng
            if n:
                grads[o:o + n] = p[key].grad.reshape(-1).numpy()
    probs = torch.softmax(logits.detach(), dim=1).numpy()
    new_state = np.asarray(s, dtype=np.float32).copy()      # Keras moving statistics: moving -= (moving - batch) * (1 - momentum)
    for p, (mean, var) in zip(params, stats):
        for rng, old, batch in ((p["spec"].moving_mean, p["mm"], mean), (p["spec"].moving_variance, p["mv"], var)):
            o, n = rng
            new_state[o:o + n] = (old - (old - batch) * (1.0 - cfg.bn_momentum)).numpy()
    return dict(loss=float(loss.detach()), probs=probs, grads=grads, stats=stats, new_state=new_state)


def time_reference_path(cfg, specs, w, s, graphs, n_steps=3, warmup=1, train=True, lr=0.0002,
                        threads=None):
    """Timed CPU baseline: host scipy collate + forward (+ backward + SGD) per step on the
    same batch of ``graphs`` [(x, a, y)], as the reference does per batch.  Returns median
    seconds per step and the thread count used."""
    from . import batching_ref
    if threads:
        torch.set_num_threads(threads)
    params = unpack(specs, w, s, requires_grad=train)
    leaves = [p[k] for p in params for k in ("W", "b", "gamma", "beta", "alpha") if p[k] is not None]
    times = []
    for it in range(warmup + n_steps):
        t0 = time.perf_counter()
        (x, (indices, _vals, _shape), seg), y = batching_ref.collate(graphs)
        xt = torch.tensor(np.asarray(x, dtype=np.float32))
        rt = torch.tensor(indices[:, 0])
        ct = torch.tensor(indices[:, 1])
        st = torch.tensor(seg)
        yt = torch.tensor(np.asarray(y, dtype=np.float32))
        if train:
            logits = forward(cfg, params, xt, rt, ct, st, len(graphs), True)
            loss = -(yt * torch.log_softmax(logits, 1)).sum(1).mean()
            for t in leaves:
                t.grad = None
            loss.backward()
            with torch.no_grad():
                for t in leaves:
                    t -= lr * t.grad
        else:
            with torch.no_grad():
                forward(cfg, params, xt, rt, ct, st, len(graphs), False)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.median(times)), torch.get_num_threads()
