"""CPU oracle for the GeneralGNN hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker / the timed
CPU baseline.  Nothing under ``gcn-string_b200/`` imports it; the product path has no CPU
fallback and fails loudly when the CUDA library is missing.

PARITY UNPINNED.  The arithmetic of the reference's hot path lives in third-party
packages that are not vendored and cannot be installed here: ``spektral`` (PyPI, version
unpinned; imported at /root/reference/src/scripts/gcn.py:4,8-10) on ``tensorflow``/Keras
(unpinned; TF 2.6-2.9 era per src/configs/env.yml:252 and gcn.py:328).  The reference
ships no tests, golden vectors or fixtures.  This oracle therefore
restates the PUBLISHED upstream algorithms, anchored on the reference's call sites
(gcn.py:316-317 loader, :320 model, :326 loss, :321-325 optimizer, :334-339 train step,
:351 eval).  What IS pinned: the index construction in ``batching_ref`` executes the
literal scipy/numpy sequence Spektral's collate runs, so integer outputs are checked
against the real libraries; the float path is cross-checked between two independent
restatements (NumPy float64 with a hand-derived backward vs PyTorch-CPU float32 autograd).

  batching_ref.py     O3  scipy/numpy disjoint collate (bit-exact authority for indices)
  model_ref_np.py     O1  NumPy float64 forward + manual backward + SGD/Adam (truth)
  model_ref_torch.py  O2  PyTorch-CPU float32 autograd in the reference's op sequence
                          (independent gradients; the timed CPU baseline)
"""
