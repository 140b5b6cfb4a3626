"""``SGD`` / ``Adam`` and ``schedules.PiecewiseConstantDecay`` with Keras semantics
(reference: src/scripts/gcn.py:321-325 schedule + SGD, :338 ``apply_gradients``; SURVEY.md §8
a11).  ``apply_gradients`` over a model's full variable list is ONE fused kernel over the
flat parameter buffer (csrc/optim.cu); partial lists fall back to one launch per tensor."""
from __future__ import annotations

from . import _lib, ops


class PiecewiseConstantDecay:
    """values[0] for step <= boundaries[0]; values[k] for boundaries[k-1] < step <=
    boundaries[k]; values[-1] afterwards (tf.keras.optimizers.schedules)."""

    def __init__(self, boundaries, values, name=None):
        if len(values) != len(boundaries) + 1:
            raise ValueError("The length of boundaries should be 1 less than the length of values")
        self.boundaries, self.values = list(boundaries), list(values)

    def __call__(self, step):
        for b, v in zip(self.boundaries, self.values):
            if step <= b:
                return v
        return self.values[-1]


class schedules:  # namespace, like tf.keras.optimizers.schedules
    PiecewiseConstantDecay = PiecewiseConstantDecay


class _Optimizer:
    def __init__(self, learning_rate):
        self.learning_rate = learning_rate
        self.iterations = 0

    def _lr(self):
        lr = self.learning_rate
        return float(lr(self.iterations)) if callable(lr) else float(lr)

    @staticmethod
    def _flat(grads_and_vars):
        """If the list covers exactly one model's trainable variables in order, return the
        flat (params, grads) pair so that a single fused launch can be used."""
        gv = list(grads_and_vars)
        if not gv:
            return gv, None
        first = gv[0][1]
        base_w = getattr(first.value, "_base", None)
        base_w = first.value._base if first.value._base is not None else None
        if base_w is None:
            return gv, None
        total = 0
        expect = base_w.data_ptr()
        gbase = gv[0][0]._base if gv[0][0]._base is not None else None
        if gbase is None:
            return gv, None
        gexpect = gbase.data_ptr()
        for g, v in gv:
            if v.value._base is not base_w or v.value.data_ptr() != expect + 4 * total:
                return gv, None
            if g._base is not gbase or g.data_ptr() != gexpect + 4 * total:
                return gv, None
            total += v.value.numel()
        if total != base_w.numel() or gbase.numel() != total:
            return gv, None
        return gv, (base_w, gbase)

    def apply_gradients(self, grads_and_vars, grad_scale: float = 1.0):
        gv, flat = self._flat(grads_and_vars)
        if flat is not None:
            self._step(flat[0], flat[1], "flat", grad_scale)
        else:
            for g, v in gv:
                self._step(v.value.view(-1), g.reshape(-1), v.name, grad_scale)
        self.iterations += 1

    def apply_flat(self, params, grads, grad_scale: float = 1.0):
        """Fused step over explicit flat buffers (the data-parallel trainer's path)."""
        self._step(params, grads, "flat", grad_scale)
        self.iterations += 1


class SGD(_Optimizer):
    def __init__(self, learning_rate=0.01, momentum=0.0, nesterov=False, name="SGD"):
        if momentum != 0.0 or nesterov:
            raise NotImplementedError("the reference uses plain SGD (gcn.py:325); momentum is not built")
        super().__init__(learning_rate)

    def _step(self, w, g, key, grad_scale):
        ops.sgd_step(w, g, self._lr(), grad_scale)


class Adam(_Optimizer):
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, name="Adam"):
        if amsgrad:
            raise NotImplementedError("amsgrad is not built")
        super().__init__(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self._slots = {}

    def _step(self, w, g, key, grad_scale):
        torch = _lib.require_cuda()
        if key not in self._slots:
            self._slots[key] = (torch.zeros_like(w), torch.zeros_like(w))
        m, v = self._slots[key]
        ops.adam_step(w, g, m, v, self.iterations + 1, self._lr(), self.beta_1, self.beta_2, self.epsilon, grad_scale)
