"""``CategoricalCrossentropy`` and ``categorical_accuracy`` as the reference uses them
(src/scripts/gcn.py:6-7 imports, :326 ``loss_fn = CategoricalCrossentropy()``, :335,:353
calls, :339,:355 accuracy).  Keras evaluates the loss on the LOGITS of a softmax output
(it recovers them from the activation); the model attaches them to its output the same way
(``_gcs_logits``), and the native kernel computes loss, accuracy and dLoss/dlogits in one
launch (csrc/loss.cu).  Reduction is the Keras default: mean over the batch."""
from __future__ import annotations

from . import _lib, ops


class CategoricalCrossentropy:
    def __init__(self, from_logits=False, reduction="sum_over_batch_size", name="categorical_crossentropy"):
        if reduction not in ("sum_over_batch_size", "auto"):
            raise NotImplementedError("only the default mean-over-batch reduction is built")
        self.from_logits = from_logits
        self.name = name

    def __call__(self, y_true, y_pred):
        torch = _lib.require_cuda()
        y = _lib.as_tensor(y_true).to(device="cuda", dtype=torch.float32)
        logits = y_pred if self.from_logits else getattr(y_pred, "_gcs_logits", None)
        if logits is None:
            # probabilities of unknown origin: Keras' fallback (normalise, clip, -sum y log p)
            p = y_pred / y_pred.sum(dim=-1, keepdim=True)
            p = p.clamp(1e-7, 1.0 - 1e-7)
            return -(y * p.log()).sum(dim=-1).mean()
        tape = getattr(y_pred, "_gcs_tape", None)
        _, loss_acc, dlogits = ops.softmax_xent(logits, y, want_grad=tape is not None)
        loss = loss_acc[0]
        if tape is not None:
            tape.loss_record = dict(dlogits=dlogits, loss_acc=loss_acc)
        return loss


def categorical_accuracy(y_true, y_pred):
    """Per-sample 0/1 accuracy, like tf.keras.metrics.categorical_accuracy."""
    torch = _lib.require_cuda()
    y = _lib.as_tensor(y_true).to(device="cuda")
    return (y.argmax(dim=-1) == y_pred.argmax(dim=-1)).to(torch.float32)
