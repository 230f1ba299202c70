"""``tf.keras.optimizers.Adam`` as the reference trainer uses it (workers/trainer.py:138-139, 348-349, 420-425): an object per
model with ``apply_gradients(zip(grads, model.trainable_variables))``, on top of the C-ABI entry ``avd_adam_apply``.

The variables of a drop-in model (avddpg_b200.model) are views into ONE flat parameter vector whose trainable prefix is laid out in
``trainable_variables`` order, so a whole ``apply_gradients`` call is one kernel launch over that prefix; the Adam moments and the
step counter live in this object, like the slots of a Keras optimizer.  Formula (TF-Keras 2.4.1, non-amsgrad, epsilon 1e-7):
    t += 1;  lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m += (g - m)(1 - b1);  v += (g^2 - v)(1 - b2);  theta -= lr_t * m / (sqrt(v) + eps)
"""
from __future__ import annotations

import torch

from . import _lib


class Adam:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = float(learning_rate), float(beta_1), float(beta_2), float(epsilon)
        self._m = self._v = self._step = self._flat = None
        self._lib = _lib.load()

    @property
    def iterations(self) -> int:
        return 0 if self._step is None else int(self._step.item())

    def apply_gradients(self, grads_and_vars):
        pairs = list(grads_and_vars)
        if not pairs:
            return
        grads, variables = zip(*pairs)
        first = variables[0]
        if not (torch.is_tensor(first) and first.is_cuda):
            raise TypeError("apply_gradients expects the trainable_variables of an avddpg_b200.model network (CUDA views)")
        n = sum(int(v.numel()) for v in variables)
        # the variables must be the contiguous trainable prefix of one flat parameter vector, in order
        ptr = first.data_ptr()
        for v in variables:
            if v.data_ptr() != ptr or not v.is_contiguous():
                raise ValueError("variables are not the trainable prefix of one model, in trainable_variables order")
            ptr += v.numel() * 4
        if self._m is None:
            self._m = torch.zeros(1, n, dtype=torch.float32, device=first.device)
            self._v = torch.zeros_like(self._m)
            self._step = torch.zeros(1, dtype=torch.int32, device=first.device)
            self._flat = torch.zeros(1, n, dtype=torch.float32, device=first.device)
        elif self._m.shape[1] != n:
            raise ValueError("this optimizer has slots for a different model")
        off = 0
        for g, v in zip(grads, variables):
            k = int(v.numel())
            self._flat[0, off:off + k].copy_(torch.as_tensor(g, dtype=torch.float32, device=first.device).reshape(-1))
            off += k
        import ctypes as C
        _lib.check(self._lib.avd_adam_apply(C.c_void_p(first.data_ptr()), n, _lib.ptr(self._flat), n, _lib.ptr(self._m), _lib.ptr(self._v),
                                            _lib.ptr(self._step), None, 1, n, self.learning_rate, self.beta_1, self.beta_2, self.epsilon,
                                            _lib.current_stream()))
