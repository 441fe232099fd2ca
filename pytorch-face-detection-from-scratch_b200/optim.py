"""Optimizer step on the flat parameter / gradient buffers of a BackboneEngine.

The reference trains with ``SAMSGD(Adam)`` whose ``step`` perturbs and un-perturbs the weights without
recomputing the gradients (models/ModelMeta.py:12-82), i.e. numerically ``torch.optim._multi_tensor.Adam``
(SURVEY.md 3.2).  ``FlatAdam`` is that update as ONE kernel over the engine's flat fp32 buffers -- which are also
the data-parallel all-reduce unit -- instead of ~10 foreach kernels over 44 tensors.
"""
from __future__ import annotations

import torch

from . import ops


class FlatAdam:
    """``capturable=True`` keeps the step count and the learning rate in device memory (the kernel advances the
    count), so ``step()`` can be captured into the train step's CUDA graph and replayed."""

    def __init__(self, engine, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 capturable: bool = False):
        self.engine, self.betas, self.eps, self.weight_decay = engine, betas, eps, weight_decay
        self.capturable = capturable
        self.m = self.v = self.state = None
        self.steps = 0
        self._lr = lr

    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):                   # e.g. the reference's MultiStepLR(milestones=[40], gamma=0.1)
        self._lr = float(value)
        if self.state is not None:
            self.state[2:3].copy_(torch.tensor([self._lr], dtype=torch.float32).view(torch.int32))

    def _ensure_state(self):
        eng = self.engine
        if self.m is None or self.m.device != eng.opt_params().device:
            self.m = torch.zeros_like(eng.opt_params())
            self.v = torch.zeros_like(eng.opt_params())
            if self.capturable:
                bits = torch.tensor([self._lr], dtype=torch.float32).view(torch.int32).item()
                self.state = torch.tensor([self.steps, 0, bits, 0], dtype=torch.int32, device=eng.opt_params().device)

    def step(self):
        eng = self.engine
        self._ensure_state()
        self.steps += 1                    # host-side mirror; under graph replay the device count is authoritative
        ops.adam_flat(eng.opt_params(), eng.opt_grads(), self.m, self.v, self._lr, self.betas[0], self.betas[1], self.eps,
                      self.weight_decay, self.steps, self.state)
        eng.weights_dirty = True          # the packed bf16 weights are rebuilt by the next forward

    def device_steps(self) -> int:
        return int(self.state[0].item()) if self.state is not None else self.steps

    def zero_grad(self):
        pass                               # run_backward overwrites the gradient buffer every step
