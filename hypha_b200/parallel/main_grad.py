"""fp32 main gradients for gradient accumulation.

The wgrad GEMMs write bf16 gradients straight into the flat gradient buffer, and with
``param.grad_accumulate`` their epilogue adds to what is already there -- in bf16.  That is exact enough for
one micro-batch per inner step (the shipped default, and all the reference does: its accelerate loop has no
accumulation, training.py:105-125), but a running bf16 sum over k micro-batches rounds k-1 times and starts
dropping contributions once the sum is ~2^8 times larger than an addend.

``MainGrad`` keeps the running sum in fp32 instead: every micro-batch's gradients are produced in OVERWRITE mode
(one bf16 rounding of that micro-batch's own fp32 accumulator, as today), added to an fp32 buffer, and the mean is
rounded to bf16 ONCE before the fused AdamW reads it.  Cost per micro-batch: one read of the bf16 gradients and
one read-modify-write of the fp32 buffer (10 bytes per parameter; ~10 ms for 8 B parameters at the measured
6.5 TB/s) and 4 bytes per parameter of HBM (32 GB for 8 B parameters, inside the 180 GB budget next to the
16+16+32+32+32 GB of weights, gradients, master weights and moments).
"""
from __future__ import annotations

from typing import Optional

import torch


class MainGrad:
    """fp32 running sum over micro-batches of a flat (bf16) gradient buffer."""

    def __init__(self, grads: torch.Tensor):
        self.grads = grads
        self.acc: Optional[torch.Tensor] = None       # allocated on first use: a plain inner step never pays for it
        self.count = 0

    def add(self) -> None:
        """Fold the gradients of the micro-batch that just finished its backward into the fp32 sum."""
        if self.acc is None:
            self.acc = torch.zeros(self.grads.shape, dtype=torch.float32, device=self.grads.device)
        if self.count == 0:
            self.acc.copy_(self.grads)                 # first micro-batch: no need to clear, then add
        else:
            self.acc.add_(self.grads)
        self.count += 1

    def finalize(self) -> int:
        """Write the MEAN over the accumulated micro-batches into the gradient buffer (one rounding) and reset.
        Returns how many micro-batches went in."""
        k = self.count
        if k == 0:
            raise RuntimeError("MainGrad.finalize() without a micro-batch")
        if k > 1:
            self.acc.mul_(1.0 / k)
        self.grads.copy_(self.acc)
        self.count = 0
        return k

    def release(self) -> None:
        self.acc, self.count = None, 0
