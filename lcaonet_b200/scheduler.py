"""Learning-rate schedule of the reference's training utilities (lcaonet/train/scheduler.py:10-89): linear warm-up from
the base rate to `lr_max`, then cosine annealing with period 2 * T_max between the current rate and `eta_min`, damped by
a slow global cosine decay and clamped to [lr_min, lr_max].  Host-side only — nothing here touches the GPU path; it
exists so that a training script written against the reference runs unchanged."""
from __future__ import annotations

import math

from torch.optim import Optimizer
from torch.optim.lr_scheduler import LRScheduler


class WarmupCosineDecayAnnealingLR(LRScheduler):
    def __init__(self, optimizer: Optimizer, num_epoch: int, num_warmup: int, T_max: int, eta_min: float = 1e-7,
                 lr_max: float = 1e-3, lr_min: float = 1e-10, decay_coef: float = 1.5, last_epoch: int = -1,
                 verbose: bool = False):
        if num_warmup >= num_epoch:
            raise ValueError("Please set 'num_warmup' lower than 'num_epoch'")
        if decay_coef <= 0:
            raise ValueError("Please set 'decay_coef' higher than 0.")
        self.num_epoch, self.num_warmup, self.T_max = num_epoch, num_warmup, T_max
        self.eta_min, self.lr_max, self.lr_min, self.decay_coef = eta_min, lr_max, lr_min, decay_coef
        super().__init__(optimizer, last_epoch)  # (`verbose` is accepted for signature compatibility; torch dropped it)

    def _clamp(self, lr: float) -> float:
        return max(min(lr, self.lr_max), self.lr_min)

    def get_lr(self):
        t = self.last_epoch
        if t == 0:
            return list(self.base_lrs)
        if t < self.num_warmup:  # straight line from the base rate (epoch 0) to lr_max (epoch num_warmup - 1)
            return [b + t * (self.lr_max - b) / (self.num_warmup - 1) for b in self.base_lrs]
        k = t - self.num_warmup  # epochs into the annealing phase
        damp = math.cos(k / self.decay_coef / self.num_epoch * math.pi / 2)
        current = [group["lr"] for group in self.optimizer.param_groups]
        if (k - 1 - self.T_max) % (2 * self.T_max) == 0:  # restart step of the chained cosine-annealing form
            step = (self.lr_max - self.eta_min) * (1 - math.cos(math.pi / self.T_max)) / 2
            return [self._clamp(damp * (lr + step)) for lr in current]
        ratio = (1 + math.cos(math.pi * k / self.T_max)) / (1 + math.cos(math.pi * (k - 1) / self.T_max))
        return [self._clamp(damp * (ratio * (lr - self.eta_min) + self.eta_min)) for lr in current]
