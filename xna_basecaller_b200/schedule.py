"""Learning-rate schedules with the reference's surface (ub-bonito/bonito/schedule.py): each *_schedule returns a function
of the training progress t in [0, 1]; func_scheduler (:108-118) wraps one into a torch LambdaLR with an optional linear
warm-up; linear_warmup_cosine_decay (:7-17) is the trainer's default (training.py:83)."""
import math

import numpy as np
from torch.optim.lr_scheduler import LambdaLR


def const_schedule(y):
    return lambda t: y


def linear_schedule(y0, y1):
    return lambda t: y0 + (y1 - y0) * t


def cosine_decay_schedule(y0, y1):
    return lambda t: y1 + 0.5 * (y0 - y1) * (np.cos(t * np.pi) + 1.0)


def inverse_sqrt_decay_schedule(scale):
    return lambda t: 1.0 / math.sqrt(1 + scale * t)


def piecewise_schedule(knots, funcs):
    """funcs[i] on the i-th interval between the knots, each re-parametrised to [0, 1]."""
    def f(t):
        i = int(np.searchsorted(knots, t))
        t0 = 0.0 if i == 0 else knots[i - 1]
        t1 = 1.0 if i == len(knots) else knots[i]
        return funcs[i]((t - t0) / (t1 - t0))
    return f


def lr_factor(func, total_steps, warmup_steps=None, warmup_ratio=0.1, start_step=0):
    """step -> multiplier of the base learning rate (the lambda func_scheduler hands to LambdaLR)."""
    if warmup_steps:
        y0 = func(0.0)
        func = piecewise_schedule([warmup_steps / total_steps], [linear_schedule(warmup_ratio * y0, y0), func])
    return lambda step: func((step + start_step) / total_steps)


def func_scheduler(optimizer, func, total_steps, warmup_steps=None, warmup_ratio=0.1, start_step=0):
    return LambdaLR(optimizer, lr_factor(func, total_steps, warmup_steps, warmup_ratio, start_step))


def linear_warmup_cosine_decay(end_ratio=0.01, warmup_steps=500, **kwargs):
    return lambda optimizer, train_loader, epochs, last_epoch: func_scheduler(
        optimizer=optimizer, func=cosine_decay_schedule(1.0, end_ratio), total_steps=epochs * len(train_loader),
        warmup_steps=warmup_steps, start_step=last_epoch * len(train_loader))


def linear_cooldown(end_ratio=0.0, **kwargs):
    return lambda optimizer, train_loader, epochs, last_epoch: func_scheduler(
        optimizer=optimizer, func=linear_schedule(1.0, end_ratio), total_steps=epochs * len(train_loader), start_step=0)
