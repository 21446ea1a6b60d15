"""Initialisers imported at `/root/reference/experiments/optimized_layers.py:13`."""
import math


def glorot(tensor):
    # U(-a, a), a = sqrt(6 / (fan_in + fan_out)) over the last two dims
    if tensor is not None:
        a = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
        tensor.data.uniform_(-a, a)


def zeros(tensor):
    if tensor is not None:
        tensor.data.fill_(0)
