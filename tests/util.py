"""Shared helpers for the parity tests."""
import math

import torch


def psnr(a, b):
    """metric.py:7-16 of the reference: PIXEL_MAX = 1, 20*log10(1/sqrt(mse))."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    if mse == 0:
        return 100.0
    return 20.0 * math.log10(1.0 / math.sqrt(mse))


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (torch.linalg.norm(a - b) / (torch.linalg.norm(b) + 1e-30)).item()


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (torch.dot(a, b) / (torch.linalg.norm(a) * torch.linalg.norm(b) + 1e-30)).item()
