"""Inert stand-in: the reference constructs lpips.LPIPS('alex') (attack_rd.py:581, train.py:47) but never calls it
on the hot path; the real package downloads weights, which is impossible offline."""
import torch


class LPIPS(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):
        raise NotImplementedError("lpips is not part of the attack hot path")
