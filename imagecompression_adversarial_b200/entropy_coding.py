"""Entropy coding at the operator surface: the tables and strings behind
``EntropyBottleneck`` / ``GaussianConditional`` ``.update()`` / ``.compress()`` / ``.decompress()`` and
``model.update()`` / ``model.compress(x)`` / ``model.decompress(strings, shape)``.

Replaces compressai's CPU path (compressai/entropy_models/entropy_models.py on the C++ ``compressai.ans`` coder;
reference call sites ``attack_TIC.py:106-110``, ``InvCompress/attack_inv.py:112-116``, ``InvCompress/ours.py:100-175``,
``InvCompress/train.py:452``).  The range coder itself runs on the GPU (``csrc/icadv_rans.cu``): one thread per
(image, lane), ``lanes`` independent rANS streams per image, so a 64-image batch is coded by thousands of threads instead of
one CPU core; with ``lanes == 1`` the bytes are the ones compressai's coder emits for the same symbols and tables.
Only table construction (parameter-sized, once per ``update()``) is host code.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L
from .ops import _p, _stream

DEFAULT_LANES = 64


# ------------------------------------------------------------------------------------------ tables
def pmf_to_quantized_cdf(pmf, precision=16):
    """compressai ``pmf_to_quantized_cdf`` (host, parameter-sized): int32 [len(pmf) + 1]."""
    a = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.empty(a.size + 1, dtype=np.int32)
    L.call("icadv_pmf_to_quantized_cdf", a.ctypes.data_as(C.POINTER(C.c_float)), int(a.size), int(precision),
           out.ctypes.data_as(C.POINTER(C.c_int)))
    return out


def pmf_to_cdf(pmf, tail_mass, pmf_length, max_length):
    """compressai ``EntropyModel._pmf_to_cdf``: one quantised CDF row per channel / scale level."""
    pmf, tail_mass = pmf.detach().float().cpu(), tail_mass.detach().float().cpu()
    cdf = torch.zeros(len(pmf_length), max_length + 2, dtype=torch.int32)
    for i in range(len(pmf_length)):
        prob = torch.cat((pmf[i, : int(pmf_length[i])], tail_mass[i].reshape(-1)))
        row = pmf_to_quantized_cdf(prob.numpy())
        cdf[i, : row.size] = torch.from_numpy(row)
    return cdf


def get_scale_table(lo=0.11, hi=256.0, levels=64):
    """compressai.zoo / examples ``get_scale_table``."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


class CoderTables:
    """Device copies of (quantized_cdf [n, stride], cdf_length [n], offset [n])."""

    def __init__(self, cdf, length, offset, device):
        self.cdf = cdf.to(device=device, dtype=torch.int32).contiguous()
        self.length = length.to(device=device, dtype=torch.int32).contiguous()
        self.offset = offset.to(device=device, dtype=torch.int32).contiguous()
        if self.cdf.dim() != 2 or self.cdf.shape[0] != self.length.numel() or self.length.numel() != self.offset.numel():
            raise L.IcadvError("entropy coder tables: inconsistent shapes (was update() called?)")


# ------------------------------------------------------------------------------------------ coder
def _order_tensor(order, device):
    if order is None:
        return None
    return order.to(device=device, dtype=torch.int32).contiguous()


def rans_encode(symbols, indexes, tables, *, mode=0, order=None, lanes=DEFAULT_LANES):
    """symbols / indexes: int32 channels-last [N, H, W, C] on the device -> list of N byte strings."""
    if symbols.dtype != torch.int32 or indexes.dtype != torch.int32 or not symbols.is_cuda:
        raise L.IcadvError("rans_encode: int32 CUDA tensors expected")
    symbols, indexes = symbols.contiguous(), indexes.contiguous()
    n, c = symbols.shape[0], symbols.shape[-1]
    hw = symbols[0].numel() // c
    dev = symbols.device
    order = _order_tensor(order, dev)
    n_pos = int(order.numel()) if order is not None else hw
    lanes = int(min(lanes, c if mode == 1 else hw * c))
    cap = L.lib().icadv_rans_lane_capacity(n, hw, c, mode, n_pos, lanes)
    stride = lanes * cap + lanes
    scratch = torch.empty(n * lanes * cap, device=dev, dtype=torch.int32)
    lane_words = torch.empty(n * lanes, device=dev, dtype=torch.int32)
    packed = torch.empty(n, stride, device=dev, dtype=torch.int32)
    total = torch.empty(n, device=dev, dtype=torch.int32)
    L.call("icadv_rans_encode", _p(symbols), _p(indexes), n, hw, c, _p(tables.cdf), _p(tables.length), _p(tables.offset),
           int(tables.cdf.shape[1]), mode, _p(order), n_pos, lanes, _p(scratch), _p(lane_words), _p(packed), stride,
           _p(total), _stream())
    words = total.cpu()
    host = packed[:, : int(words.max())].cpu().numpy()
    return [host[i, : int(words[i])].tobytes() for i in range(n)]


def _upload(strings, device):
    n = len(strings)
    words = [len(s) // 4 for s in strings]
    if any(len(s) % 4 or len(s) < 8 for s in strings):
        raise L.IcadvError("rans_decode: a string is not a whole number of 32-bit words")
    stride = max(words)
    host = np.zeros((n, stride), dtype=np.uint32)
    for i, s in enumerate(strings):
        host[i, : words[i]] = np.frombuffer(s, dtype=np.uint32)
    return torch.from_numpy(host.view(np.int32)).to(device), stride


def rans_decode(strings, indexes, tables, *, means=None, mode=0, order=None, lanes=DEFAULT_LANES):
    """-> fp32 channels-last tensor of decoded symbols (+ means), shape of ``indexes``."""
    indexes = indexes.contiguous()
    n, c = indexes.shape[0], indexes.shape[-1]
    hw = indexes[0].numel() // c
    dev = indexes.device
    order = _order_tensor(order, dev)
    n_pos = int(order.numel()) if order is not None else hw
    lanes = int(min(lanes, c if mode == 1 else hw * c))
    packed, stride = _upload(strings, dev)
    out = torch.empty(indexes.shape, device=dev, dtype=torch.float32)
    if means is not None:
        means = means.contiguous()
    L.call("icadv_rans_decode", _p(packed), stride, _p(indexes), _p(means), _p(out), n, hw, c, _p(tables.cdf),
           _p(tables.length), _p(tables.offset), int(tables.cdf.shape[1]), mode, _p(order), n_pos, lanes, _stream())
    return out


class StepDecoder:
    """Incremental decoder of the autoregressive models: per-lane coder state stays on the device between steps."""

    def __init__(self, strings, tables, n_channels, device, lanes=DEFAULT_LANES):
        self.n = len(strings)
        self.tables = tables
        self.lanes = int(min(lanes, n_channels))
        self.packed, self.stride = _upload(strings, device)
        self.state = torch.empty(self.n * self.lanes, device=device, dtype=torch.int64)
        self.cursor = torch.empty(self.n * self.lanes, device=device, dtype=torch.int32)
        L.call("icadv_rans_decode_init", _p(self.packed), self.stride, self.n, self.lanes, _p(self.state),
               _p(self.cursor), _stream())

    def step(self, indexes, means, out, positions):
        """indexes / means / out: full channels-last tensors [N, H, W, C]; positions: int32 hw-indexes of this step."""
        c = indexes.shape[-1]
        hw = indexes[0].numel() // c
        t = self.tables
        L.call("icadv_rans_decode_step", _p(self.packed), self.stride, _p(self.state), _p(self.cursor), _p(indexes),
               _p(means), _p(out), self.n, hw, c, _p(t.cdf), _p(t.length), _p(t.offset), int(t.cdf.shape[1]),
               _p(positions), int(positions.numel()), self.lanes, _stream())


def build_indexes(scales, scale_table, bound=0.11):
    """compressai ``GaussianConditional.build_indexes`` on any dense fp32 CUDA tensor (same memory order out)."""
    out = torch.empty(scales.shape, device=scales.device, dtype=torch.int32)
    if scales.is_contiguous(memory_format=torch.channels_last) and scales.dim() == 4:
        out = out.contiguous(memory_format=torch.channels_last)
    elif not scales.is_contiguous():
        scales = scales.contiguous()
    L.call("icadv_build_indexes", _p(scales), _p(scale_table), int(scale_table.numel()), float(bound), _p(out),
           int(scales.numel()), _stream())
    return out


def ar_schedule(h, w, wavefront=True):
    """Decoding order of a type-A 5x5 masked context: list of int32 tensors of hw-indexes, one per step.  Position
    (i, j) needs rows i-2, i-1 (columns j-2 .. j+2) and (i, j-2), (i, j-1): time j + 3 i is the earliest consistent
    wavefront; ``wavefront=False`` is compressai's raster order (one position per step)."""
    if not wavefront:
        return [torch.tensor([p], dtype=torch.int32) for p in range(h * w)]
    steps = []
    for t in range(w + 3 * (h - 1)):
        pos = [i * w + (t - 3 * i) for i in range(h) if 0 <= t - 3 * i < w]
        if pos:
            steps.append(torch.tensor(pos, dtype=torch.int32))
    return steps
