"""Oracle for SURVEY.md section 8(f) rank 4: actual entropy coding (``compress()`` / ``decompress()``, range-ANS).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Plain Python integers / torch on CPU; small cases only.

The reference never implements the coder itself: its models inherit ``compress`` / ``decompress`` / ``update`` from the
un-vendored, unpinned third-party package ``compressai`` (API era 1.1.x-1.2.x; call sites
``/root/reference/attack_TIC.py:106-110``, ``/root/reference/InvCompress/attack_inv.py:112-116``,
``/root/reference/InvCompress/ours.py:100-175`` shows the model-level shape, ``/root/reference/InvCompress/train.py:452``
calls ``net.update()``).  The package is absent from ``/root/reference`` and from this image, so its published algorithm
is restated here -- **parity unpinned** (no reference-side vector exists; what is checked is self-consistency:
decode(encode(s)) == s, frequencies sum to 2^16, the code length tracks the likelihood estimate):

* ``pmf_to_quantized_cdf``         -- compressai/cpp_exts/ops/ops.cpp
* ``RansEncoder.encode_with_indexes`` / ``RansDecoder.decode_with_indexes`` -- compressai/cpp_exts/rans/rans_interface.cpp on
  ryg_rans ``rans64.h`` (64-bit state, 32-bit renormalisation, 16-bit precision, 4-bit bypass coding of out-of-range
  symbols), symbols pushed in order and coded last-to-first so the decoder reads them first-to-last
* ``EntropyBottleneck.update`` / ``GaussianConditional.update`` / ``build_indexes`` / ``get_scale_table`` --
  compressai/entropy_models/entropy_models.py
* model-level ``compress`` / ``decompress`` of the four families -- compressai/models/google.py (the autoregressive
  pair follows ``_compress_ar`` / ``_decompress_ar``: raster order, position-major / channel-minor symbol order).
"""
import math

import torch
import torch.nn.functional as F

PRECISION = 16
BYPASS_PRECISION = 4
MAX_BYPASS = (1 << BYPASS_PRECISION) - 1
RANS64_L = 1 << 31
M32 = (1 << 32) - 1


def _round_half_away(v):
    return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))


def pmf_to_quantized_cdf(pmf, precision=PRECISION):
    """ops.cpp: scale to 2^precision, renormalise with integer division, cumulate, then give every zero-width symbol one
    count stolen from the least frequent symbol that can spare it."""
    cdf = [0] + [_round_half_away(float(p) * (1 << precision)) for p in pmf]
    total = sum(cdf)
    assert total > 0
    cdf = [((1 << precision) * c) // total for c in cdf]
    for i in range(1, len(cdf)):
        cdf[i] += cdf[i - 1]
    cdf[-1] = 1 << precision
    for i in range(len(cdf) - 1):
        if cdf[i] == cdf[i + 1]:
            best_freq, best_steal = None, -1
            for j in range(len(cdf) - 1):
                freq = cdf[j + 1] - cdf[j]
                if freq > 1 and (best_freq is None or freq < best_freq):
                    best_freq, best_steal = freq, j
            assert best_steal != -1
            if best_steal < i:
                for j in range(best_steal + 1, i + 1):
                    cdf[j] -= 1
            else:
                for j in range(i + 1, best_steal + 1):
                    cdf[j] += 1
    return cdf


def _pmf_to_cdf(pmf, tail_mass, pmf_length, max_length):
    cdf = torch.zeros(len(pmf_length), max_length + 2, dtype=torch.int32)
    for i, p in enumerate(pmf):
        prob = torch.cat((p[: int(pmf_length[i])], tail_mass[i]), dim=0)
        c = pmf_to_quantized_cdf(prob.tolist())
        cdf[i, : len(c)] = torch.tensor(c, dtype=torch.int32)
    return cdf


def eb_tables(eb):
    """EntropyBottleneck.update(): (quantized_cdf [C, L+2], cdf_length [C], offset [C]) as int32 tensors."""
    with torch.no_grad():
        q = eb.quantiles.detach().cpu()
        medians = q[:, 0, 1]
        minima = torch.ceil(medians - q[:, 0, 0]).int().clamp(min=0)
        maxima = torch.ceil(q[:, 0, 2] - medians).int().clamp(min=0)
        offset = -minima
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max())
        samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
        lower = eb._logits_cumulative(samples - 0.5, stop_gradient=True)
        upper = eb._logits_cumulative(samples + 0.5, stop_gradient=True)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        return _pmf_to_cdf(pmf, tail_mass, pmf_length, max_length), (pmf_length + 2).int(), offset.int()


def get_scale_table(lo=0.11, hi=256.0, levels=64):
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


def _std_cum(t):
    return 0.5 * torch.erfc(-(2 ** -0.5) * t)


def gc_tables(scale_table, tail_mass=1e-9):
    """GaussianConditional.update()."""
    # scipy.stats.norm.ppf(tail_mass / 2) without scipy: inverse of Phi by bisection in float64
    lo, hi = -40.0, 0.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if 0.5 * math.erfc(-mid / math.sqrt(2.0)) < tail_mass / 2:
            lo = mid
        else:
            hi = mid
    multiplier = -0.5 * (lo + hi)
    st = scale_table.float()
    pmf_center = torch.ceil(st * multiplier).int()
    pmf_length = 2 * pmf_center + 1
    max_length = int(pmf_length.max())
    samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
    sc = st[:, None]
    upper = _std_cum((0.5 - samples) / sc)
    lower = _std_cum((-0.5 - samples) / sc)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    return _pmf_to_cdf(pmf, tail, pmf_length, max_length), (pmf_length + 2).int(), (-pmf_center).int()


def build_indexes(scales, scale_table, scale_bound=0.11):
    scales = torch.clamp(scales, min=scale_bound)
    idx = torch.full(scales.shape, len(scale_table) - 1, dtype=torch.int32)
    for s in scale_table[:-1]:
        idx -= (scales <= s).int()
    return idx


# ------------------------------------------------------------------------------------------ rans64
def rans_encode_with_indexes(symbols, indexes, cdfs, cdf_sizes, offsets):
    """-> bytes.  All arguments are Python lists of ints (cdfs: list of lists)."""
    syms = []          # (start, range, bypass)
    for s, ci in zip(symbols, indexes):
        cdf = cdfs[ci]
        max_value = cdf_sizes[ci] - 2
        value = s - offsets[ci]
        raw = 0
        if value < 0:
            raw = -2 * value - 1
            value = max_value
        elif value >= max_value:
            raw = 2 * (value - max_value)
            value = max_value
        syms.append((cdf[value], cdf[value + 1] - cdf[value], False))
        if value == max_value:
            n_bypass = 0
            while (raw >> (n_bypass * BYPASS_PRECISION)) != 0:
                n_bypass += 1
            val = n_bypass
            while val >= MAX_BYPASS:
                syms.append((MAX_BYPASS, 1, True))
                val -= MAX_BYPASS
            syms.append((val, 1, True))
            for j in range(n_bypass):
                syms.append(((raw >> (j * BYPASS_PRECISION)) & MAX_BYPASS, 1, True))
    x = RANS64_L
    words = []          # emitted back to front
    for start, rng, bypass in reversed(syms):
        if not bypass:
            x_max = ((RANS64_L >> PRECISION) << 32) * rng
            if x >= x_max:
                words.append(x & M32)
                x >>= 32
            x = ((x // rng) << PRECISION) + (x % rng) + start
        else:
            freq = 1 << (16 - BYPASS_PRECISION)
            x_max = ((RANS64_L >> 16) << 32) * freq
            if x >= x_max:
                words.append(x & M32)
                x >>= 32
            x = (x << BYPASS_PRECISION) | start
    words.append((x >> 32) & M32)
    words.append(x & M32)
    words.reverse()
    return b"".join(int(w).to_bytes(4, "little") for w in words)


def rans_decode_with_indexes(data, indexes, cdfs, cdf_sizes, offsets):
    words = [int.from_bytes(data[i:i + 4], "little") for i in range(0, len(data), 4)]
    x = words[0] | (words[1] << 32)
    pos = 2

    def get_bits(n):
        nonlocal x, pos
        val = x & ((1 << n) - 1)
        x >>= n
        if x < RANS64_L:
            x = (x << 32) | words[pos]
            pos += 1
        return val

    out = []
    for ci in indexes:
        cdf = cdfs[ci]
        size = cdf_sizes[ci]
        max_value = size - 2
        cum = x & ((1 << PRECISION) - 1)
        s = 0
        while s + 1 < size and cdf[s + 1] <= cum:
            s += 1
        start, rng = cdf[s], cdf[s + 1] - cdf[s]
        x = rng * (x >> PRECISION) + (x & ((1 << PRECISION) - 1)) - start
        if x < RANS64_L:
            x = (x << 32) | words[pos]
            pos += 1
        value = s
        if value == max_value:
            val = get_bits(BYPASS_PRECISION)
            n_bypass = val
            while val == MAX_BYPASS:
                val = get_bits(BYPASS_PRECISION)
                n_bypass += val
            raw = 0
            for j in range(n_bypass):
                raw |= get_bits(BYPASS_PRECISION) << (j * BYPASS_PRECISION)
            value = raw >> 1
            if raw & 1:
                value = -value - 1
            else:
                value += max_value
        out.append(value + offsets[ci])
    return out


# ------------------------------------------------------------------------------------------ entropy models
def _lists(tables):
    cdf, length, offset = tables
    return cdf.tolist(), length.tolist(), offset.tolist()


def eb_compress(eb, x, tables=None):
    """EntropyBottleneck.compress: one string per image; symbol order (c, h, w)."""
    tables = tables or eb_tables(eb)
    cdf, length, offset = _lists(tables)
    med = eb.quantiles.detach()[:, 0, 1].view(1, -1, 1, 1)
    sym = torch.round(x - med).int()
    C = x.shape[1]
    idx = torch.arange(C, dtype=torch.int32).view(1, C, 1, 1).expand_as(sym)
    return [rans_encode_with_indexes(sym[i].reshape(-1).tolist(), idx[i].reshape(-1).tolist(), cdf, length, offset)
            for i in range(x.shape[0])]


def eb_decompress(eb, strings, size, tables=None):
    tables = tables or eb_tables(eb)
    cdf, length, offset = _lists(tables)
    C = eb.channels
    med = eb.quantiles.detach()[:, 0, 1].view(1, C, 1, 1)
    idx = torch.arange(C, dtype=torch.int32).view(C, 1, 1).expand(C, size[0], size[1]).reshape(-1).tolist()
    out = [torch.tensor(rans_decode_with_indexes(s, idx, cdf, length, offset), dtype=torch.float32).view(C, *size)
           for s in strings]
    return torch.stack(out) + med


def gc_compress(y, indexes, tables, means=None):
    cdf, length, offset = _lists(tables)
    sym = torch.round(y - means).int() if means is not None else torch.round(y).int()
    return [rans_encode_with_indexes(sym[i].reshape(-1).tolist(), indexes[i].reshape(-1).tolist(), cdf, length, offset)
            for i in range(y.shape[0])]


def gc_decompress(strings, indexes, tables, means=None):
    cdf, length, offset = _lists(tables)
    out = [torch.tensor(rans_decode_with_indexes(s, indexes[i].reshape(-1).tolist(), cdf, length, offset),
                        dtype=torch.float32).view(indexes[i].shape) for i, s in enumerate(strings)]
    out = torch.stack(out)
    return out + means if means is not None else out


# ------------------------------------------------------------------------------------------ models (google.py)
def compress(net, x, family):
    """compressai ``model.compress(x)`` -> {"strings": [...], "shape": ...} for the four families."""
    st = get_scale_table()
    with torch.no_grad():
        y = net.g_a(x)
        if family == "factorized":
            return {"strings": [eb_compress(net.entropy_bottleneck, y)], "shape": tuple(y.shape[-2:])}
        z = net.h_a(torch.abs(y)) if family == "hyper" else net.h_a(y)
        z_strings = eb_compress(net.entropy_bottleneck, z)
        z_hat = eb_decompress(net.entropy_bottleneck, z_strings, tuple(z.shape[-2:]))
        gct = gc_tables(st)
        if family == "hyper":
            scales = net.h_s(z_hat)
            y_strings = gc_compress(y, build_indexes(scales, st), gct)
        elif family == "mean_scale":
            scales, means = net.h_s(z_hat).chunk(2, 1)
            y_strings = gc_compress(y, build_indexes(scales, st), gct, means)
        else:                                       # context / cheng2020: _compress_ar, raster order
            params = net.h_s(z_hat)
            cdf, length, offset = _lists(gct)
            w = net.context_prediction.weight * net.context_prediction.mask
            y_pad = F.pad(y, (2, 2, 2, 2))
            y_strings, y_hats = [], []
            H, W = y.shape[-2:]
            for i in range(y.shape[0]):
                yh = y_pad[i:i + 1].clone()
                syms, idxs = [], []
                for h in range(H):
                    for ww in range(W):
                        crop = yh[:, :, h:h + 5, ww:ww + 5]
                        ctx = F.conv2d(crop, w, net.context_prediction.bias)
                        gp = net.entropy_parameters(torch.cat((params[i:i + 1, :, h:h + 1, ww:ww + 1], ctx), 1))
                        sc, mu = gp.squeeze(3).squeeze(2).chunk(2, 1)
                        idx = build_indexes(sc, st)
                        q = torch.round(crop[:, :, 2, 2] - mu).int()
                        yh[:, :, h + 2, ww + 2] = q + mu
                        syms.extend(q.reshape(-1).tolist())
                        idxs.extend(idx.reshape(-1).tolist())
                y_strings.append(rans_encode_with_indexes(syms, idxs, cdf, length, offset))
                y_hats.append(yh[:, :, 2:-2, 2:-2])
            return {"strings": [y_strings, z_strings], "shape": tuple(z.shape[-2:]), "y_hat": torch.cat(y_hats)}
        return {"strings": [y_strings, z_strings], "shape": tuple(z.shape[-2:])}


def decompress(net, strings, shape, family):
    st = get_scale_table()
    with torch.no_grad():
        if family == "factorized":
            y_hat = eb_decompress(net.entropy_bottleneck, strings[0], shape)
            return {"x_hat": net.g_s(y_hat).clamp_(0, 1), "y_hat": y_hat}
        z_hat = eb_decompress(net.entropy_bottleneck, strings[1], shape)
        gct = gc_tables(st)
        if family == "hyper":
            scales = net.h_s(z_hat)
            y_hat = gc_decompress(strings[0], build_indexes(scales, st), gct)
        elif family == "mean_scale":
            scales, means = net.h_s(z_hat).chunk(2, 1)
            y_hat = gc_decompress(strings[0], build_indexes(scales, st), gct, means)
        else:
            params = net.h_s(z_hat)
            cdf, length, offset = _lists(gct)
            w = net.context_prediction.weight * net.context_prediction.mask
            M = w.shape[1]
            H, W = shape[0] * 4, shape[1] * 4
            outs = []
            for i, s in enumerate(strings[0]):
                words = s
                yh = torch.zeros(1, M, H + 4, W + 4)
                # sequential: each position's parameters need the symbols decoded so far -> decode in one pass with a
                # decoder object that is advanced position by position
                dec = _Stream(words)
                for h in range(H):
                    for ww in range(W):
                        crop = yh[:, :, h:h + 5, ww:ww + 5]
                        ctx = F.conv2d(crop, w, net.context_prediction.bias)
                        gp = net.entropy_parameters(torch.cat((params[i:i + 1, :, h:h + 1, ww:ww + 1], ctx), 1))
                        sc, mu = gp.squeeze(3).squeeze(2).chunk(2, 1)
                        idx = build_indexes(sc, st).reshape(-1).tolist()
                        q = torch.tensor(dec.decode(idx, cdf, length, offset), dtype=torch.float32).view(1, -1)
                        yh[:, :, h + 2, ww + 2] = q + mu
                outs.append(yh[:, :, 2:-2, 2:-2])
            y_hat = torch.cat(outs)
        return {"x_hat": net.g_s(y_hat).clamp_(0, 1), "y_hat": y_hat}


class _Stream:
    """Incremental form of rans_decode_with_indexes (the autoregressive decoder interleaves decoding with the model)."""

    def __init__(self, data):
        self.words = [int.from_bytes(data[i:i + 4], "little") for i in range(0, len(data), 4)]
        self.x = self.words[0] | (self.words[1] << 32)
        self.pos = 2

    def _bits(self, n):
        val = self.x & ((1 << n) - 1)
        self.x >>= n
        if self.x < RANS64_L:
            self.x = (self.x << 32) | self.words[self.pos]
            self.pos += 1
        return val

    def decode(self, indexes, cdfs, cdf_sizes, offsets):
        out = []
        for ci in indexes:
            cdf, size = cdfs[ci], cdf_sizes[ci]
            max_value = size - 2
            cum = self.x & 0xFFFF
            s = 0
            while s + 1 < size and cdf[s + 1] <= cum:
                s += 1
            self.x = (cdf[s + 1] - cdf[s]) * (self.x >> 16) + cum - cdf[s]
            if self.x < RANS64_L:
                self.x = (self.x << 32) | self.words[self.pos]
                self.pos += 1
            value = s
            if value == max_value:
                val = self._bits(4)
                n = val
                while val == MAX_BYPASS:
                    val = self._bits(4)
                    n += val
                raw = 0
                for j in range(n):
                    raw |= self._bits(4) << (j * 4)
                value = raw >> 1
                value = -value - 1 if raw & 1 else value + max_value
            out.append(value + offsets[ci])
        return out
