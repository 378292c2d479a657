"""Launch programs for a conv/GDN stack (``net.g_a`` / ``net.g_s``): fixed buffers, cached TMA plans.

A *unit* is one Conv2d/ConvTranspose2d optionally followed by GDN/IGDN.  Forward fuses the
normalisation into the contraction's epilogue; backward fuses the normalisation's backward into the
epilogue of the NEXT unit's input-gradient contraction (kernels K1-K5).  Shapes the
tcgen05 path does not take (3-channel end layers) run on the CUDA-core kernels with a stand-alone
(I)GDN launch.  Reference semantics: ``y = net.g_a(x)``, ``x_ = net.g_s(y)`` + autograd to the input
(attack_rd.py:344,349,547).
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops


class UnitSpec:
    """Static description of one unit, extracted from the modules of a stack."""

    def __init__(self, conv, gdn):
        self.transposed = bool(conv.transposed)
        self.k, self.s = int(conv.ksize), int(conv.stride)
        self.cin, self.cout = int(conv.in_channels), int(conv.out_channels)
        self.conv, self.gdn = conv, gdn

    @property
    def fwd_form(self):
        return L.FORM_TCONV if self.transposed else L.FORM_SCONV

    @property
    def bwd_form(self):
        return L.FORM_SCONV if self.transposed else L.FORM_TCONV


def parse_stack(seq):
    """Sequential-like container of our Conv2d/ConvTranspose2d/GDN modules -> list of UnitSpec."""
    from . import models as M
    mods = list(seq._modules.values()) if hasattr(seq, "_modules") else list(seq)
    units, i = [], 0
    while i < len(mods):
        m = mods[i]
        if not isinstance(m, (M.Conv2d, M.ConvTranspose2d)):
            raise L.IcadvError(f"stack program: unsupported module {type(m).__name__} at position {i}")
        gdn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], M.GDN) else None
        units.append(UnitSpec(m, gdn))
        i += 2 if gdn is not None else 1
    return units


def _tc_ok(k_ch, n_ch, gdn):
    return k_ch % 32 == 0 and n_ch % 32 == 0 and 32 <= n_ch <= 256 and (not gdn or 2 * n_ch <= 512)


class StackProgram:
    def __init__(self, units, n_img, in_h, in_w, device, *, x_in=None, g_out=None, g_in=None, active=None,
                 n_active=None, need_grad=True, round_final_out=False, round_final_gin=False):
        self.units, self.n_img, self.device = units, n_img, device
        self.round_final_out, self.round_final_gin = round_final_out, round_final_gin
        self.active, self.n_active = active, n_active
        f = lambda *shape: torch.empty(*shape, device=device, dtype=torch.float32)
        U = len(units)
        self.x_in = x_in if x_in is not None else f(n_img, in_h, in_w, units[0].cin)
        assert tuple(self.x_in.shape) == (n_img, in_h, in_w, units[0].cin), self.x_in.shape
        # ---- shapes
        self.hw = [(in_h, in_w)]
        for u in units:
            self.hw.append(ops.out_hw(u.fwd_form, u.k, u.s, *self.hw[-1]))
        # ---- forward buffers: y[j] = unit output (after GDN if any); sc[j] = GDN scale; u[j] = pre-GDN (unfused only)
        self.y = [f(n_img, *self.hw[j + 1], units[j].cout) for j in range(U)]
        self.sc = [torch.empty_like(self.y[j]) if units[j].gdn is not None else None for j in range(U)]
        self.fused_fwd = [units[j].gdn is not None and _tc_ok(units[j].cin, units[j].cout, True) for j in range(U)]
        self.u = [torch.empty_like(self.y[j]) if (units[j].gdn is not None and not self.fused_fwd[j]) else None
                  for j in range(U)]
        self.out = self.y[-1]
        # ---- parameters in kernel layouts
        self.w_fwd, self.w_bwd, self.bias = [None] * U, [None] * U, [None] * U
        self.beta, self.gamma, self.gammaT = [None] * U, [None] * U, [None] * U
        self.refresh_parameters()
        # ---- backward buffers: gu[j] = gradient wrt the contraction output of unit j
        self.need_grad = need_grad
        if need_grad:
            self.gu = [torch.empty_like(self.y[j]) for j in range(U)]
            if units[-1].gdn is None and g_out is not None:
                assert g_out.shape == self.y[-1].shape
                self.gu[-1] = g_out
            self.g_out = self.gu[-1] if units[-1].gdn is None else torch.empty_like(self.y[-1])
            self.fused_bwd = [j > 0 and units[j - 1].gdn is not None and _tc_ok(units[j].cout, units[j].cin, True)
                              for j in range(U)]
            self.gy = [torch.empty_like(self.y[j - 1]) if (j > 0 and units[j - 1].gdn is not None and
                                                          not self.fused_bwd[j]) else None for j in range(U)]
            self.g_in = g_in if g_in is not None else torch.empty_like(self.x_in)
            assert self.g_in.shape == self.x_in.shape
        self._build()

    # ------------------------------------------------------------------ parameters
    def refresh_parameters(self):
        """(Re)pack weights and reparametrise GDN parameters; call after a codec update."""
        for j, u in enumerate(self.units):
            w = u.conv.weight
            # tensor-path operands are rounded to TF32 (nearest) once, here; CUDA-core layers keep fp32 weights
            wf = ops.pack_weight(w, L.PACK_CONVT_FWD if u.transposed else L.PACK_CONV_FWD,
                                 round_tf32=_tc_ok(u.cin, u.cout, False))
            wb = ops.pack_weight(w, L.PACK_CONVT_DGRAD if u.transposed else L.PACK_CONV_DGRAD,
                                 round_tf32=_tc_ok(u.cout, u.cin, False))
            if self.w_fwd[j] is None:       # plans bake these pointers in: later refreshes copy in place
                self.w_fwd[j], self.w_bwd[j] = wf, wb
            else:
                self.w_fwd[j].copy_(wf)
                self.w_bwd[j].copy_(wb)
            b = u.conv.bias.detach().contiguous().clone() if u.conv.bias is not None else None
            if self.bias[j] is None:
                self.bias[j] = b
            elif b is not None:
                self.bias[j].copy_(b)
            if u.gdn is not None:
                be, ga, gaT = u.gdn.effective_parameters(round_tf32=True)
                for store, val in ((self.beta, be), (self.gamma, ga), (self.gammaT, gaT)):
                    if store[j] is None:
                        store[j] = val
                    else:
                        store[j].copy_(val)

    # ------------------------------------------------------------------ launch lists
    def _launch(self, x, wpack, bias, out, *, form, u, n_ch, **kw):
        d = ops.make_desc(x, wpack, bias, out, form=form, ksize=kw.pop("ksize", u.k), stride=kw.pop("stride", u.s),
                          n_ch=n_ch, active=self.active, n_active=self.n_active, **kw)
        keep = (x, wpack, bias, out) + tuple(v for v in kw.values() if torch.is_tensor(v))
        if L.lib().icadv_conv_tc_supported(C.byref(d)) == 1:
            return ops.ConvPlan(d, keep)
        if kw.get("epi", L.EPI_LINEAR) != L.EPI_LINEAR:
            raise L.IcadvError("normalisation epilogue requested on a shape the tensor path does not take")
        return ops.SimtLaunch(d, keep)

    def _build(self):
        U, units = len(self.units), self.units
        self.fwd, self.bwd = [], []
        x = self.x_in
        # an activation is rounded to TF32 where it is produced iff its consumer is a tensor-path contraction
        fwd_round = [(_tc_ok(units[j + 1].cin, units[j + 1].cout, False) if j + 1 < U else self.round_final_out)
                     for j in range(U)]
        bwd_round = [(_tc_ok(units[j].cout, units[j].cin, False)) for j in range(U)]  # gu[j] feeds dgrad of unit j
        for j, u in enumerate(units):
            if u.gdn is None:
                self.fwd.append(self._launch(x, self.w_fwd[j], self.bias[j], self.y[j], form=u.fwd_form, u=u,
                                             n_ch=u.cout, round_out=fwd_round[j]))
            elif self.fused_fwd[j]:
                epi = L.EPI_IGDN_FWD if u.gdn.inverse else L.EPI_GDN_FWD
                self.fwd.append(self._launch(x, self.w_fwd[j], self.bias[j], self.y[j], form=u.fwd_form, u=u,
                                             n_ch=u.cout, epi=epi, gmat=self.gamma[j], beta=self.beta[j],
                                             out_scale=self.sc[j], round_out=fwd_round[j]))
            else:
                epi = L.EPI_IGDN_FWD if u.gdn.inverse else L.EPI_GDN_FWD
                self.fwd.append(self._launch(x, self.w_fwd[j], self.bias[j], self.u[j], form=u.fwd_form, u=u,
                                             n_ch=u.cout))
                self.fwd.append(self._launch(self.u[j], None, None, self.y[j], form=L.FORM_SCONV, u=u, n_ch=u.cout,
                                             ksize=1, stride=1, epi=epi, gmat=self.gamma[j], beta=self.beta[j],
                                             out_scale=self.sc[j], acc_from_in=True, round_out=fwd_round[j]))
            x = self.y[j]
        if not self.need_grad:
            return
        if units[-1].gdn is not None:
            g = units[-1].gdn
            self.bwd.append(self._launch(self.g_out, None, None, self.gu[-1], form=L.FORM_SCONV, u=units[-1],
                                         n_ch=units[-1].cout, ksize=1, stride=1,
                                         epi=L.EPI_IGDN_BWD if g.inverse else L.EPI_GDN_BWD, gmat=self.gammaT[-1],
                                         y_prev=self.y[-1], sc_prev=self.sc[-1], acc_from_in=True,
                                         round_out=bwd_round[-1]))
        for j in range(U - 1, 0, -1):
            u, prev = units[j], units[j - 1]
            if prev.gdn is None:
                self.bwd.append(self._launch(self.gu[j], self.w_bwd[j], None, self.gu[j - 1], form=u.bwd_form, u=u,
                                             n_ch=u.cin, round_out=bwd_round[j - 1]))
                continue
            epi = L.EPI_IGDN_BWD if prev.gdn.inverse else L.EPI_GDN_BWD
            if self.fused_bwd[j]:
                self.bwd.append(self._launch(self.gu[j], self.w_bwd[j], None, self.gu[j - 1], form=u.bwd_form, u=u,
                                             n_ch=u.cin, epi=epi, gmat=self.gammaT[j - 1], y_prev=self.y[j - 1],
                                             sc_prev=self.sc[j - 1], round_out=bwd_round[j - 1]))
            else:
                self.bwd.append(self._launch(self.gu[j], self.w_bwd[j], None, self.gy[j], form=u.bwd_form, u=u,
                                             n_ch=u.cin))
                self.bwd.append(self._launch(self.gy[j], None, None, self.gu[j - 1], form=L.FORM_SCONV, u=prev,
                                             n_ch=prev.cout, ksize=1, stride=1, epi=epi, gmat=self.gammaT[j - 1],
                                             y_prev=self.y[j - 1], sc_prev=self.sc[j - 1], acc_from_in=True,
                                             round_out=bwd_round[j - 1]))
        u0 = units[0]
        self.bwd.append(self._launch(self.gu[0], self.w_bwd[0], None, self.g_in, form=u0.bwd_form, u=u0, n_ch=u0.cin,
                                     round_out=self.round_final_gin))

    def forward(self):
        for p in self.fwd:
            p.launch()
        return self.out

    def backward(self):
        for p in self.bwd:
            p.launch()
        return self.g_in

    def n_kernels(self):
        """Kernel launches per forward / backward pass."""
        return sum(p.kernels for p in self.fwd), sum(p.kernels for p in self.bwd)
