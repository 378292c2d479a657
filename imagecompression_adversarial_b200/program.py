"""Launch programs for a conv/GDN stack (``net.g_a`` / ``net.g_s``): fixed buffers, cached TMA plans.

A *unit* is one Conv2d/ConvTranspose2d optionally followed by GDN/IGDN.  Forward fuses the
normalisation into the contraction's epilogue; backward fuses the normalisation's backward into the
epilogue of the NEXT unit's input-gradient contraction (kernels K1-K5).  The RGB end
layers run on the tensor path too: 3 -> N strided convs read a padded RGB0 copy of the image through an
overlapping-window tensor map (one K-block per kernel row), N -> 3 transposed convs run as a 1x1 GEMM with
a col2im epilogue.  Anything else the tensor path does not take falls to the CUDA-core kernels with a
stand-alone (I)GDN launch.  Reference semantics: ``y = net.g_a(x)``, ``x_ = net.g_s(y)`` + autograd to
the input (attack_rd.py:344,349,547).
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops
from . import precision


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


def tc_kind(form, k, s, k_ch, n_ch):
    """Which tensor-path mode takes this contraction (mirrors tc_mode() in csrc/icadv_conv_tc.cu):
    "generic", "rgb_in" (3 -> N, 5x5/2, padded RGB0 input), "col2im" (N -> <=4, 5x5/2 transposed) or None."""
    if form == L.FORM_SCONV and k == 5 and s == 2 and k_ch == 3 and n_ch % 32 == 0 and 32 <= n_ch <= 256:
        return "rgb_in"
    if form == L.FORM_TCONV and k == 5 and s == 2 and n_ch <= 4 and k_ch % 32 == 0 and k_ch >= 32:
        return "col2im"
    if k_ch % 32 == 0 and k_ch >= 32 and n_ch % 32 == 0 and 32 <= n_ch <= 256:
        return "generic"
    return None


class PadLaunch:
    """dense RGB [n,h,w,3] -> padded RGB0 layout (input of the rgb_in contraction)."""
    kernels = 1

    def __init__(self, src, dst, active, n_active):
        self.src, self.dst, self.active, self.n_active = src, dst, active, n_active

    def launch(self):
        ops.pad_rgb4(self.src, self.dst, self.active, self.n_active)


class FnLaunch:
    """One elementwise library call with its arguments bound (same interface as ConvPlan)."""
    kernels = 1

    def __init__(self, fn, *args):
        self.fn, self.args = fn, args

    def launch(self):
        self.fn(*self.args)


class StackProgram:
    def __init__(self, units, n_img, in_h, in_w, device, *, x_in=None, g_out=None, g_in=None, active=None,
                 n_active=None, need_grad=True, round_final_out=False, round_final_gin=False, name="stack"):
        self.units, self.n_img, self.device = units, n_img, device
        self.name = name
        self.fwd_info, self.bwd_info = [], []
        self.split = precision.split()
        if self.split:
            self._init_split(in_h, in_w, x_in, g_out, g_in, active, n_active, need_grad)
            return
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
        # ---- which path takes each contraction
        self.fwd_kind = [tc_kind(u.fwd_form, u.k, u.s, u.cin, u.cout) for u in units]
        self.bwd_kind = [tc_kind(u.bwd_form, u.k, u.s, u.cout, u.cin) for u in units]
        for j in range(U):   # the padded-RGB form needs even image sizes
            if self.fwd_kind[j] == "rgb_in" and (self.hw[j][0] % 2 or self.hw[j][1] % 2):
                self.fwd_kind[j] = None
            if self.bwd_kind[j] == "rgb_in" and (self.hw[j + 1][0] % 2 or self.hw[j + 1][1] % 2):
                self.bwd_kind[j] = None
        gdn_ok = lambda j: units[j].gdn is not None and 2 * units[j].cout <= 512
        # ---- forward buffers: y[j] = unit output (after GDN if any); sc[j] = GDN scale; u[j] = pre-GDN (unfused only)
        if need_grad:
            self.y = [f(n_img, *self.hw[j + 1], units[j].cout) for j in range(U)]
            self.sc = [torch.empty_like(self.y[j]) if units[j].gdn is not None else None for j in range(U)]
        else:
            # inference-only program (clean pass, final eval): nothing is saved for a backward pass, so the unit outputs
            # ping-pong between two flat buffers and every scale goes to one scratch buffer
            shapes = [(n_img, *self.hw[j + 1], units[j].cout) for j in range(U)]
            numel = [s[0] * s[1] * s[2] * s[3] for s in shapes]
            flat = [f(max(numel[0::2])), f(max(numel[1::2]) if U > 1 else 1), f(max(numel))]
            self.y = [flat[j % 2][:numel[j]].view(shapes[j]) for j in range(U)]
            self.sc = [flat[2][:numel[j]].view(shapes[j]) if units[j].gdn is not None else None for j in range(U)]
        self.fused_fwd = [gdn_ok(j) and self.fwd_kind[j] in ("generic", "rgb_in") for j in range(U)]
        self.u = [torch.empty_like(self.y[j]) if (units[j].gdn is not None and not self.fused_fwd[j]) else None
                  for j in range(U)]
        self.out = self.y[-1]
        self.x_pad = {j: ops.alloc_pad4(n_img, *self.hw[j], device) for j in range(U) if self.fwd_kind[j] == "rgb_in"}
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
            # backward of GDN_{j-1} rides in the epilogue of unit j's input-gradient contraction
            self.fused_bwd = [j > 0 and gdn_ok(j - 1) and self.bwd_kind[j] in ("generic", "rgb_in")
                              for j in range(U)]
            self.gy = [torch.empty_like(self.y[j - 1]) if (j > 0 and units[j - 1].gdn is not None and
                                                          not self.fused_bwd[j]) else None for j in range(U)]
            self.g_in = g_in if g_in is not None else torch.empty_like(self.x_in)
            assert self.g_in.shape == self.x_in.shape
            self.gu_pad = {j: ops.alloc_pad4(n_img, *self.hw[j + 1], device) for j in range(U)
                           if self.bwd_kind[j] == "rgb_in"}
        self._build()

    # ------------------------------------------------------------------ 3xTF32 parity mode
    def _init_split(self, in_h, in_w, x_in, g_out, g_in, active, n_active, need_grad):
        """Launch lists of the parity mode (precision.py, csrc/icadv_split.cu): every contraction input is expanded to
        the three-term split form [hi | lo | hi] along the channels, cut into K-slices, and contracted slice by slice
        with [Whi | Whi | Wlo] weights on the same tensor-path kernel (linear epilogue); the partial outputs are added in
        fp32.  GDN / IGDN are unfused (square -> split -> 1x1 contraction with gamma -> apply; backward: operand ->
        split -> 1x1 with gamma^T -> combine).  Nothing is rounded to TF32 in memory.  Scratch buffers are shared by all
        units (everything is stream-ordered)."""
        units, n_img, device = self.units, self.n_img, self.device
        self.round_final_out = self.round_final_gin = False
        self.active, self.n_active = active, n_active
        f = lambda *shape: torch.empty(*shape, device=device, dtype=torch.float32)
        U = len(units)
        self.x_in = x_in if x_in is not None else f(n_img, in_h, in_w, units[0].cin)
        assert tuple(self.x_in.shape) == (n_img, in_h, in_w, units[0].cin), self.x_in.shape
        self.hw = [(in_h, in_w)]
        for u in units:
            self.hw.append(ops.out_hw(u.fwd_form, u.k, u.s, *self.hw[-1]))
        self.fwd_kind = self.bwd_kind = ["split"] * U
        self.need_grad = need_grad
        self.y = [f(n_img, *self.hw[j + 1], units[j].cout) for j in range(U)]
        self.sc = [torch.empty_like(self.y[j]) if units[j].gdn is not None else None for j in range(U)]
        self.out = self.y[-1]
        px = lambda j: n_img * self.hw[j][0] * self.hw[j][1]       # pixels of the INPUT of unit j (px(j+1): its output)
        width = lambda geom: geom[0] * geom[1]
        self.geom_f = [ops.split_geom(u.cin, u.fwd_form, u.k, u.s) for u in units]       # forward contraction of unit j
        self.geom_b = [ops.split_geom(u.cout, u.bwd_form, u.k, u.s) for u in units]      # its input gradient
        self.geom_n = [ops.split_geom(u.cout) if u.gdn is not None else None for u in units]   # 1x1 normalisation GEMM
        n_split, n_parts = 0, 0
        for j, u in enumerate(units):
            n_split = max(n_split, px(j) * width(self.geom_f[j]), px(j + 1) * width(self.geom_b[j]),
                          px(j + 1) * width(self.geom_n[j]) if u.gdn is not None else 0)
            n_parts = max(n_parts, self.geom_f[j][1] * px(j + 1) * u.cout, self.geom_b[j][1] * px(j) * u.cin,
                          self.geom_n[j][1] * px(j + 1) * u.cout if u.gdn is not None else 0)
        n_act = max([self.y[j].numel() for j in range(U)] + [self.x_in.numel()])
        self._s = f(n_split)                                      # split operand of the next contraction
        self._p = f(n_parts)                                      # partial outputs of the K-slices
        self._a, self._b, self._c = f(n_act), f(n_act), f(n_act)  # pre-GDN output / gradient ping-pong / norm
        self.w_fwd, self.w_bwd, self.bias = [None] * U, [None] * U, [None] * U
        self.beta, self.gamma, self.gammaT = [None] * U, [None] * U, [None] * U
        self.refresh_parameters()
        if need_grad:
            if g_out is not None:
                assert g_out.shape == self.y[-1].shape
            self.g_out = g_out if g_out is not None else torch.empty_like(self.y[-1])
            self.gu = [None] * (U - 1) + [self.g_out]
            self.g_in = g_in if g_in is not None else torch.empty_like(self.x_in)
            assert self.g_in.shape == self.x_in.shape
        self.fwd, self.bwd = [], []
        aview = lambda buf, like: buf[:like.numel()].view(like.shape)

        def sview(geom, j_px, hw):
            ks, g = geom
            return self._s[:g * j_px * ks].view(g, n_img, hw[0], hw[1], ks)

        def contraction(lst, xs, ws, bias, dst, form, u, n_ch, **kw):
            """G slice launches + the fp32 sum of their partial outputs into dst."""
            g = xs.shape[0]
            if g == 1:
                lst.append(self._launch(xs[0], ws[0], bias, dst, form=form, u=u, n_ch=n_ch, **kw))
                return
            parts = self._p[:g * dst.numel()].view(g, *dst.shape)
            for i in range(g):
                lst.append(self._launch(xs[i], ws[i], bias if i == 0 else None, parts[i], form=form, u=u, n_ch=n_ch, **kw))
            lst.append(FnLaunch(ops.sum_slices, parts, dst))

        x = self.x_in
        for j, u in enumerate(units):
            xs = sview(self.geom_f[j], px(j), self.hw[j])
            self.fwd.append(FnLaunch(ops.split3, x, *self.geom_f[j], 0, xs))
            dst = self.y[j] if u.gdn is None else aview(self._a, self.y[j])
            contraction(self.fwd, xs, self.w_fwd[j], self.bias[j], dst, u.fwd_form, u, u.cout)
            if u.gdn is not None:
                sq = sview(self.geom_n[j], px(j + 1), self.hw[j + 1])
                nrm = aview(self._c, self.y[j])
                self.fwd.append(FnLaunch(ops.split3, dst, *self.geom_n[j], 1, sq))
                contraction(self.fwd, sq, self.gamma[j], self.beta[j], nrm, L.FORM_SCONV, u, u.cout, ksize=1, stride=1)
                self.fwd.append(FnLaunch(ops.gdn_apply, dst, nrm, self.y[j], self.sc[j], u.gdn.inverse))
            x = self.y[j]
        if not need_grad:
            return
        gcur = self.g_out                                          # gradient wrt the OUTPUT y[j] of the current unit
        for j in range(U - 1, -1, -1):
            u = units[j]
            gsrc = gcur
            if u.gdn is not None:
                t = sview(self.geom_n[j], px(j + 1), self.hw[j + 1])
                w = aview(self._c, self.y[j])
                gsrc = aview(self._a, self.y[j])
                self.bwd.append(FnLaunch(ops.gdn_bwd_operand_split3, gcur, self.y[j], self.sc[j], u.gdn.inverse,
                                         *self.geom_n[j], t))
                contraction(self.bwd, t, self.gammaT[j], None, w, L.FORM_SCONV, u, u.cout, ksize=1, stride=1)
                self.bwd.append(FnLaunch(ops.gdn_bwd_combine, gcur, self.y[j], self.sc[j], w, gsrc, u.gdn.inverse))
            gs = sview(self.geom_b[j], px(j + 1), self.hw[j + 1])
            self.bwd.append(FnLaunch(ops.split3, gsrc, *self.geom_b[j], 0, gs))
            dst = self.g_in if j == 0 else aview(self._b, self.y[j - 1])
            contraction(self.bwd, gs, self.w_bwd[j], None, dst, u.bwd_form, u, u.cin)
            gcur = dst

    def _refresh_split(self):
        for j, u in enumerate(self.units):
            w = u.conv.weight
            wf = ops.split3_weight(ops.pack_weight(w, L.PACK_CONVT_FWD if u.transposed else L.PACK_CONV_FWD), *self.geom_f[j])
            wb = ops.split3_weight(ops.pack_weight(w, L.PACK_CONVT_DGRAD if u.transposed else L.PACK_CONV_DGRAD),
                                   *self.geom_b[j])
            vals = [(self.w_fwd, wf), (self.w_bwd, wb)]
            if u.conv.bias is not None:
                vals.append((self.bias, u.conv.bias.detach().contiguous().clone()))
            if u.gdn is not None:
                be, ga, gaT = u.gdn.effective_parameters(round_tf32=False)
                c = u.cout
                vals += [(self.beta, be), (self.gamma, ops.split3_weight(ga.contiguous().view(1, c, c), *self.geom_n[j])),
                         (self.gammaT, ops.split3_weight(gaT.contiguous().view(1, c, c), *self.geom_n[j]))]
            for store, val in vals:
                if store[j] is None:      # plans bake these pointers in: later refreshes copy in place
                    store[j] = val
                else:
                    store[j].copy_(val)

    # ------------------------------------------------------------------ parameters
    def refresh_parameters(self):
        """(Re)pack weights and reparametrise GDN parameters; call after a codec update.
        Tensor-path operands are rounded to TF32 (nearest) once, here; CUDA-core layers keep fp32 weights."""
        if self.split:
            return self._refresh_split()
        for j, u in enumerate(self.units):
            w = u.conv.weight
            if self.fwd_kind[j] == "rgb_in":
                wf = ops.pack_weight_rgb(w, round_tf32=True)
            else:
                wf = ops.pack_weight(w, L.PACK_CONVT_FWD if u.transposed else L.PACK_CONV_FWD,
                                     round_tf32=self.fwd_kind[j] is not None)
            if self.bwd_kind[j] == "rgb_in":
                wb = ops.pack_weight_rgb(w, round_tf32=True)
            else:
                wb = ops.pack_weight(w, L.PACK_CONVT_DGRAD if u.transposed else L.PACK_CONV_DGRAD,
                                     round_tf32=self.bwd_kind[j] is not None)
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

    def _gdn_alone(self, x, out, j, epi, **kw):
        u = self.units[j]
        return self._launch(x, None, None, out, form=L.FORM_SCONV, u=u, n_ch=u.cout, ksize=1, stride=1, epi=epi,
                            acc_from_in=True, **kw)

    # ------------------------------------------------------------------ algorithmic work per launch (roofline tables)
    def _conv_macs(self, j):
        u = self.units[j]
        px = self.hw[j][0] * self.hw[j][1] if u.transposed else self.hw[j + 1][0] * self.hw[j + 1][1]
        return self.n_img * px * u.k * u.k * u.cin * u.cout

    def _act_bytes(self, j):
        """fp32 bytes of the OUTPUT activation of unit j (j = -1: the stack input)."""
        c = self.units[0].cin if j < 0 else self.units[j].cout
        return 4 * self.n_img * self.hw[j + 1][0] * self.hw[j + 1][1] * c

    def _note(self, lst, name, macs, nbytes, bound):
        """Record the algorithmic work of the launch just appended to ``lst`` (SURVEY.md section 8d accounting:
        every tensor read once and written once; weights stay in L2)."""
        info = self.fwd_info if lst is self.fwd else self.bwd_info
        while len(info) < len(lst) - 1:
            info.append(None)
        info.append({"name": name, "flops": 2.0 * macs, "bytes": float(nbytes), "bound": bound})

    def _build(self):
        U, units = len(self.units), self.units
        self.fwd, self.bwd = [], []
        self.fwd_info, self.bwd_info = [], []
        nm = lambda j, d: f"{self.name}.{j * 2} {'deconv' if units[j].transposed else 'conv'} {units[j].cin}->{units[j].cout} {d}"
        end = lambda kind: "hbm" if kind in ("rgb_in", "col2im", None) else "tensor"
        gmacs = lambda j: self.n_img * self.hw[j + 1][0] * self.hw[j + 1][1] * units[j].cout ** 2
        # an activation is rounded to TF32 where it is produced iff its consumer is a tensor-path contraction
        fwd_round = [(self.fwd_kind[j + 1] is not None if j + 1 < U else self.round_final_out) for j in range(U)]
        bwd_round = [self.bwd_kind[j] is not None for j in range(U)]   # gu[j] feeds the input-gradient of unit j
        x = self.x_in
        for j, u in enumerate(units):
            extra = {}
            in_bytes = self._act_bytes(j - 1)
            if self.fwd_kind[j] == "rgb_in":
                self.fwd.append(PadLaunch(x, self.x_pad[j], self.active, self.n_active))
                in_bytes = 4 * self.x_pad[j].numel()
                self._note(self.fwd, f"{self.name}.{j * 2} pad RGB->RGB0", 0, self._act_bytes(j - 1) + in_bytes, "hbm")
                x, extra = self.x_pad[j], {"in_pad4": True}
            if u.gdn is None:
                self.fwd.append(self._launch(x, self.w_fwd[j], self.bias[j], self.y[j], form=u.fwd_form, u=u,
                                             n_ch=u.cout, round_out=fwd_round[j], **extra))
                self._note(self.fwd, nm(j, "fwd"), self._conv_macs(j), in_bytes + self._act_bytes(j), end(self.fwd_kind[j]))
            else:
                epi = L.EPI_IGDN_FWD if u.gdn.inverse else L.EPI_GDN_FWD
                if self.fused_fwd[j]:
                    self.fwd.append(self._launch(x, self.w_fwd[j], self.bias[j], self.y[j], form=u.fwd_form, u=u,
                                                 n_ch=u.cout, epi=epi, gmat=self.gamma[j], beta=self.beta[j],
                                                 out_scale=self.sc[j], round_out=fwd_round[j], **extra))
                    self._note(self.fwd, nm(j, "fwd + " + ("IGDN" if u.gdn.inverse else "GDN")),
                               self._conv_macs(j) + gmacs(j), in_bytes + 2 * self._act_bytes(j), end(self.fwd_kind[j]))
                else:
                    self.fwd.append(self._launch(x, self.w_fwd[j], self.bias[j], self.u[j], form=u.fwd_form, u=u,
                                                 n_ch=u.cout, **extra))
                    self._note(self.fwd, nm(j, "fwd"), self._conv_macs(j), in_bytes + self._act_bytes(j), end(self.fwd_kind[j]))
                    self.fwd.append(self._gdn_alone(self.u[j], self.y[j], j, epi, gmat=self.gamma[j],
                                                    beta=self.beta[j], out_scale=self.sc[j],
                                                    round_out=fwd_round[j]))
                    self._note(self.fwd, f"{self.name}.{j * 2 + 1} (I)GDN fwd", gmacs(j), 3 * self._act_bytes(j), "hbm")
            x = self.y[j]
        if not self.need_grad:
            return
        if units[-1].gdn is not None:
            g = units[-1].gdn
            self.bwd.append(self._gdn_alone(self.g_out, self.gu[-1], U - 1,
                                            L.EPI_IGDN_BWD if g.inverse else L.EPI_GDN_BWD, gmat=self.gammaT[-1],
                                            y_prev=self.y[-1], sc_prev=self.sc[-1], round_out=bwd_round[-1]))
            self._note(self.bwd, f"{self.name}.{2 * U - 1} (I)GDN bwd", gmacs(U - 1), 4 * self._act_bytes(U - 1), "hbm")
        for j in range(U - 1, -1, -1):
            u = units[j]
            gsrc, extra = self.gu[j], {}
            g_bytes = self._act_bytes(j)
            if self.bwd_kind[j] == "rgb_in":
                self.bwd.append(PadLaunch(self.gu[j], self.gu_pad[j], self.active, self.n_active))
                g_bytes = 4 * self.gu_pad[j].numel()
                self._note(self.bwd, f"{self.name}.{j * 2} pad RGB->RGB0 (gradient)", 0, self._act_bytes(j) + g_bytes, "hbm")
                gsrc, extra = self.gu_pad[j], {"in_pad4": True}
            if j == 0:
                self.bwd.append(self._launch(gsrc, self.w_bwd[0], None, self.g_in, form=u.bwd_form, u=u, n_ch=u.cin,
                                             round_out=self.round_final_gin, **extra))
                self._note(self.bwd, nm(0, "dgrad"), self._conv_macs(0), g_bytes + self._act_bytes(-1), end(self.bwd_kind[0]))
                continue
            prev = units[j - 1]
            if prev.gdn is None:
                self.bwd.append(self._launch(gsrc, self.w_bwd[j], None, self.gu[j - 1], form=u.bwd_form, u=u,
                                             n_ch=u.cin, round_out=bwd_round[j - 1], **extra))
                self._note(self.bwd, nm(j, "dgrad"), self._conv_macs(j), g_bytes + self._act_bytes(j - 1), end(self.bwd_kind[j]))
                continue
            epi = L.EPI_IGDN_BWD if prev.gdn.inverse else L.EPI_GDN_BWD
            if self.fused_bwd[j]:
                self.bwd.append(self._launch(gsrc, self.w_bwd[j], None, self.gu[j - 1], form=u.bwd_form, u=u,
                                             n_ch=u.cin, epi=epi, gmat=self.gammaT[j - 1], y_prev=self.y[j - 1],
                                             sc_prev=self.sc[j - 1], round_out=bwd_round[j - 1], **extra))
                self._note(self.bwd, nm(j, "dgrad + " + ("IGDN" if prev.gdn.inverse else "GDN") + " bwd"),
                           self._conv_macs(j) + gmacs(j - 1), g_bytes + 3 * self._act_bytes(j - 1), end(self.bwd_kind[j]))
            else:
                self.bwd.append(self._launch(gsrc, self.w_bwd[j], None, self.gy[j], form=u.bwd_form, u=u,
                                             n_ch=u.cin, **extra))
                self._note(self.bwd, nm(j, "dgrad"), self._conv_macs(j), g_bytes + self._act_bytes(j - 1), end(self.bwd_kind[j]))
                self.bwd.append(self._gdn_alone(self.gy[j], self.gu[j - 1], j - 1, epi, gmat=self.gammaT[j - 1],
                                                y_prev=self.y[j - 1], sc_prev=self.sc[j - 1],
                                                round_out=bwd_round[j - 1]))
                self._note(self.bwd, f"{self.name}.{j * 2 - 1} (I)GDN bwd", gmacs(j - 1), 4 * self._act_bytes(j - 1), "hbm")

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
