"""Static launch programs for codecs whose ``g_a`` / ``g_s`` are not plain conv/GDN stacks (cheng2020_anchor: residual
blocks, sub-pixel convolutions, skip connections; anchors/model.py:77).

The module graph is TRACED once (one forward through the operator surface with a recorder in ``functional``), then
compiled into two launch lists -- forward, and backward to the input -- over pre-allocated buffers with cached TMA
plans.  What the trace buys over walking the modules through autograd every iteration (the previous
``GenericAttackEngine``): no per-call plan creation / destruction, no autograd bookkeeping or allocation, the whole
attack iteration is capturable in ONE CUDA graph, and three peephole fusions:

  * LeakyReLU / ReLU following a contraction (directly, or behind a PixelShuffle: an activation commutes with a
    permutation) runs in that contraction's epilogue;
  * an activation that feeds only tensor-path contractions is rounded to TF32 where it is produced (``round_out``) instead
    of by a separate rounding pass in front of every contraction;
  * the gradient of a tensor with several consumers (block input: first conv + skip) is summed by one add launch.

Weight gradients are not produced (the attack never reads them: attack_rd.py:546-548).  Reference semantics:
``x_ = net.g_s(net.g_a(im_in))`` and autograd to ``im_in`` (attack_rd.py:344-349,547).
"""
import torch

from . import _lib as L
from . import functional as Fn
from . import ops
from .ops import _p, _stream
from .program import FnLaunch

_UN = {L.ACT_ABS: 0, L.ACT_RELU: 1, L.ACT_LEAKY: 2}


class Recorder:
    """Receives one ``note()`` per operator-surface call while a trace is active (``functional._REC``).  Tensors are
    identified by the allocation behind them (channels-last buffers; NCHW views of one buffer share its id)."""

    def __init__(self):
        self.nodes, self.keep, self.ids = [], [], {}
        self.current_gdn = None

    def tid(self, t_nhwc):
        key = (t_nhwc.data_ptr(), tuple(t_nhwc.shape))
        if key not in self.ids:
            self.ids[key] = len(self.ids)
            self.keep.append(t_nhwc)     # keep the allocation alive so data_ptr stays unique for the whole trace
        return self.ids[key]

    def note(self, kind, inputs, output, **attrs):
        """``inputs`` / ``output``: channels-last [N,H,W,C] tensors (views are fine)."""
        self.nodes.append({"kind": kind, "in": [self.tid(t) for t in inputs], "out": self.tid(output),
                           "shape": tuple(output.shape), **attrs})


def trace(stack, x_nchw):
    """Run ``stack(x)`` once under the recorder; returns (nodes, id of the input buffer, id of the output buffer)."""
    rec = Recorder()
    x = x_nchw.contiguous(memory_format=torch.channels_last)
    in_id = rec.tid(Fn.to_nhwc(x))
    Fn._REC = rec
    try:
        with torch.no_grad():
            out = stack(x)
    finally:
        Fn._REC = None
    return rec.nodes, in_id, rec.tid(Fn.to_nhwc(out))


def _unary(x, b, y, op):
    L.call("icadv_unary", _p(x), _p(b), _p(y), x.numel(), int(op), _stream())


def _act_bwd(x, g, gx, act, rnd=False):
    L.call("icadv_act_backward", _p(x), _p(g), _p(gx), x.numel(), _UN[act] | (256 if rnd else 0), _stream())


def _gate(a, b, x, y):
    L.call("icadv_attention_gate", _p(a), _p(b), _p(x), _p(y), a.numel(), _stream())


def _gate_bwd(a, b, g, ga, gb):
    L.call("icadv_attention_gate_backward", _p(a), _p(b), _p(g), _p(ga), _p(gb), a.numel(), _stream())


def _copy_ch(src, dst, count):
    """dst[..., :count] = src[..., :count] (channels-last tensors of any channel counts)."""
    L.call("icadv_copy_channels", _p(src), _p(dst), src.numel() // src.shape[-1], src.shape[-1], dst.shape[-1], 0, 0, count,
           _stream())


def _pad32(c):
    return (c + 31) // 32 * 32


def _shuffle(src, dst, n, lo_h, lo_w, c_out, r, inverse):
    L.call("icadv_pixel_shuffle", _p(src), _p(dst), n, lo_h, lo_w, c_out, r, 1 if inverse else 0, _stream())


class TapeProgram:
    """Forward + input-gradient launch lists of one traced stack; same interface as ``program.StackProgram``."""

    def __init__(self, stack, n_img, in_h, in_w, device, *, x_in=None, g_out=None, g_in=None, active=None, n_active=None,
                 name="stack"):
        self.name, self.n_img, self.device = name, n_img, device
        self.active, self.n_active = active, n_active
        f = lambda *shape: torch.empty(*shape, device=device, dtype=torch.float32)
        cin = self._first_in_channels(stack)
        self.x_in = x_in if x_in is not None else f(n_img, in_h, in_w, cin)
        probe = torch.zeros(n_img, cin, in_h, in_w, device=device).contiguous(memory_format=torch.channels_last)
        nodes, in_id, out_id = trace(stack, probe)
        del probe
        self.nodes, alias = self._fuse_activations(nodes)
        out_id = alias.get(out_id, out_id)
        # ---- buffers: one per node output; the input buffer is the caller's
        self.buf = {in_id: self.x_in}
        for nd in self.nodes:
            self.buf[nd["out"]] = f(*nd["shape"])
        self.out = self.buf[out_id]
        self._out_id = out_id
        self.consumers = {}
        for i, nd in enumerate(self.nodes):
            for t in nd["in"]:
                self.consumers.setdefault(t, []).append(i)
        self.producer = {nd["out"]: i for i, nd in enumerate(self.nodes)}
        self._keep = []
        self.fwd, self.bwd = [], []
        self.fwd_info, self.bwd_info = [], []
        self._weights = []           # (module weight, kind, packed buffer) for refresh_parameters
        self._gdn = []               # (gdn module, beta buffer, gamma buffer, gammaT buffer)
        self._build_forward(in_id)
        self.g_out = g_out if g_out is not None else torch.empty_like(self.out)
        assert self.g_out.shape == self.out.shape
        self.g_in = g_in if g_in is not None else torch.empty_like(self.x_in)
        self._build_backward(in_id, out_id)
        self._describe()

    @staticmethod
    def _first_in_channels(stack):
        for m in stack.modules():
            if hasattr(m, "in_channels"):
                return int(m.in_channels)
        raise L.IcadvError("tape program: no contraction in the stack")

    # ------------------------------------------------------------------ peephole: activations into epilogues
    @staticmethod
    def _fuse_activations(nodes):
        """conv -> act and conv -> shuffle -> act (each link the only consumer) become conv(act) [-> shuffle]."""
        uses = {}
        for nd in nodes:
            for t in nd["in"]:
                uses[t] = uses.get(t, 0) + 1
        prod = {nd["out"]: nd for nd in nodes}
        drop, alias = set(), {}
        for nd in nodes:
            if nd["kind"] != "act":
                continue
            src = prod.get(nd["in"][0])
            via = None
            if src is not None and src["kind"] == "shuffle" and uses[src["out"]] == 1:
                via, src = src, prod.get(src["in"][0])
            if src is None or src["kind"] != "conv" or src["act"] != L.ACT_NONE or uses[src["out"]] != 1:
                continue
            src["act"] = nd["act"]
            drop.add(id(nd))
            alias[nd["out"]] = via["out"] if via is not None else src["out"]
        out = []
        for nd in nodes:
            if id(nd) in drop:
                continue
            nd["in"] = [alias.get(t, t) for t in nd["in"]]
            out.append(nd)
        return out, alias

    # ------------------------------------------------------------------ helpers
    def _tc(self, k_ch, n_ch):
        return Fn.tc_shape(k_ch, n_ch)

    def _plan(self, x, w, bias, out, **kw):
        d = ops.make_desc(x, w, bias, out, active=self.active, n_active=self.n_active, **kw)
        keep = (x, w, bias, out) + tuple(v for v in kw.values() if torch.is_tensor(v))
        import ctypes as C
        if L.lib().icadv_conv_tc_supported(C.byref(d)) == 1:
            return ops.ConvPlan(d, keep)
        if kw.get("epi", L.EPI_LINEAR) != L.EPI_LINEAR or kw.get("acc_from_in"):
            raise L.IcadvError("tape program: normalisation on a shape the tensor path does not take")
        return ops.SimtLaunch(d, keep)

    def _rounded_source(self, lst, src, rounded_cache):
        """Buffer holding ``src`` rounded to TF32 for a tensor-path consumer: ``src`` itself when its producer rounds on
        store, else one rounding launch into a scratch copy (shared by all consumers)."""
        if src in self._round_at_producer:
            return self.buf[src]
        if src not in rounded_cache:
            r = torch.empty_like(self.buf[src])
            lst.append(FnLaunch(_unary, self.buf[src], None, r, 5))
            rounded_cache[src] = r
        return rounded_cache[src]

    def _pack(self, weight, kind, tc, pad_to=None):
        """Packed weight [taps][n][k]; ``pad_to = (n32, k32)``: zero-padded to channel counts the tensor path takes."""
        pk = ops.pack_weight(weight, kind, round_tf32=tc)
        if pad_to is None:
            self._weights.append((weight, kind, tc, pk, None))
            return pk
        w = torch.zeros(pk.shape[0], pad_to[0], pad_to[1], device=pk.device, dtype=torch.float32)
        w[:, :pk.shape[1], :pk.shape[2]].copy_(pk)
        self._weights.append((weight, kind, tc, w, (pk.shape[1], pk.shape[2])))
        return w

    def _padded(self, k_ch, n_ch):
        """Channel counts the tensor path does not take as they are (RGB ends, the 12-channel sub-pixel conv) but does
        once zero-padded to multiples of 32: the padded contraction costs a few MB of extra traffic; the CUDA-core
        kernel it replaces ran at 1 - 5 % of its HBM bound (profiles/r2_bench_config4.json: 8.3 ms for conv 192 -> 12)."""
        return not self._tc(k_ch, n_ch) and self._tc(_pad32(k_ch), _pad32(n_ch))

    def _padded_input(self, lst, x, k_ch):
        """[.., k_ch] -> zero-padded, TF32-rounded [.., pad32(k_ch)] copy (the padding channels are written once, here)."""
        xp = torch.zeros(*x.shape[:-1], _pad32(k_ch), device=x.device, dtype=torch.float32)
        lst.append(FnLaunch(_copy_ch, x, xp, k_ch))
        lst.append(FnLaunch(_unary, xp, None, xp, 5))
        return xp

    def refresh_parameters(self):
        """Re-pack weights / re-parametrise GDN in place after a codec update (plans bake the pointers in)."""
        for weight, kind, tc, w, sub in self._weights:
            pk = ops.pack_weight(weight, kind, round_tf32=tc)
            (w if sub is None else w[:, :sub[0], :sub[1]]).copy_(pk)
        for nd in self.nodes:
            if nd["kind"] == "conv" and nd.get("bias_buf") is not None:
                nd["bias_buf"][:nd["bias"].numel()].copy_(nd["bias"].detach())
        for gdn, be, ga, gaT in self._gdn:
            b2, g2, gT2 = gdn.effective_parameters(round_tf32=True)
            be.copy_(b2); ga.copy_(g2); gaT.copy_(gT2)

    # ------------------------------------------------------------------ forward
    def _build_forward(self, in_id):
        # a tensor is rounded where it is produced iff its producer is a contraction / GDN launch and EVERY consumer is a
        # tensor-path contraction
        self._round_at_producer = set()
        for t, cons in self.consumers.items():
            pi = self.producer.get(t)
            if pi is None or self.nodes[pi]["kind"] not in ("conv", "gdn"):
                continue
            if all(self.nodes[c]["kind"] == "conv" and self._tc(self.buf[t].shape[-1], self.nodes[c]["n_ch"]) for c in cons):
                self._round_at_producer.add(t)
        # (a shuffle is a permutation: rounding its source conv covers its consumers)
        for nd in self.nodes:
            if nd["kind"] == "shuffle":
                t = nd["out"]
                cons = self.consumers.get(t, [])
                src = nd["in"][0]
                if cons and self.producer.get(src) is not None and self.nodes[self.producer[src]]["kind"] == "conv" and \
                        len(self.consumers.get(src, [])) == 1 and \
                        all(self.nodes[c]["kind"] == "conv" and self._tc(self.buf[t].shape[-1], self.nodes[c]["n_ch"]) for c in cons):
                    self._round_at_producer.add(src)
                    self._round_at_producer.add(t)
        rounded = {}
        for nd in self.nodes:
            out = self.buf[nd["out"]]
            if nd["kind"] == "conv":
                src = nd["in"][0]
                k_ch, n_ch = self.buf[src].shape[-1], nd["n_ch"]
                tc = self._tc(k_ch, n_ch)
                kind = L.PACK_CONVT_FWD if nd["transposed"] else L.PACK_CONV_FWD
                if self._padded(k_ch, n_ch):
                    kp, npad = _pad32(k_ch), _pad32(n_ch)
                    if kp != k_ch:
                        x = self._padded_input(self.fwd, self.buf[src], k_ch)
                    else:
                        x = self._rounded_source(self.fwd, src, rounded)
                    w = self._pack(nd["weight"], kind, True, pad_to=(npad, kp))
                    bias = None
                    if nd["bias"] is not None:
                        bias = torch.zeros(npad, device=out.device, dtype=torch.float32)
                        bias[:n_ch].copy_(nd["bias"].detach())
                    nd["bias_buf"] = bias
                    outp = out if npad == n_ch else torch.empty(*out.shape[:-1], npad, device=out.device, dtype=torch.float32)
                    plan = self._plan(x, w, bias, outp, form=L.FORM_TCONV if nd["transposed"] else L.FORM_SCONV,
                                      ksize=nd["ksize"], stride=nd["stride"], n_ch=npad, act=nd["act"],
                                      round_out=npad == n_ch and nd["out"] in self._round_at_producer)
                    if not isinstance(plan, ops.ConvPlan):
                        raise L.IcadvError("tape program: padded contraction not taken by the tensor path")
                    self.fwd.append(plan)
                    if outp is not out:
                        self.fwd.append(FnLaunch(_copy_ch, outp, out, n_ch))
                        self._round_at_producer.discard(nd["out"])
                        for c in self.consumers.get(nd["out"], []):
                            if self.nodes[c]["kind"] == "shuffle":
                                self._round_at_producer.discard(self.nodes[c]["out"])
                    continue
                x = self._rounded_source(self.fwd, src, rounded) if tc else self.buf[src]
                w = self._pack(nd["weight"], kind, tc)
                bias = nd["bias"].detach().contiguous().clone() if nd["bias"] is not None else None
                nd["bias_buf"] = bias
                # conv -> (I)GDN with nothing else reading the conv output: the normalisation rides in the contraction's
                # epilogue (the same fused form program.StackProgram uses); the GDN node keeps only its backward launch
                cons = self.consumers.get(nd["out"], [])
                gnd = self.nodes[cons[0]] if len(cons) == 1 and self.nodes[cons[0]]["kind"] == "gdn" else None
                if tc and gnd is not None and nd["act"] == L.ACT_NONE and n_ch <= 256 and nd["out"] != self._out_id:
                    g = gnd["module"]
                    be, ga, gaT = g.effective_parameters(round_tf32=True)
                    gout = self.buf[gnd["out"]]
                    sc = torch.empty_like(gout)
                    try:
                        plan = self._plan(x, w, bias, gout, form=L.FORM_TCONV if nd["transposed"] else L.FORM_SCONV,
                                          ksize=nd["ksize"], stride=nd["stride"], n_ch=n_ch,
                                          epi=L.EPI_IGDN_FWD if g.inverse else L.EPI_GDN_FWD, gmat=ga, beta=be, out_scale=sc,
                                          round_out=gnd["out"] in self._round_at_producer)
                    except L.IcadvError:
                        plan = None
                    if isinstance(plan, ops.ConvPlan):
                        self._gdn.append((g, be, ga, gaT))
                        gnd["gaT"], gnd["sc"], gnd["fused_fwd"] = gaT, sc, True
                        self.fwd.append(plan)
                        continue
                plan = self._plan(x, w, bias, out, form=L.FORM_TCONV if nd["transposed"] else L.FORM_SCONV,
                                  ksize=nd["ksize"], stride=nd["stride"], n_ch=n_ch, act=nd["act"],
                                  round_out=nd["out"] in self._round_at_producer)
                if not isinstance(plan, ops.ConvPlan):
                    # the CUDA-core kernels do not round on store: consumers of this tensor get a rounding launch
                    self._round_at_producer.discard(nd["out"])
                    for c in self.consumers.get(nd["out"], []):
                        if self.nodes[c]["kind"] == "shuffle":
                            self._round_at_producer.discard(self.nodes[c]["out"])
                self.fwd.append(plan)
            elif nd["kind"] == "gdn":
                if nd.get("fused_fwd"):
                    continue
                g = nd["module"]
                be, ga, gaT = g.effective_parameters(round_tf32=True)
                self._gdn.append((g, be, ga, gaT))
                nd["gaT"] = gaT
                nd["sc"] = torch.empty_like(out)
                c = out.shape[-1]
                self.fwd.append(self._plan(self.buf[nd["in"][0]], None, None, out, form=L.FORM_SCONV, ksize=1, stride=1,
                                           n_ch=c, epi=L.EPI_IGDN_FWD if g.inverse else L.EPI_GDN_FWD, gmat=ga, beta=be,
                                           out_scale=nd["sc"], acc_from_in=True,
                                           round_out=nd["out"] in self._round_at_producer))
            elif nd["kind"] == "act":
                self.fwd.append(FnLaunch(_unary, self.buf[nd["in"][0]], None, out, _UN[nd["act"]]))
            elif nd["kind"] == "add":
                self.fwd.append(FnLaunch(_unary, self.buf[nd["in"][0]], self.buf[nd["in"][1]], out, 4))
            elif nd["kind"] == "shuffle":
                n, h, w, c = self.buf[nd["in"][0]].shape
                r = nd["r"]
                self.fwd.append(FnLaunch(_shuffle, self.buf[nd["in"][0]], out, n, h, w, c // (r * r), r, False))
            elif nd["kind"] == "gate":       # AttentionBlock: a * sigmoid(b) + x
                a, b, x = (self.buf[t] for t in nd["in"])
                self.fwd.append(FnLaunch(_gate, a, b, x, out))
            else:
                raise L.IcadvError(f"tape program: unsupported operator '{nd['kind']}'")

    # ------------------------------------------------------------------ backward (to the input)
    def _build_backward(self, in_id, out_id):
        contrib = {out_id: [self.g_out]}            # gradient contributions per tensor, one per consumer
        rounded_g = set()                           # gradient buffers whose producer already rounded them to TF32

        def grad_of(t, want_round=False):
            """One buffer holding dL/dt: the contributions of several consumers are summed by add launches (the last one
            rounds to TF32 on store when the only reader is a tensor-path input gradient)."""
            parts = contrib.get(t)
            if not parts:
                raise L.IcadvError("tape program: a tensor of the stack has no path to the output")
            acc = parts[0]
            for i, extra in enumerate(parts[1:]):
                dst = torch.empty_like(acc)
                last = want_round and i == len(parts) - 2
                self.bwd.append(FnLaunch(_unary, acc, extra, dst, 4 | (256 if last else 0)))
                if last:
                    rounded_g.add(dst.data_ptr())
                acc = dst
            return acc

        def rounds_its_gradient(nd):
            """A conv node whose input-gradient launch reads dL/d(out) as a plain TF32-rounded tensor (no activation
            gradient in between, no channel padding of the gradient)."""
            if nd["kind"] != "conv" or nd["act"] != L.ACT_NONE:
                return False
            n_out, k_in = self.buf[nd["out"]].shape[-1], self.buf[nd["in"][0]].shape[-1]
            return self._tc(n_out, k_in) or (self._padded(n_out, k_in) and n_out % 32 == 0)

        single_input_use = len(self.consumers.get(in_id, [])) == 1
        for nd in reversed(self.nodes):
            g = grad_of(nd["out"], want_round=rounds_its_gradient(nd))
            src = nd["in"][0]
            if nd["kind"] == "conv":
                k_ch = self.buf[src].shape[-1]
                tc = self._tc(g.shape[-1], k_ch)
                padded = self._padded(g.shape[-1], k_ch)
                plain_round = tc or (padded and g.shape[-1] % 32 == 0)
                if nd["act"] != L.ACT_NONE:         # activation fused into the forward epilogue: its gradient from the output,
                    gz = torch.empty_like(g)        # rounded on store when the input gradient reads it as it is
                    self.bwd.append(FnLaunch(_act_bwd, self.buf[nd["out"]], g, gz, nd["act"], plain_round))
                    if plain_round:
                        rounded_g.add(gz.data_ptr())
                    g = gz
                if padded and g.shape[-1] % 32:
                    g = self._padded_input(self.bwd, g, g.shape[-1])
                elif plain_round and g.data_ptr() not in rounded_g:
                    gr = torch.empty_like(g)
                    self.bwd.append(FnLaunch(_unary, g, None, gr, 5))
                    g = gr
                dkind = L.PACK_CONVT_DGRAD if nd["transposed"] else L.PACK_CONV_DGRAD
                w = self._pack(nd["weight"], dkind, True, pad_to=(_pad32(k_ch), g.shape[-1])) if padded else \
                    self._pack(nd["weight"], dkind, tc)
                if not nd["transposed"] and nd["ksize"] < nd["stride"]:
                    # input gradient of a strided 1x1 conv: the pixels no tap reaches are never written and stay zero
                    dst = torch.zeros_like(self.buf[src])
                elif src == in_id and single_input_use:
                    dst = self.g_in
                else:
                    dst = torch.empty_like(self.buf[src])
                if padded and k_ch % 32:
                    dstp = torch.zeros(*dst.shape[:-1], _pad32(k_ch), device=dst.device, dtype=torch.float32)
                    plan = self._plan(g, w, None, dstp, form=L.FORM_SCONV if nd["transposed"] else L.FORM_TCONV,
                                      ksize=nd["ksize"], stride=nd["stride"], n_ch=_pad32(k_ch))
                    if not isinstance(plan, ops.ConvPlan):
                        raise L.IcadvError("tape program: padded contraction not taken by the tensor path")
                    self.bwd.append(plan)
                    self.bwd.append(FnLaunch(_copy_ch, dstp, dst, k_ch))
                else:
                    self.bwd.append(self._plan(g, w, None, dst, form=L.FORM_SCONV if nd["transposed"] else L.FORM_TCONV,
                                               ksize=nd["ksize"], stride=nd["stride"], n_ch=k_ch))
                contrib.setdefault(src, []).append(dst)
            elif nd["kind"] == "gdn":
                dst = torch.empty_like(self.buf[src])
                gm = nd["module"]
                # dL/d(conv out) goes straight into that conv's input-gradient launch: rounded on store
                pi = self.producer.get(src)
                rnd = pi is not None and len(self.consumers.get(src, [])) == 1 and rounds_its_gradient(self.nodes[pi])
                self.bwd.append(self._plan(g, None, None, dst, form=L.FORM_SCONV, ksize=1, stride=1, n_ch=dst.shape[-1],
                                           epi=L.EPI_IGDN_BWD if gm.inverse else L.EPI_GDN_BWD, gmat=nd["gaT"],
                                           y_prev=self.buf[nd["out"]], sc_prev=nd["sc"], acc_from_in=True, round_out=rnd))
                if rnd:
                    rounded_g.add(dst.data_ptr())
                contrib.setdefault(src, []).append(dst)
            elif nd["kind"] == "act":
                dst = torch.empty_like(g)
                self.bwd.append(FnLaunch(_act_bwd, self.buf[src], g, dst, nd["act"]))
                contrib.setdefault(src, []).append(dst)
            elif nd["kind"] == "add":
                for t in nd["in"]:
                    contrib.setdefault(t, []).append(g)
            elif nd["kind"] == "shuffle":
                n, h, w, c = self.buf[src].shape
                r = nd["r"]
                dst = torch.empty_like(self.buf[src])
                self.bwd.append(FnLaunch(_shuffle, g, dst, n, h, w, c // (r * r), r, True))
                contrib.setdefault(src, []).append(dst)
            elif nd["kind"] == "gate":
                ta, tb, tx = nd["in"]
                ga, gb = torch.empty_like(g), torch.empty_like(g)
                self.bwd.append(FnLaunch(_gate_bwd, self.buf[ta], self.buf[tb], g, ga, gb))
                contrib.setdefault(ta, []).append(ga)
                contrib.setdefault(tb, []).append(gb)
                contrib.setdefault(tx, []).append(g)
        gin = grad_of(in_id)
        if gin is not self.g_in:
            self.bwd.append(FnLaunch(_unary, gin, None, self.g_in, 7))      # op 7: copy

    # ------------------------------------------------------------------ algorithmic work per launch (roofline tables)
    def _describe(self):
        """fwd_info / bwd_info for engine.launch_table(): every tensor a launch touches read or written once (fp32),
        FLOPs = 2 x MACs of the contraction (GDN: the C x C normalisation GEMM)."""
        def info(p, direction, k):
            if isinstance(p, (ops.ConvPlan, ops.SimtLaunch)):
                d = p._d if isinstance(p, ops.SimtLaunch) else p.desc
                oh, ow = ops.out_hw(d.form, d.ksize, d.stride, d.in_h, d.in_w)
                px = d.in_h * d.in_w if d.form == L.FORM_TCONV else oh * ow
                if d.acc_from_in:
                    macs, nbytes = d.n_img * oh * ow * d.n_ch * d.n_ch, 4.0 * d.n_img * oh * ow * d.n_ch * (3 if d.epi <= L.EPI_IGDN_FWD else 4)
                    name = "(I)GDN " + ("fwd" if d.epi <= L.EPI_IGDN_FWD else "bwd")
                else:
                    macs = d.n_img * px * d.ksize * d.ksize * d.k_ch * d.n_ch
                    nbytes = 4.0 * d.n_img * (d.in_h * d.in_w * d.k_ch + oh * ow * d.n_ch)
                    name = f"{'deconv' if (d.form == L.FORM_TCONV) == (direction == 'fwd') else 'conv'} {d.ksize}x{d.ksize}/{d.stride} " \
                           f"{d.k_ch}->{d.n_ch} {'' if direction == 'fwd' else 'dgrad'}"
                bound = "tensor" if Fn.tc_shape(d.k_ch, d.n_ch) or d.acc_from_in else "hbm"
                return {"name": f"{self.name}[{k}] {name}", "flops": 2.0 * macs, "bytes": nbytes, "bound": bound}
            t = [a for a in p.args if torch.is_tensor(a)]
            fn = p.fn.__name__.strip('_')
            if fn == "unary":       # which elementwise pass: add / round / copy ... (+ "r": rounds on store)
                op = int(p.args[3])
                fn = {0: "abs", 1: "relu", 2: "leaky", 3: "rint", 4: "add", 5: "round", 6: "clamp", 7: "copy"}[op & 255] + \
                    ("+r" if op & 256 else "")
            return {"name": f"{self.name}[{k}] elementwise ({fn})", "flops": 0.0,
                    "bytes": 4.0 * sum(a.numel() for a in t), "bound": "hbm"}
        self.fwd_info = [info(p, "fwd", k) for k, p in enumerate(self.fwd)]
        self.bwd_info = [info(p, "bwd", k) for k, p in enumerate(self.bwd)]

    # ------------------------------------------------------------------ run
    def forward(self):
        for p in self.fwd:
            p.launch()
        return self.out

    def backward(self):
        for p in self.bwd:
            p.launch()
        return self.g_in

    def n_kernels(self):
        return sum(p.kernels for p in self.fwd), sum(p.kernels for p in self.bwd)
