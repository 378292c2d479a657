"""Device-resident attack loop: the body of ``attack_`` (attack_rd.py:506-560) for a batch of images with
PER-IMAGE semantics (every image behaves like its own N=1 call of the reference), no host sync.

Per iteration, all stream-ordered and captured in one CUDA graph:
  1. perturb_forward    noise -> eps-clamp -> im_in = clamp(im_s + nc, 0, 1); loss_i; branch A/B per image
                        (attack_rd.py:507,517,333-334); branch-B images are compacted into an active list;
                        LR schedule (MultiStepLR, attack_rd.py:503,553-554) and Adam coefficients on device
  2. g_a forward, g_s forward   (attack_rd.py:344,349) over the active images only
  3. output_loss        clamp + 1 - MSE(output_s, output_) + gradient seed  (attack_rd.py:353-364)
  4. g_s backward, g_a backward to the input  (attack_rd.py:547, input gradient only)
  5. perturb_update     clamp backward rule + Adam on the perturbation  (attack_rd.py:546-548)
"""
import ctypes as C
import os

import torch

from . import _lib as L
from . import ops
from .program import StackProgram, parse_stack


class RoiSpec:
    """Targeted / ROI attack configuration (flags ``-t``, ``--mask_loc``, ``-la_bkg_in``, ``-la_bkg_out``, ``-la_tar``;
    coder.py:198-203).  ``mask_loc = (x0, x1, y0, y1)``: W range then H range of the target area (attack_cv.py:160-161);
    None = the whole image is the target area."""

    def __init__(self, mask_loc=None, lamb_bkg_in=1.0, lamb_bkg_out=1.0, lamb_tar=1.0):
        self.mask_loc = tuple(mask_loc) if mask_loc is not None else None
        self.lamb_bkg_in, self.lamb_bkg_out, self.lamb_tar = float(lamb_bkg_in), float(lamb_bkg_out), float(lamb_tar)

    def key(self):
        return (self.mask_loc, self.lamb_bkg_in, self.lamb_bkg_out, self.lamb_tar)

    def maps(self, height, width, device):
        """(mask_tar, w_in, w_out) as [H, W, 3] channels-last maps (set-up, once per engine)."""
        tar = torch.ones(height, width, 3, device=device)
        if self.mask_loc is not None:
            x0, x1, y0, y1 = self.mask_loc
            tar.zero_()
            tar[y0:y1, x0:x1, :] = 1.0
        bkg = 1.0 - tar
        return tar, (tar + self.lamb_bkg_in * bkg).contiguous(), (self.lamb_tar * tar + self.lamb_bkg_out * bkg).contiguous()


class AttackEngine:
    def __init__(self, net, n_img, height, width, *, steps, epsilon=16.0, noise_budget=1e-4, lr_attack=0.01,
                 clamp=True, att_metric="L2", force_branch=-1, use_graph=True, device=None, roi=None,
                 budget_scope="image"):
        """``budget_scope``: "image" = every image tests its own loss_i against the budget (N independent runs of the
        reference CLI, which attacks one image at a time); "batch" = the reference's literal batch semantics -- loss_i,
        loss_o and the ms-ssim terms are means over the whole batch and ONE branch is taken per iteration
        (attack_rd.py:333-364 on a batch, i.e. what train.py:342 runs)."""
        ops.require_device()
        if budget_scope not in ("image", "batch"):
            raise L.IcadvError(f"AttackEngine: budget_scope {budget_scope!r} (image, batch)")
        self.batch_budget = budget_scope == "batch"
        # a batch-mean loss spreads its gradient: every per-image seed carries 1 / n_img (Adam's eps makes the scale matter)
        self._bscale = 1.0 / n_img if self.batch_budget else 1.0
        if att_metric not in ("L2", "ms-ssim"):
            raise L.IcadvError(f"AttackEngine: -att_metric {att_metric} is not supported (L2, ms-ssim)")
        self.att_metric = att_metric
        self.roi = roi
        if roi is not None and att_metric != "L2":
            raise L.IcadvError("AttackEngine: the targeted / ROI loss is defined for -att_metric L2 only")
        if steps < 3:
            raise ZeroDivisionError("integer division or modulo by zero")  # attack_rd.py:553 with steps//3 == 0
        dev = device or next(net.parameters()).device
        self.net, self.n_img, self.H, self.W, self.device = net, n_img, height, width, dev
        self.steps, self.eps, self.budget, self.lr0 = steps, epsilon / 255.0, noise_budget, lr_attack
        self.clamp, self.force_branch, self.use_graph = clamp, force_branch, use_graph
        f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        shape = (n_img, height, width, 3)
        self.per_img = 3 * height * width
        self.im_s, self.output_s = f(*shape), f(*shape)
        self.noise, self.m, self.v, self.im_in = f(*shape), f(*shape), f(*shape), f(*shape)
        self.st = ops.PerturbState(n_img, dev)
        self.loss_o_sum = f(n_img)
        self.ws_out = f(n_img * L.RED_BLOCKS)
        self.w_in = self.w_out = self.mask_tar = None
        if roi is not None:
            self.mask_tar, self.w_in, self.w_out = roi.maps(height, width, dev)
        act, nact = self.st.active, self.st.n_active
        self._build_network(net, n_img, height, width, dev, act, nact)
        self._graph = self._graph_k = None
        self._if_capture, self._if_stream = False, None
        self._graph_if_ok = True           # engines whose network pass allocates (autograd walk) switch it off
        self.iterations_done = 0
        self.im_s_nchw = self.output_s_nchw = None
        self._ones = torch.full((n_img,), self._bscale, device=dev, dtype=torch.float32)   # upstream of the ms-ssim terms
        if att_metric == "ms-ssim":
            from . import metrics
            metrics._weights_tensor(dev, metrics.WEIGHTS)   # created outside any graph capture
        self.last_loss_A = f(n_img)   # loss of the budget branch per image (loss_i, or 1 - ms_ssim(im_s, im_in))
        self.last_loss_B = f(n_img)   # loss of the network branch per image (1 - MSE, ms_ssim(out, output_s), or loss_o)

    def _build_network(self, net, n_img, height, width, dev, act, nact):
        f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        ga_units, gs_units = parse_stack(net.g_a), parse_stack(net.g_s)
        lat_h, lat_w = height, width
        for u in ga_units:
            lat_h, lat_w = ops.out_hw(u.fwd_form, u.k, u.s, lat_h, lat_w)
        # gradient wrt the latent: written by g_s.backward, seed of g_a.backward (one shared buffer)
        self.g_lat = f(n_img, lat_h, lat_w, ga_units[-1].cout)
        self.ga = StackProgram(ga_units, n_img, height, width, dev, x_in=self.im_in, g_out=self.g_lat, active=act,
                               n_active=nact, round_final_out=True, name="g_a")
        self.gs = StackProgram(gs_units, n_img, lat_h, lat_w, dev, x_in=self.ga.out, g_in=self.g_lat, active=act,
                               n_active=nact, round_final_gin=True, name="g_s")
        self.x_out, self.g_x = self.gs.out, self.gs.g_out
        assert self.x_out.shape == self.im_s.shape, (self.x_out.shape, self.im_s.shape)

    def launch_table(self):
        """Every launch of one L2-attack iteration, in stream order, with its algorithmic work:
        a list of dicts {name, launch (callable), kernels, flops, bytes, bound}.  ``bytes`` counts every tensor the
        launch must read or write once in fp32; ``bound`` names the roofline that limits it ("tensor" / "hbm").
        Used by bench.py's per-launch roofline table and scripts/step_breakdown.py (speed mode only)."""
        img = 4.0 * self.n_img * self.per_img           # bytes of one RGB tensor of the batch
        rows = [{"name": "perturb_forward + finalize", "launch": self._perturb_forward, "kernels": 1, "flops": 0.0,
                 "bytes": 3 * img, "bound": "hbm"}]
        msssim = self.att_metric == "ms-ssim"
        # MS-SSIM value + gradient (one pass per level, metrics.ms_ssim_value_and_grad): both images' 5-level pyramids
        # (4/3 of the image) read once, the gradient pyramid written once, + the two layout copies (read + write each).
        # Library kernels: 2 layout + 5 level value-and-gradient + 4 pools (the reference pyramid is cached) + 1 for all
        # the scalar work + 5 combine.
        ms_bytes = (4.0 / 3.0) * (2 + 1) * img + 4 * img   # (the cached reference pyramid is still READ every iteration)
        if msssim:
            rows.append({"name": "1 - ms_ssim(im_s, im_in) value + gradient (5 fused level launches + combine + layout copies)",
                         "launch": self._msssim_budget_branch, "kernels": 17, "flops": 0.0, "bytes": ms_bytes,
                         "bound": "hbm"})

        def stack(prog, lst, info):
            for p, i in zip(lst, info):
                rows.append({"name": i["name"], "launch": p.launch, "kernels": p.kernels, "flops": i["flops"],
                             "bytes": i["bytes"], "bound": i["bound"]})

        stack(self.ga, self.ga.fwd, self.ga.fwd_info)
        stack(self.gs, self.gs.fwd, self.gs.fwd_info)
        if msssim:
            rows.append({"name": "ms_ssim(clamp(x_out), output_s) value + gradient + clamp rules",
                         "launch": self._msssim_output_loss, "kernels": 17, "flops": 0.0, "bytes": ms_bytes + img,
                         "bound": "hbm"})
        else:
            rows.append({"name": "output_loss (clamp + MSE + gradient seed)",
                         "launch": lambda: self._output_loss(self.x_out, self.g_x), "kernels": 2, "flops": 0.0,
                         "bytes": 3 * img, "bound": "hbm"})
        stack(self.gs, self.gs.bwd, self.gs.bwd_info)
        stack(self.ga, self.ga.bwd, self.ga.bwd_info)
        if msssim:
            rows.append({"name": "perturb_update_adam (clamp backward + Adam, external budget-branch gradient)",
                         "launch": self._update_msssim, "kernels": 1, "flops": 0.0, "bytes": 9 * img, "bound": "hbm"})
            return rows
        rows.append({"name": "perturb_update_adam (clamp backward + Adam)",
                     "launch": lambda: ops.perturb_update_adam(self.im_s, self.noise, self.ga.g_in, self.m, self.v, self.st,
                                                               eps=self.eps, gradA_scale=self._bscale / self.per_img,
                                                               gradB_scale=1.0, w_in=self.w_in),
                     "kernels": 1, "flops": 0.0, "bytes": 8 * img, "bound": "hbm"})
        return rows

    # ------------------------------------------------------------------ state
    def load(self, im_s_nchw, output_s_nchw, noise_init_nchw=None, output_t_nchw=None):
        """Start a new attack on a batch: copies inputs in, zeroes the perturbation and Adam state.
        ROI attack: ``output_t_nchw`` is the clean reconstruction of the target image; the distortion reference becomes
        output_t inside the target area and output_s outside (a one-time select, not part of the loop)."""
        self.im_s.copy_(im_s_nchw.permute(0, 2, 3, 1))
        self.output_s.copy_(output_s_nchw.permute(0, 2, 3, 1))
        if self.roi is not None:
            if output_t_nchw is None:
                raise L.IcadvError("ROI attack: load() needs output_t")
            self.output_s.copy_(torch.where(self.mask_tar.unsqueeze(0) > 0, output_t_nchw.permute(0, 2, 3, 1),
                                            self.output_s))
        if self.att_metric == "ms-ssim":
            from . import metrics
            # persistent buffers, refilled in place: the captured iteration reads these addresses on every replay, also
            # after the engine is re-loaded with the next batch (a fresh clone per load would leave the graph reading
            # the previous batch's freed tensors)
            if self.im_s_nchw is None:
                self.im_s_nchw = torch.empty(im_s_nchw.shape, device=self.device, dtype=torch.float32)
                self.output_s_nchw = torch.empty_like(self.im_s_nchw)
                self._pyr_im_s = self._pyr_out_s = None
            self.im_s_nchw.copy_(im_s_nchw.detach())
            self.output_s_nchw.copy_(output_s_nchw.detach())
            # both reference images are fixed for the whole attack: their pool pyramids are built once per load
            self._pyr_im_s = metrics.reference_pyramid(self.im_s_nchw, out=self._pyr_im_s)
            self._pyr_out_s = metrics.reference_pyramid(self.output_s_nchw, out=self._pyr_out_s)
        if noise_init_nchw is None:
            self.noise.zero_()
        else:
            self.noise.copy_(noise_init_nchw.permute(0, 2, 3, 1))
        self.m.zero_()
        self.v.zero_()
        self.st.step.zero_()
        self.iterations_done = 0

    def refresh_parameters(self):
        self.ga.refresh_parameters()
        self.gs.refresh_parameters()

    # ------------------------------------------------------------------ one iteration
    def _perturb_forward(self):
        ops.perturb_forward(self.im_s, self.noise, self.im_in, self.st, eps=self.eps, budget=self.budget,
                            force_branch=self.force_branch, lr0=self.lr0, lr_gamma=0.33,
                            sched_period=self.steps // 3, w_in=self.w_in, ge_test=self.roi is not None,
                            batch_budget=self.batch_budget)

    def _output_loss(self, x_out, g_x):
        # untargeted: loss = 1 - mean(d^2) (attack_rd.py:364), seed +2d/P; ROI: loss = +mean(w d^2), seed -2wd/P
        sign = -1.0 if self.roi is not None else 1.0
        ops.output_loss(x_out, self.output_s, g_x, self.ws_out, self.loss_o_sum, do_clamp=self.clamp,
                        grad_scale=sign * self._bscale / self.per_img, active=self.st.active, n_active=self.st.n_active,
                        w_out=self.w_out)

    def _network_pass(self):
        self.ga.forward()
        self.gs.forward()
        self._output_loss(self.x_out, self.g_x)
        self.gs.backward()
        self.ga.backward()
        return self.ga.g_in

    def _iteration(self):
        if self.att_metric == "ms-ssim":
            return self._iteration_msssim()
        if self._if_capture:
            # graph capture of an un-forced loop: the network launches go into the body of an IF node on the device-side
            # count of images that take the network branch (csrc/icadv_graph.cu); perturb_forward sets the node's condition
            # itself (handle in the state it is launched with); nothing in the pass allocates
            cur = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            body = C.c_void_p(self._if_stream.cuda_stream)
            handle = C.c_uint64(0)
            L.call("icadv_graph_if_create", cur, C.byref(handle))
            self.st.c.cond_handle = handle.value
            try:
                self._perturb_forward()
            finally:
                self.st.c.cond_handle = 0
            L.call("icadv_graph_if_begin_handle", handle, cur, body)
            with torch.cuda.stream(self._if_stream):
                g_in = self._network_pass()
            L.call("icadv_graph_if_end", body)
        else:
            self._perturb_forward()
            g_in = self._network_pass()
        ops.perturb_update_adam(self.im_s, self.noise, g_in, self.m, self.v, self.st, eps=self.eps,
                                gradA_scale=self._bscale / self.per_img, gradB_scale=1.0, w_in=self.w_in)

    def _msssim_budget_branch(self):
        """Budget branch of -att_metric ms-ssim: loss = 1 - ms_ssim(im_s, im_in) (attack_rd.py:335-336) and its gradient
        with respect to im_in (channels-last), for every image."""
        from . import metrics
        ms_a, g_a = metrics.ms_ssim_value_and_grad(ops.nhwc_to_nchw(self.im_in), self.im_s_nchw, -self._ones,
                                                   y_pyramid=self._pyr_im_s)
        self.last_loss_A = 1.0 - ms_a
        self._g_a_ext = ops.nchw_to_nhwc(g_a)

    def _msssim_output_loss(self):
        """Network branch: loss = ms_ssim(clamp(x_out), output_s) (attack_rd.py:353-362); writes the gradient seed g_x."""
        from . import metrics
        x = self.x_out
        # the clamp and its gradient rules ride on the two layout copies around the pyramid (one launch each side)
        out = ops.clamp01_nhwc_to_nchw(x) if self.clamp else ops.nhwc_to_nchw(x)
        ms_b, g_o = metrics.ms_ssim_value_and_grad(out, self.output_s_nchw, self._ones, y_pyramid=self._pyr_out_s)
        direct = self.g_x.is_contiguous() and self.g_x.shape == x.shape
        if self.clamp:
            g = ops.clamp01_backward_nchw_to_nhwc(g_o, x, out=self.g_x if direct else None)
        else:
            g = ops.nchw_to_nhwc(g_o)
        if not (direct and self.clamp):
            self.g_x.copy_(g.view_as(self.g_x))
        self.last_loss_B = ms_b

    def _update_msssim(self):
        ops.perturb_update_adam(self.im_s, self.noise, self.ga.g_in, self.m, self.v, self.st, eps=self.eps,
                                gradA_scale=1.0, gradB_scale=1.0, g_a_ext=self._g_a_ext)

    def _iteration_msssim(self):
        """-att_metric ms-ssim (attack_rd.py:335-336, 360-362): budget branch loss = 1 - ms_ssim(im_s, im_in),
        network branch loss = ms_ssim(output_, output_s); the branch test itself stays on the L2 budget (:333-334).
        Every tensor this composition allocates (the two 5-level pyramids, the per-level sums, the layout copies) comes
        from the CUDA graph's private pool during capture and is reused by every replay."""
        ops.perturb_forward(self.im_s, self.noise, self.im_in, self.st, eps=self.eps, budget=self.budget,
                            force_branch=self.force_branch, lr0=self.lr0, lr_gamma=0.33,
                            sched_period=self.steps // 3, batch_budget=self.batch_budget)
        self._msssim_budget_branch()
        self.ga.forward()
        self.gs.forward()
        self._msssim_output_loss()
        self.gs.backward()
        self.ga.backward()
        self._update_msssim()

    def kernels_per_iteration(self):
        fa, ba = self.ga.n_kernels()
        fs, bs = self.gs.n_kernels()
        if self.att_metric == "ms-ssim":   # the two value-and-gradient compositions: library kernels only (launch_table)
            return 1 + 17 + fa + fs + 17 + bs + ba + 1
        return 1 + fa + fs + 2 + bs + ba + 1  # perturb fwd (1) + stacks + output_loss (2) + update (1)

    def run(self, iterations, record=None):
        """Run ``iterations`` loop iterations.  ``record`` (list) receives per-iteration
        (branch[n], loss_i[n], loss_o[n]) tensors on the host -- this syncs and is for tests only."""
        done = 0
        if self.use_graph and record is None and self.att_metric == "L2" and iterations >= 2 * self.GRAPH_UNROLL:
            # several iterations per replay: at small batches an iteration is a handful of launch-latency-bound nodes
            # (budget-branch iterations of an un-forced loop: four kernels) and the per-replay cost shows
            if self._graph_k is None:
                self._graph_k = self._capture(self.GRAPH_UNROLL)
            for _ in range(iterations // self.GRAPH_UNROLL):
                self._graph_k.replay()
            done = (iterations // self.GRAPH_UNROLL) * self.GRAPH_UNROLL
            self.iterations_done += done
        for _ in range(iterations - done):
            if self.use_graph and record is None:
                if self._graph is None:
                    self._graph = self._capture(1)
                self._graph.replay()
            else:
                self._iteration()
            if record is not None:
                if self.att_metric == "L2":
                    self.last_loss_A = self.st.loss_i
                    mse_o = self.loss_o_sum / self.per_img
                    self.last_loss_B = mse_o if self.roi is not None else 1.0 - mse_o
                br = self.st.branch.cpu().clone()
                loss = torch.where(br == 1, self.last_loss_B.cpu(), self.last_loss_A.cpu())
                record.append((br, self.st.loss_i.cpu().clone(), loss))
            self.iterations_done += 1

    GRAPH_UNROLL = 8

    def _capture(self, n_iter=1):
        # warm-up outside capture (lazy one-time initialisation inside the library), on a side stream
        state = [t.clone() for t in (self.noise, self.m, self.v, self.st.step)]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._iteration()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for t, c in zip((self.noise, self.m, self.v, self.st.step), state):
            t.copy_(c)
        g = torch.cuda.CUDAGraph()
        # un-forced L2 loops skip the network launches of iterations in which no image takes the network branch
        # (ICADV_GRAPH_IF=0 keeps them unconditional); a forced branch needs no test
        self._if_capture = (self.att_metric == "L2" and self.force_branch != 1 and self._graph_if_ok and
                            os.environ.get("ICADV_GRAPH_IF", "1") != "0")
        if self._if_capture:
            self._if_stream = torch.cuda.Stream()
        try:
            with torch.cuda.graph(g):
                for _ in range(n_iter):
                    self._iteration()
        finally:
            self._if_capture = False
        for t, c in zip((self.noise, self.m, self.v, self.st.step), state):
            t.copy_(c)
        return g

    # ------------------------------------------------------------------ results
    def im_in_nchw(self):
        """Current adversarial input (after the last perturb_forward) as NCHW."""
        return self.im_in.permute(0, 3, 1, 2)

    def finalize(self):
        """Recompute im_in from the final perturbation?  No: the reference evaluates the im_in of the LAST
        iteration (attack_rd.py:561,573), i.e. before the last optimizer step.  Returns that tensor."""
        return self.im_in_nchw()


class GenericAttackEngine(AttackEngine):
    """Same loop for codecs whose ``g_a`` / ``g_s`` are not plain conv/GDN stacks (cheng2020_anchor: residual blocks,
    sub-pixel convolutions): the network pass runs module by module through the autograd Functions of ``functional``
    (every FLOP still in the library), over the whole batch; the perturbation kernels, branch logic and Adam are the
    fused ones.  No CUDA graph (autograd allocates)."""

    def _build_network(self, net, n_img, height, width, dev, act, nact):
        if self.att_metric != "L2":
            raise L.IcadvError("GenericAttackEngine: only -att_metric L2 is built for non-stack codecs")
        self.use_graph = False
        self.g_x = torch.zeros(n_img, height, width, 3, device=dev, dtype=torch.float32)

    def refresh_parameters(self):
        pass

    def kernels_per_iteration(self):
        return -1

    def _network_pass(self):
        from . import functional as Fn
        net = self.net
        params = [p for p in net.parameters() if p.requires_grad]
        for p in params:               # the attack never reads parameter gradients (attack_rd.py:546-548)
            p.requires_grad_(False)
        try:
            x_in = self.im_in.permute(0, 3, 1, 2).detach().requires_grad_(True)
            with torch.enable_grad():
                out = net.g_s(net.g_a(x_in))
            x_out = Fn.to_nhwc(out.detach()).contiguous()
            self.g_x.zero_()           # rows of budget-branch images stay zero
            self._output_loss(x_out, self.g_x)
            out.backward(Fn.to_nchw(self.g_x))
        finally:
            for p in params:
                p.requires_grad_(True)
        return Fn.to_nhwc(x_in.grad).contiguous()


class TapeAttackEngine(AttackEngine):
    """The same device-resident loop for codecs whose ``g_a`` / ``g_s`` are not plain conv/GDN stacks (cheng2020_anchor:
    residual blocks, sub-pixel convolutions; BASELINE config 4): the two stacks are traced once into static launch
    programs (``tape.TapeProgram``: pre-allocated buffers, cached TMA plans, activations fused into contraction
    epilogues), so the iteration is the same stream-ordered launch sequence as AttackEngine's and replays from one CUDA
    graph -- no autograd walk, no per-call plan creation."""

    def _build_network(self, net, n_img, height, width, dev, act, nact):
        from .tape import TapeProgram
        f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        with torch.no_grad():
            probe = torch.zeros(1, 3, height, width, device=dev).contiguous(memory_format=torch.channels_last)
            lat = net.g_a(probe)
        lat_c, lat_h, lat_w = lat.shape[1], lat.shape[2], lat.shape[3]
        del probe, lat
        self.g_lat = f(n_img, lat_h, lat_w, lat_c)
        self.ga = TapeProgram(net.g_a, n_img, height, width, dev, x_in=self.im_in, g_out=self.g_lat, active=act,
                              n_active=nact, name="g_a")
        self.gs = TapeProgram(net.g_s, n_img, lat_h, lat_w, dev, x_in=self.ga.out, g_in=self.g_lat, active=act,
                              n_active=nact, name="g_s")
        self.x_out, self.g_x = self.gs.out, self.gs.g_out
        assert self.x_out.shape == self.im_s.shape, (self.x_out.shape, self.im_s.shape)


class CwEngine(AttackEngine):
    """One step of the C&W-style search of attack_cw.py (:111-140, 149-166): every iteration runs the network,
    loss = loss_i + c (1 - MSE_o) with a per-image weight c (zeroed on device once the image's reconstruction error exceeds
    1.1 x its target level), Adam with a CONSTANT learning rate (no scheduler in attack_cw.py).  ``c`` and ``level`` are
    device vectors the driver (attack.attack_cw) rewrites between blocks of iterations; the iteration itself is the same
    stream-ordered launch sequence (CUDA graph) as AttackEngine's network branch plus one combine kernel."""

    def __init__(self, net, n_img, height, width, *, epsilon=16.0, lr_attack=0.01, clamp=True, use_graph=True, device=None):
        super().__init__(net, n_img, height, width, steps=3, epsilon=epsilon, noise_budget=0.0, lr_attack=lr_attack,
                         clamp=clamp, att_metric="L2", force_branch=1, use_graph=use_graph, device=device)
        self.c = torch.zeros(n_img, device=self.device, dtype=torch.float32)
        self.level = torch.zeros(n_img, device=self.device, dtype=torch.float32)
        self.g_cw = torch.zeros_like(self.im_s)

    def _iteration(self):
        ops.perturb_forward(self.im_s, self.noise, self.im_in, self.st, eps=self.eps, budget=0.0, force_branch=1,
                            lr0=self.lr0, lr_gamma=1.0, sched_period=1 << 30)
        g_net = self._network_pass()       # d(1 - MSE_o)/d im_in; loss_o_sum[n] = sum (output_s - out)^2 of this forward
        ops.cw_combine(g_net, self.im_in, self.im_s, self.g_cw, self.c, self.level, self.loss_o_sum)
        ops.perturb_update_adam(self.im_s, self.noise, self.g_cw, self.m, self.v, self.st, eps=self.eps,
                                gradA_scale=0.0, gradB_scale=1.0)

    def kernels_per_iteration(self):
        return super().kernels_per_iteration() + 1


class IfgsmEngine:
    """I-FGSM / PGD / MI-FGSM loop of attack_ifgsm.py:364-438 on the same launch programs: every step is a network
    step (no budget branch): loss = mean((output_s - g_s(g_a(x)))^2), x += (eps/steps) sign(grad) [or the momentum
    form], projection onto [x0 - eps, x0 + eps]."""

    def __init__(self, net, n_img, height, width, *, steps, epsilon=16.0, momentum=False, device=None):
        ops.require_device()
        dev = device or next(net.parameters()).device
        self.n_img, self.steps, self.eps, self.momentum = n_img, steps, epsilon / 255.0, momentum
        f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        shape = (n_img, height, width, 3)
        self.per_img = 3 * height * width
        self.im_s, self.output_s, self.im_adv, self.g_mom = f(*shape), f(*shape), f(*shape), f(*shape)
        self.loss_sum, self.l1 = f(n_img), f(n_img)
        self.ws = f(n_img * L.RED_BLOCKS)
        ga_units, gs_units = parse_stack(net.g_a), parse_stack(net.g_s)
        lat_h, lat_w = height, width
        for u in ga_units:
            lat_h, lat_w = ops.out_hw(u.fwd_form, u.k, u.s, lat_h, lat_w)
        self.g_lat = f(n_img, lat_h, lat_w, ga_units[-1].cout)
        self.ga = StackProgram(ga_units, n_img, height, width, dev, x_in=self.im_adv, g_out=self.g_lat,
                               round_final_out=True)
        self.gs = StackProgram(gs_units, n_img, lat_h, lat_w, dev, x_in=self.ga.out, g_in=self.g_lat,
                               round_final_gin=True)

    def load(self, im_s_nchw, output_s_nchw, start_nchw=None):
        self.im_s.copy_(im_s_nchw.permute(0, 2, 3, 1))
        self.output_s.copy_(output_s_nchw.permute(0, 2, 3, 1))
        self.im_adv.copy_((im_s_nchw if start_nchw is None else start_nchw).permute(0, 2, 3, 1))
        self.g_mom.zero_()

    def run(self, iterations, record=None):
        for _ in range(iterations):
            self.ga.forward()
            self.gs.forward()
            # loss = +mean(d^2): d loss / d out = -2 d / P  (no output clamp in attack_ifgsm.py:394-396)
            ops.output_loss(self.gs.out, self.output_s, self.gs.g_out, self.ws, self.loss_sum, do_clamp=False,
                            grad_scale=-1.0 / self.per_img)
            self.gs.backward()
            self.ga.backward()
            if self.momentum:
                ops.mifgsm_update(self.im_s, self.im_adv, self.ga.g_in, self.g_mom, self.ws, self.l1,
                                  alpha=self.eps / self.steps, eps=self.eps, mu=1.0)
            else:
                ops.ifgsm_update(self.im_s, self.im_adv, self.ga.g_in, alpha=self.eps / self.steps, eps=self.eps)
            if record is not None:
                record.append((self.loss_sum / self.per_img).cpu().clone())

    def im_adv_nchw(self):
        return self.im_adv.permute(0, 3, 1, 2)
