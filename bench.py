"""Benchmark of the adversarial-perturbation hot path (attack_rd.py:506-560) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4|5]
  (N > 1: launched under torch.distributed.run, one rank per GPU)

Default = BASELINE.json configs[1], the configuration the headline metric is quoted on: Balle2018 hyperprior q=3, MSE
attack, 64 synthetic 768x512 images sharded over the N GPUs.  The other BASELINE configs are selectable:
  1  factorized q1, one 256x256 image, 100-step schedule (the reference's CPU-runnable case)
  3  Minnen2018 context model q4, -att_metric ms-ssim, 768x512 batch
  4  Cheng2020 (anchor) q6, targeted attack with an ROI mask, 768x512 batch
  5  hyperprior adversarial fine-tuning step (train.py --adv, 300 attack steps + codec update with the gradient
     all-reduce over NCCL), 8 x 256x256 crops per GPU

One *step* = one iteration of the loop body attack_rd.py:507-548 for every image of the rank's shard with the network
branch forced (branch B: g_a -> g_s forward, loss, backward to the input, Adam update), so no step skips the expensive
work; the natural branch mix of a real trajectory is reported in ``config.natural_mix``.  (Config 5: one step = one
training iteration = 300 attack iterations with the natural mix + one codec update.)

Timing: W >= 3 warm-up steps, then blocks of EXACTLY K steps, each bracketed by barrier + synchronize and timed with
CUDA events (max over ranks); blocks are repeated until >= 2 s have been timed and the MEDIAN block is reported
(``blocks`` lists all of them).  ``value`` = image-iterations/s with inputs resident in HBM; ``e2e`` = the same metric
through the public ``attack_()`` call with pinned host buffers (H2D, clean pass, loop, final eval, D2H inside the
timed region).  ``roofline`` = the launch that takes the most time, plus a per-launch table of every launch of the
step against its own bound (measured HBM copy bandwidth / measured tcgen05 kind::tf32 rate).

``--impl reference`` times the reference's algorithm on the host CPU cores: the oracle restatement (plain torch fp32
eager, parameter gradients left on as the reference does) on a bounded sample.  ``gpu_eager_baseline`` in our own line
is the same oracle under stock PyTorch on the GPU (cuDNN, TF32 allowed) -- the like-for-like bar of SURVEY.md 8(d).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    1: dict(model="factorized", quality=1, H=256, W=256, batch=1, att_metric="L2", sched_steps=100,
            name="Balle2016 factorized q1 MSE attack, one 256x256 image, 100-step schedule"),
    2: dict(model="hyper", quality=3, H=512, W=768, batch=64, att_metric="L2", sched_steps=1001,
            name="hyperprior(Balle2018) q3 MSE attack, 64 x 768x512"),
    3: dict(model="context", quality=4, H=512, W=768, batch=32, att_metric="ms-ssim", sched_steps=1001,
            name="Minnen2018 context q4 MS-SSIM attack, 32 x 768x512"),
    4: dict(model="cheng2020", quality=6, H=512, W=768, batch=8, att_metric="L2", sched_steps=1001, roi=True,
            name="Cheng2020-anchor q6 targeted attack with ROI mask, 8 x 768x512"),
    5: dict(model="hyper", quality=1, H=256, W=256, batch=8, att_metric="L2", sched_steps=300, train=True,
            name="hyperprior q1 adversarial fine-tuning step (train.py --adv -steps 300), 8 x 256x256 per GPU"),
}
CPU_SAMPLE_IMAGES = 2                # images of the workload the CPU legs time (bounded sample)
MIN_TIMED_S = 2.0


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, dev_index):
        super().__init__(daemon=True)
        self.idx, self.rows, self.stop_flag = dev_index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        pw = []
        for r in self.rows:
            try:
                pw.append(float(r[6]))
            except Exception:
                pass
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "power_w_max": max(pw) if pw else None}


def dist_env():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    return rank, world, int(os.environ.get("LOCAL_RANK", 0))


def workload_string(cfg):
    if cfg.get("train"):
        return cfg["name"]
    return cfg["name"] + ", forced branch B (every iteration = g_a,g_s fwd + input-grad bwd + Adam)"


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def _oracle_iteration(cfg, device, n_img, tf32=False):
    """One forced-branch-B iteration of attack_rd.py:507-548 on the oracle restatement (stock torch eager, parameter
    gradients left on as the reference does).  Returns a callable."""
    from oracle import attack as oatk
    from oracle import models as om
    from oracle.layers import low_bound, up_bound
    from oracle.msssim import ms_ssim
    net = om.init_model(cfg["model"], cfg["quality"], seed=0).to(device)
    x = torch.cat([oatk.synthetic_image(i, cfg["H"], cfg["W"]) for i in range(n_img)]).to(device)
    a = oatk.default_args(model=cfg["model"], quality=cfg["quality"], metric="mse")
    output_s, _, _ = oatk.clean_pass(x, net, a)
    net.train()
    noise = torch.zeros_like(x, requires_grad=True)
    opt = torch.optim.Adam([noise], lr=a.lr_attack)
    eps = a.epsilon / 255.0

    def one_iter():
        nc = up_bound(low_bound(noise, -eps), eps)
        im_in = up_bound(low_bound(x + nc, 0.0), 1.0)
        out = up_bound(low_bound(net.g_s(net.g_a(im_in)), 0.0), 1.0)
        if cfg["att_metric"] == "ms-ssim":
            loss = ms_ssim(out, output_s, data_range=1.0, size_average=True)          # attack_rd.py:362
        else:
            loss = 1.0 - torch.mean((output_s - out) * (output_s - out))               # attack_rd.py:364
        opt.zero_grad()
        loss.backward()
        opt.step()

    return one_iter


def cpu_sample(cfg, target_s, max_iters):
    """The oracle port on the host cores: a bounded sample of the workload (CPU_SAMPLE_IMAGES images, forced-branch-B
    iterations until about ``target_s`` seconds have been timed after one warm-up iteration)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_img = min(CPU_SAMPLE_IMAGES, cfg["batch"])
    one_iter = _oracle_iteration(cfg, torch.device("cpu"), n_img)
    times = []
    while len(times) < 2 or (sum(times[1:]) < target_s and len(times) < max_iters + 1):
        t0 = time.perf_counter()
        one_iter()
        times.append(time.perf_counter() - t0)
    dt = sum(times[1:]) / len(times[1:])
    return {"value": n_img / dt, "unit": "image-iterations/s", "cores": cores, "kind": "port",
            "sample": f"{n_img} images x {len(times) - 1} forced-branch-B iterations after 1 warm-up "
                      f"({sum(times[1:]):.1f} s of CPU time) of '{cfg['name']}', oracle port (torch eager fp32, "
                      f"wgrad on), {cores} threads"}, dt, len(times) - 1


def run_reference(args, cfg):
    """The reference algorithm on the box's host cores (oracle restatement; see oracle/__init__.py)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    if cfg.get("train"):
        print(json.dumps({"impl": "reference", "unavailable": "config 5 reference arm: use --config 2 (the attack loop "
                          "is 300/301 of the training step)"}), flush=True)
        return
    cpu, dt, iters = cpu_sample(cfg, target_s=30.0, max_iters=max(1, min(args.steps, 40)))
    v = cpu["value"]
    line = {"impl": "reference", "metric": "attack_image_iterations_per_sec", "value": v,
            "unit": "image-iterations/s", "n_gpus": args.gpus, "steps": iters, "warmup": 1,
            "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "gpu_launches": 0,
            "config": {"workload": workload_string(cfg), "model": cfg["model"], "quality": cfg["quality"],
                       "global_batch": cfg["batch"], "image": [cfg["H"], cfg["W"]]},
            "cpu_baseline": cpu,
            "e2e": {"value": v, "unit": "image-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(cfg, dev, batches=(8, 64)):
    """The like-for-like GPU bar: the oracle under stock PyTorch eager
    on this GPU -- cuDNN with TF32 allowed (PyTorch's default for convolutions), parameter gradients left on as the
    reference does -- forced-branch-B iterations at the given batch sizes."""
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    out = {"what": "oracle restatement under stock PyTorch eager (cuDNN, TF32 allowed, cudnn.benchmark, wgrad on), "
                   "forced branch B, same GPU", "unit": "image-iterations/s", "by_batch": {}}
    try:
        for b in batches:
            b = min(b, cfg["batch"])
            if str(b) in out["by_batch"]:
                continue
            try:
                one_iter = _oracle_iteration(cfg, dev, b, tf32=True)
                for _ in range(3):
                    one_iter()
                torch.cuda.synchronize()
                n, t0 = 0, time.perf_counter()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                while n < 5 or (time.perf_counter() - t0 < 1.5 and n < 200):
                    one_iter()
                    n += 1
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                out["by_batch"][str(b)] = {"value": b / (ms / 1e3), "ms_per_step": ms, "steps": n}
            except torch.OutOfMemoryError:
                out["by_batch"][str(b)] = {"value": None, "note": "out of memory"}
            del one_iter
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = prev
    vals = [v["value"] for v in out["by_batch"].values() if v.get("value")]
    out["value"] = max(vals) if vals else None
    return out


# ------------------------------------------------------------------------------------------------ our arm
class Timer:
    """Blocks of exactly K steps, CUDA events on the launching stream, barrier + synchronize on both sides, max over
    ranks; repeated until MIN_TIMED_S seconds have been timed."""

    def __init__(self, world, dev):
        self.world, self.dev = world, dev

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world > 1:
            t = torch.tensor([v], device=self.dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            return float(t)
        return v

    def block(self, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def blocks(self, fn, min_s=MIN_TIMED_S, max_blocks=200):
        ms = [self.block(fn)]
        n = int(min(max_blocks, max(3, math.ceil(min_s * 1e3 / max(ms[0], 1e-3)))))
        for _ in range(n - 1):
            ms.append(self.block(fn))
        return ms


def median(v):
    s = sorted(v)
    return s[len(s) // 2]


def measured_tf32_peak():
    """Burst (one ~1 ms launch timed alone) and sustained (~1 s of back-to-back launches) kind::tf32 tensor rate."""
    import ctypes as C
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    burst = ops.probe_tf32_peak(iters=60000, n=256, reps=4)      # best single ~16 ms launch (clocks already up)
    flops = C.c_double(0.0)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 60                      # 60 x ~16 ms
    e0.record()
    for _ in range(reps):
        L.call("icadv_probe_tf32_peak", 60000, 256, C.byref(flops), stream)
    e1.record()
    torch.cuda.synchronize()
    sustained = reps * flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return burst, sustained


def launch_table(eng, pk, tf32_burst):
    """Every launch of one iteration timed alone (CUDA events, best of 3; the batch's working set >> L2 except at small
    per-GPU batches) against its own bound: max(FLOP / measured TF32 burst rate, bytes / measured HBM copy bandwidth)."""
    rows = []
    for r in eng.launch_table():
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            r["launch"]()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        t_tensor = r["flops"] / (tf32_burst * 1e12) * 1e3
        t_hbm = r["bytes"] / (pk["hbm_gbs"] * 1e9) * 1e3
        bound_ms = max(t_tensor, t_hbm)
        rows.append({"name": r["name"], "kernels": r["kernels"], "ms": round(best, 4),
                     "gflop": round(r["flops"] / 1e9, 1), "mbytes": round(r["bytes"] / 1e6, 1),
                     "bound": "tensor" if t_tensor >= t_hbm else "hbm", "bound_ms": round(bound_ms, 4),
                     "frac": round(bound_ms / best, 3),
                     "tflops": round(r["flops"] / (best * 1e-3) / 1e12, 1), "gbs": round(r["bytes"] / (best * 1e-3) / 1e9, 1)})
    return rows


def build_attack(cfg, net, n_local, x, dev, steps_for_sched):
    """(engine, args namespace, output_s) for the attack configs."""
    from imagecompression_adversarial_b200 import attack as patk
    from imagecompression_adversarial_b200.engine import AttackEngine, RoiSpec, TapeAttackEngine
    a = argparse.Namespace(model=cfg["model"], quality=cfg["quality"], metric="mse", steps=steps_for_sched, random=1,
                           noise=1e-4, lr_attack=0.01, att_metric=cfg["att_metric"], epsilon=16.0, clamp=True, adv=False,
                           force_branch=1, mask_loc=None, lamb_bkg_in=1.0, lamb_bkg_out=1.0, lamb_tar=1.0)
    output_s, _ = patk.clean_pass(x, net, a)
    roi, output_t = None, None
    if cfg.get("roi"):
        a.mask_loc = [cfg["W"] // 4, 3 * cfg["W"] // 4, cfg["H"] // 4, 3 * cfg["H"] // 4]
        a.lamb_bkg_in, a.lamb_bkg_out, a.lamb_tar = 0.5, 2.0, 1.5
        roi = RoiSpec(a.mask_loc, a.lamb_bkg_in, a.lamb_bkg_out, a.lamb_tar)
        output_t, _ = patk.clean_pass(torch.roll(x, 1, 0) if x.shape[0] > 1 else torch.flip(x, (3,)), net, a)
    net.train()
    cls = AttackEngine if patk._fused_stacks(net) else TapeAttackEngine
    eng = cls(net, n_local, cfg["H"], cfg["W"], steps=cfg["sched_steps"], force_branch=1, att_metric=cfg["att_metric"], roi=roi)
    eng.load(x, output_s, None, output_t)
    return eng, a, output_s, output_t


def run_ours(args, cfg):
    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    from imagecompression_adversarial_b200 import attack as patk
    from imagecompression_adversarial_b200 import data
    from imagecompression_adversarial_b200 import models as pm
    from imagecompression_adversarial_b200 import ops
    from imagecompression_adversarial_b200.engine import AttackEngine
    from imagecompression_adversarial_b200.distributed import shard_indices
    ops.require_device()
    if cfg.get("train"):
        return run_train(args, cfg, rank, world, local, dev)
    torch.manual_seed(0)
    net = pm.init_model(cfg["model"], cfg["quality"], "mse", pretrained=False).to(dev)   # random-init weights (--new), seed 0
    batch = cfg["batch"]
    mine = shard_indices(batch, rank, world)
    n_local = len(mine)
    # seeded Kodak-like images on the k/255 lattice (SURVEY.md section 8d generator), this rank's shard
    host = data.synthetic_batch(mine, cfg["H"], cfg["W"]).pin_memory()
    x = host.to(dev, non_blocking=True)
    eng, a, output_s, output_t = build_attack(cfg, net, n_local, x, dev, max(args.steps, 3))
    timer = Timer(world, dev)
    warm = max(args.warmup, 3)
    eng.run(warm)
    timer.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    blocks = timer.blocks(lambda: eng.run(args.steps))
    sampler.stop_flag = True
    ms_block = median(blocks)
    ms_step = ms_block / args.steps
    value = batch * args.steps / (ms_block / 1e3)

    pk, pk_kind = peaks()
    roof, table, tf32 = None, None, None
    if rank == 0 and isinstance(eng, AttackEngine) and (eng.ga.fwd_info or type(eng) is AttackEngine):
        burst, sustained = measured_tf32_peak()
        tf32 = {"burst_tflops": round(burst, 1), "sustained_tflops": round(sustained, 1),
                "how": "bare tcgen05.mma kind::tf32 128x256x8 loop, one CTA per SM (icadv_probe_tf32_peak): best single "
                       "~16 ms launch (burst) / 60 back-to-back 16 ms launches (sustained)"}
        table = launch_table(eng, pk, burst)
        dom = max(table, key=lambda r: r["ms"])
        tc_rows = [r for r in table if r["gflop"] > 0]
        tc_ms = sum(r["ms"] for r in tc_rows)
        tc_flops = sum(r["gflop"] for r in tc_rows) * 1e9
        hbm_rows = [r for r in table if r["bound"] == "hbm"]
        if dom["bound"] == "hbm":
            achieved, peak, unit = dom["gbs"], pk["hbm_gbs"], "GB/s"
        else:
            achieved, peak, unit = dom["tflops"], burst, "TFLOP/s"
        roof = {"bound": dom["bound"], "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                "traffic": None,
                "kernel": dom["name"], "kernel_ms": dom["ms"], "share_of_step": dom["ms"] / ms_step,
                "algorithmic_flops_per_launch": dom["gflop"] * 1e9, "algorithmic_bytes_per_launch": dom["mbytes"] * 1e6,
                "peak_note": f"HBM: {pk_kind} copy bandwidth {pk['hbm_gbs']} GB/s (MEASURED_PEAKS.json); tensor: kind::tf32 "
                             f"burst rate measured in this run ({burst:.0f} TF/s; cuBLAS bf16 burst / 2 would be "
                             f"{pk['bf16_tflops'] / 2:.0f})",
                "step_bound_ms": round(sum(r["bound_ms"] for r in table), 3),
                "step_frac": round(sum(r["bound_ms"] for r in table) / ms_step, 3),
                "all_tensor_launches": {"ms_per_step": round(tc_ms, 3), "launches_per_step": len(tc_rows),
                                        "achieved_tflops": tc_flops / (tc_ms / 1e3) / 1e12,
                                        "frac_of_sustained_tf32": tc_flops / (tc_ms / 1e3) / 1e12 / sustained,
                                        "frac_of_burst_tf32": tc_flops / (tc_ms / 1e3) / 1e12 / burst,
                                        "share_of_step": tc_ms / ms_step},
                "hbm_bound_launches": {"ms_per_step": round(sum(r["ms"] for r in hbm_rows), 3),
                                       "bound_ms": round(sum(r["bound_ms"] for r in hbm_rows), 3),
                                       "gbs": sum(r["mbytes"] for r in hbm_rows) * 1e6 / (sum(r["ms"] for r in hbm_rows) * 1e-3) / 1e9},
                "launches": table}
        traffic = os.path.join(ROOT, "profiles", "r2_dominant_traffic.json")
        if os.path.exists(traffic):       # dram__bytes of the dominant launch from the committed ncu --set full capture
            t = json.load(open(traffic))
            if t.get("kernel") == dom["name"]:
                roof["traffic"] = t["dram_bytes_per_image"] * n_local
                roof["traffic_source"] = t["source"]

    # ---- natural branch mix on un-forced trajectories (reported, not timed)
    mix = None
    if rank == 0 and type(eng) is AttackEngine and cfg["att_metric"] == "L2":
        mix = {}
        for sched in sorted({60, cfg["sched_steps"]}):
            eng2 = AttackEngine(net, min(2, n_local), cfg["H"], cfg["W"], steps=sched, use_graph=False)
            eng2.load(x[:eng2.n_img], output_s[:eng2.n_img])
            rec = []
            eng2.run(sched, record=rec)
            nb = sum(int(r[0].sum()) for r in rec)
            mix[f"steps_{sched}"] = {"iterations": sched, "images": eng2.n_img,
                                     "branch_B_fraction": nb / (float(sched) * eng2.n_img)}
            del eng2

    kernels_per_iter = eng.kernels_per_iteration()
    del eng
    torch.cuda.empty_cache()

    # ---- end to end through the public API with host buffers
    e2e_steps = 60
    a2 = argparse.Namespace(**vars(a))
    a2.steps = e2e_steps
    out_pinned = torch.empty_like(host).pin_memory()
    host_t = None
    if cfg.get("roi"):
        host_t = (torch.roll(host, 1, 0) if host.shape[0] > 1 else torch.flip(host, (3,))).contiguous().pin_memory()

    def e2e_once():
        xin = host.to(dev, non_blocking=True)                       # pinned host -> device
        xt = host_t.to(dev, non_blocking=True) if host_t is not None else None
        im_adv, output_adv, _, bpp_ori, bpp, mse_r, vi_r = patk.attack_(xin, net, a2, im_t=xt)
        out_pinned.copy_(im_adv, non_blocking=True)                 # adversarial images -> pinned host
        torch.cuda.synchronize()
        return float(bpp)

    e2e_once()   # builds the engine and the cached inference programs for this shape (plans, buffers) ...
    e2e_once()   # ... and lets the caching allocator settle: the timed calls below are steady-state calls
    e2e_times = []
    while len(e2e_times) < 3 or (sum(e2e_times) < MIN_TIMED_S and len(e2e_times) < 50):
        timer.barrier()
        t0 = time.perf_counter()
        e2e_once()
        timer.barrier()
        e2e_times.append(timer.max_over_ranks(time.perf_counter() - t0))
    dt = median(e2e_times)
    h2d = host.numel() * 4 * (2 if host_t is not None else 1)
    e2e = {"value": batch * e2e_steps / dt, "unit": "image-iterations/s",
           "h2d_bytes_per_step": h2d * world // e2e_steps,
           "d2h_bytes_per_step": out_pinned.numel() * 4 * world // e2e_steps,
           "calls": len(e2e_times), "s_per_call": {"median": dt, "min": min(e2e_times), "max": max(e2e_times)},
           "note": f"attack_() with {e2e_steps} iterations per call: H2D + clean pass + loop + final eval (2x MS-SSIM) + "
                   f"D2H inside the timed region; median of {len(e2e_times)} calls, max over ranks"}

    if rank == 0:
        eager = gpu_eager_baseline(cfg, dev) if world == 1 else None
        torch.cuda.empty_cache()
        cpu = cpu_sample(cfg, target_s=12.0, max_iters=40)[0] if world == 1 else None
        line = {"metric": "attack_image_iterations_per_sec", "value": value, "unit": "image-iterations/s",
                "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "tf32",
                "data": "synthetic",
                "blocks": {"count": len(blocks), "steps_per_block": args.steps, "timed_s": sum(blocks) / 1e3,
                           "ms_per_step_median": ms_step, "ms_per_step_min": min(blocks) / args.steps,
                           "ms_per_step_max": max(blocks) / args.steps},
                "config": {"workload": workload_string(cfg), "bench_config": args.config,
                           "model": cfg["model"], "quality": cfg["quality"], "global_batch": batch,
                           "image": [cfg["H"], cfg["W"]], "images_per_gpu": n_local, "parallelism": f"image-shard x{world}",
                           "inputs": "seeded blurred-field images on the k/255 lattice (SURVEY 8d generator)",
                           "l2": "working set per step (GBs of activations) >> 126 MB L2",
                           "natural_mix": mix, "precision": "fp32 storage, TF32 tensor-core contractions (RN-rounded "
                                                            "operands), fp32 accumulate"},
                "roofline": roof, "tf32_peak_measured": tf32, "cpu_baseline": cpu, "gpu_eager_baseline": eager,
                "e2e": e2e, "gpu_launches": (kernels_per_iter or 0) * args.steps * len(blocks),
                "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_train(args, cfg, rank, world, local, dev):
    """BASELINE configs[4]: one train.py --adv iteration per step (train.py:335-366): attack_ with -steps 300 on the
    rank's batch (natural branch mix), train-mode forward, RD loss, backward with parameter gradients, ONE NCCL
    all-reduce of the flat gradient buffer, clip + Adam.  Weak scaling: 8 crops per GPU."""
    from imagecompression_adversarial_b200 import data
    from imagecompression_adversarial_b200 import models as pm
    from imagecompression_adversarial_b200 import training as ptr
    torch.manual_seed(0)
    net = pm.init_model(cfg["model"], cfg["quality"], "mse", pretrained=False).to(dev)
    n_local = cfg["batch"]
    host = data.synthetic_batch(range(rank * n_local, (rank + 1) * n_local), cfg["H"], cfg["W"]).pin_memory()
    a = argparse.Namespace(model=cfg["model"], quality=cfg["quality"], metric="mse", steps=cfg["sched_steps"], random=1,
                           noise=1e-4, lr_attack=0.01, att_metric="L2", epsilon=16.0, clamp=True, adv=True, lr_train=1e-5)
    crit = ptr.RateDistortionLoss("mse", ptr.LAMBDA_MSE[cfg["quality"]])
    opt, aux = ptr.configure_optimizers(net, a)
    timer = Timer(world, dev)
    parts = {"attack_ms": [], "update_ms": [], "allreduce_ms": []}

    def step():
        x = host.to(dev, non_blocking=True)
        out, _ = ptr.adv_train_step(x, net, a, crit, opt, aux, timings=parts)
        return float(out["loss"])            # device -> host read of the step's loss

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    for v in parts.values():
        v.clear()
    timer.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n_steps = max(1, min(args.steps, 10))
    blocks = timer.blocks(lambda: [step() for _ in range(n_steps)], max_blocks=5)
    sampler.stop_flag = True
    ms_step = median(blocks) / n_steps
    n_params = opt.flat.numel()
    if rank == 0:
        med = lambda v: median(v) if v else None
        line = {"metric": "adv_training_images_per_sec", "value": n_local * world / (ms_step / 1e3), "unit": "images/s",
                "n_gpus": world, "steps": n_steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
                "blocks": {"count": len(blocks), "timed_s": sum(blocks) / 1e3},
                "config": {"workload": cfg["name"], "bench_config": 5, "model": cfg["model"], "quality": cfg["quality"],
                           "images_per_gpu": n_local, "global_batch": n_local * world, "image": [cfg["H"], cfg["W"]],
                           "attack_steps": cfg["sched_steps"], "parallelism": f"data-parallel x{world}, one NCCL all-reduce "
                           f"of the flat gradient buffer per update ({n_params} fp32 = {n_params * 4 / 1e6:.1f} MB)"},
                "parts_ms": {"attack_300_iterations": med(parts["attack_ms"]), "codec_update": med(parts["update_ms"]),
                             "gradient_allreduce": med(parts["allreduce_ms"]),
                             "attack_image_iterations_per_sec": (n_local * world * cfg["sched_steps"] /
                                                                 (med(parts["attack_ms"]) / 1e3)) if parts["attack_ms"] else None},
                "allreduce_bytes": n_params * 4,
                "e2e": {"value": n_local * world / (ms_step / 1e3), "unit": "images/s",
                        "h2d_bytes_per_step": host.numel() * 4 * world, "d2h_bytes_per_step": 4 * world,
                        "note": "the step itself is end to end: pinned-host batch -> device, adv_train_step, loss read back"},
                "gpu_launches": None, "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
