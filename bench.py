"""Headline benchmark: attack iterations/sec on synthetic 768x512 (Kodak-shaped) images, hyperprior
(Balle2018) q=3, MSE distortion attack (BASELINE.json configs[1]): 64 images sharded over N GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched under torch.distributed.run, one rank per GPU)

One *step* = one iteration of the loop body attack_rd.py:507-548 for every image of the rank's shard,
with the network branch forced (branch B: g_a -> g_s forward, loss, backward to the input, Adam update),
so no step skips the expensive work.  The natural branch mix of a real trajectory is reported in
``config.natural_mix`` from a short un-forced run.  ``value`` = image-iterations/s with inputs resident in
HBM; ``e2e`` = the same metric through the public ``attack_()`` call with host buffers (H2D of the images,
clean pass, loop, final eval, D2H of the adversarial images inside the timed region).

``--impl reference`` times the reference's algorithm on the host CPU cores: the oracle restatement
(plain torch fp32 eager, parameter gradients left on as the reference does) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 512, 768
GLOBAL_BATCH = 64
MODEL, QUALITY = "hyper", 3
FLOPS_PER_IMAGE_ITER = 132.67e9      # fwd + input-gradient of g_a, g_s
CPU_SAMPLE_IMAGES = 2                # images of the workload the CPU legs time (bounded sample)


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
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def dist_env():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    return rank, world, int(os.environ.get("LOCAL_RANK", 0))


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference algorithm on the box's host cores (oracle restatement; see oracle/__init__.py)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import attack as oatk
    from oracle import models as om
    from oracle.layers import low_bound, up_bound
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = om.init_model(MODEL, QUALITY, seed=0)
    n_img = CPU_SAMPLE_IMAGES
    x = torch.cat([oatk.synthetic_image(i, H, W) for i in range(n_img)])
    a = oatk.default_args(model=MODEL, quality=QUALITY, metric="mse")
    output_s, _, _ = oatk.clean_pass(x, net, a)
    net.train()
    noise = torch.zeros_like(x, requires_grad=True)
    opt = torch.optim.Adam([noise], lr=a.lr_attack)
    eps = a.epsilon / 255.0

    def one_iter():  # branch B of attack_rd.py:507-548, parameter gradients left on as in the reference
        nc = up_bound(low_bound(noise, -eps), eps)
        im_in = up_bound(low_bound(x + nc, 0.0), 1.0)
        out = up_bound(low_bound(net.g_s(net.g_a(im_in)), 0.0), 1.0)
        loss = 1.0 - torch.mean((output_s - out) * (output_s - out))
        opt.zero_grad()
        loss.backward()
        opt.step()

    steps = max(1, min(args.steps, 20))      # bounded sample: ~0.3 s per image-iteration on 16 cores
    for _ in range(min(args.warmup, 2)):
        one_iter()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_iter()
    dt = time.perf_counter() - t0
    v = n_img * steps / dt
    sample = (f"{n_img} images x {steps} forced-branch-B iterations (of the 64-image workload), oracle port, torch eager "
              f"fp32, wgrad on, {cores} threads")
    line = {"impl": "reference", "metric": "attack_image_iterations_per_sec", "value": v,
            "unit": "image-iterations/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2),
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "gpu_launches": 0,
            "config": {"workload": "hyperprior(Balle2018) q3 MSE attack, 768x512, forced branch B", "model": MODEL,
                       "quality": QUALITY, "global_batch": GLOBAL_BATCH, "image": [H, W]},
            "cpu_baseline": {"value": v, "unit": "image-iterations/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": v, "unit": "image-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def cpu_baseline_sample(target_s=12.0):
    """The oracle port on the host cores: CPU_SAMPLE_IMAGES images of the workload, forced-branch-B iterations until
    about `target_s` seconds of CPU work have been timed (after one warm-up iteration)."""
    from oracle import attack as oatk
    from oracle import models as om
    from oracle.layers import low_bound, up_bound
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = om.init_model(MODEL, QUALITY, seed=0)
    n_img = CPU_SAMPLE_IMAGES
    x = torch.cat([oatk.synthetic_image(i, H, W) for i in range(n_img)])
    a = oatk.default_args(model=MODEL, quality=QUALITY, metric="mse")
    output_s, _, _ = oatk.clean_pass(x, net, a)
    net.train()
    noise = torch.zeros_like(x, requires_grad=True)
    opt = torch.optim.Adam([noise], lr=a.lr_attack)
    eps = a.epsilon / 255.0
    times = []
    while len(times) < 2 or (sum(times[1:]) < target_s and len(times) < 41):
        t0 = time.perf_counter()
        nc = up_bound(low_bound(noise, -eps), eps)
        im_in = up_bound(low_bound(x + nc, 0.0), 1.0)
        out = up_bound(low_bound(net.g_s(net.g_a(im_in)), 0.0), 1.0)
        loss = 1.0 - torch.mean((output_s - out) * (output_s - out))
        opt.zero_grad()
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    dt = sum(times[1:]) / len(times[1:])
    return {"value": n_img / dt, "unit": "image-iterations/s", "cores": cores, "kind": "port",
            "sample": f"{n_img} images x {len(times) - 1} forced-branch-B iterations after 1 warm-up "
                      f"({sum(times[1:]):.1f} s of CPU time), oracle port (torch eager fp32, wgrad on)"}


def time_tc_kernels(eng):
    """CUDA-event time of every tensor-path launch of one iteration (same buffers, steady state)."""
    from imagecompression_adversarial_b200 import ops
    plans = [p for prog in (eng.ga, eng.gs) for lst in (prog.fwd, prog.bwd) for p in lst if isinstance(p, ops.ConvPlan)]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in plans]
    torch.cuda.synchronize()
    for p, (a, b) in zip(plans, ev):
        a.record()
        p.launch()
        b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    n_kernels = sum(p.kernels for p in plans)
    return ms, n_kernels


def run_ours(args):
    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    from imagecompression_adversarial_b200 import attack as patk
    from imagecompression_adversarial_b200 import models as pm
    from imagecompression_adversarial_b200 import ops
    from imagecompression_adversarial_b200.engine import AttackEngine
    from imagecompression_adversarial_b200.distributed import shard_indices
    ops.require_device()
    torch.manual_seed(0)
    net = pm.init_model(MODEL, QUALITY, "mse", pretrained=False).to(dev)   # random-init weights (--new), seed 0
    mine = shard_indices(GLOBAL_BATCH, rank, world)
    n_local = len(mine)
    g = torch.Generator().manual_seed(1234)
    # synthetic Kodak-shaped inputs on the k/255 lattice (cheap generator for the timed path; the seeded
    # blurred-field generator of the parity tests is oracle.attack.synthetic_image)
    host = (torch.randint(0, 256, (GLOBAL_BATCH, 3, H, W), generator=g, dtype=torch.uint8)[mine].float() / 255.0)
    host = host.pin_memory()
    x = host.to(dev, non_blocking=True)
    a = argparse.Namespace(model=MODEL, quality=QUALITY, metric="mse", steps=max(args.steps, 3), random=1, noise=1e-4,
                           lr_attack=0.01, att_metric="L2", epsilon=16.0, clamp=True, adv=False, force_branch=1)
    output_s, _ = patk.clean_pass(x, net, a)
    net.train()
    eng = AttackEngine(net, n_local, H, W, steps=1001, force_branch=1)
    eng.load(x, output_s)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    eng.run(max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    eng.run(args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t)
    sampler.stop_flag = True
    value = GLOBAL_BATCH * args.steps / (ms / 1e3)

    # ---- tensor-path launches timed alone on this rank's shard (CUDA events, same buffers, steady state)
    tc_ms, tc_kernels = time_tc_kernels(eng)
    tc_ms2, _ = time_tc_kernels(eng)
    tc_ms = min(tc_ms, tc_ms2)
    tc_flops = FLOPS_PER_IMAGE_ITER * n_local      # every contraction of the step is on the tensor path
    # dominant launch: g_a.2 = conv 128->128 5x5/2 + fused GDN on the 256x384 feature map
    dom = eng.ga.fwd[2]                            # [pad, g_a.0+GDN, g_a.2+GDN, ...]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        ev[0].record()
        dom.launch()
        ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    dom_flops = 2.0 * n_local * (H // 4) * (W // 4) * 128 * (25 * 128 + 128)     # 5x5x128 taps + the GDN 1x1
    pk, pk_kind = peaks()
    tf32_peak = pk["bf16_tflops"] / 2.0             # burst figure: this launch is timed alone
    achieved = dom_flops / (best / 1e3) / 1e12
    roof = {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
            "frac": achieved / tf32_peak,
            "traffic": 71.79e6 * n_local,           # ncu --set full on this launch at 8 images: 574.4 MB DRAM (read 414.3 + write 160.1), profiles/r1_ncu_full_g_a2_conv_gdn_v10.txt
            "kernel": "conv_tc_kernel<GDN_FWD> (tcgen05 kind::tf32), g_a.2 conv 128->128 5x5/2 + GDN",
            "kernel_ms": best, "algorithmic_flops_per_launch": dom_flops,
            "algorithmic_bytes_per_launch": n_local * 4.0 * 128 * ((H // 2) * (W // 2) + 2 * (H // 4) * (W // 4)),
            "peak_note": f"{pk_kind} bf16 burst {pk['bf16_tflops']} TF/s x 1/2 (TF32 issues at half the bf16 rate)",
            "all_tensor_launches": {"ms_per_step": tc_ms, "launches_per_step": tc_kernels,
                                    "achieved_tflops": tc_flops / (tc_ms / 1e3) / 1e12,
                                    "frac_of_sustained_tf32": tc_flops / (tc_ms / 1e3) / 1e12 /
                                    (pk["bf16_tflops_sustained"] / 2.0),
                                    "share_of_step": tc_ms / (ms / args.steps)}}

    # ---- natural branch mix on a short un-forced trajectory (reported, not timed)
    mix = None
    if rank == 0:
        eng2 = AttackEngine(net, min(2, n_local), H, W, steps=60, use_graph=False)
        eng2.load(x[:eng2.n_img], output_s[:eng2.n_img])
        rec = []
        eng2.run(60, record=rec)
        nb = sum(int(r[0].sum()) for r in rec)
        mix = {"iterations": 60, "images": eng2.n_img, "branch_B_fraction": nb / (60.0 * eng2.n_img)}
        del eng2

    # ---- end to end through the public API with host buffers
    e2e_steps = 60
    a2 = argparse.Namespace(**vars(a))
    a2.steps = e2e_steps
    kernels_per_iter = eng.kernels_per_iteration()
    del eng
    torch.cuda.empty_cache()

    out_pinned = torch.empty_like(host).pin_memory()

    def e2e_once():
        xin = host.to(dev, non_blocking=True)                       # pinned host -> device
        im_adv, output_adv, _, bpp_ori, bpp, mse_r, vi_r = patk.attack_(xin, net, a2)
        out_pinned.copy_(im_adv, non_blocking=True)                 # adversarial images -> pinned host
        torch.cuda.synchronize()
        return out_pinned, float(bpp)

    e2e_once()   # builds the engine and the cached inference programs for this shape (plans, buffers) ...
    e2e_once()   # ... and lets the caching allocator settle: the timed call below is a steady-state call
    e2e_times = []
    for _ in range(3):          # three timed calls, the median is reported (all three are listed in the note)
        barrier()
        t0 = time.perf_counter()
        out_host, _ = e2e_once()
        barrier()
        e2e_times.append(time.perf_counter() - t0)
    dt = sorted(e2e_times)[1]
    if world > 1:
        t = torch.tensor([dt], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        dt = float(t)
    e2e = {"value": GLOBAL_BATCH * e2e_steps / dt, "unit": "image-iterations/s",
           "h2d_bytes_per_step": host.numel() * 4 * world // e2e_steps,
           "d2h_bytes_per_step": out_host.numel() * 4 * world // e2e_steps,
           "note": f"attack_() with {e2e_steps} iterations: H2D + clean pass + loop + final eval (2x MS-SSIM) + D2H; "
                   f"median of 3 calls ({', '.join(f'{t:.3f}' for t in e2e_times)} s on rank 0)"}

    if rank == 0:
        cpu = cpu_baseline_sample()
        line = {"metric": "attack_image_iterations_per_sec", "value": value, "unit": "image-iterations/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "tf32",
                "data": "synthetic",
                "config": {"workload": "hyperprior(Balle2018) q3 MSE attack, 64 x 768x512, forced branch B "
                                       "(every iteration = g_a,g_s fwd + input-grad bwd + Adam)",
                           "model": MODEL, "quality": QUALITY, "global_batch": GLOBAL_BATCH, "image": [H, W],
                           "images_per_gpu": n_local, "parallelism": f"image-shard x{world}",
                           "l2": "working set per step (GBs of activations) >> 126 MB L2",
                           "natural_mix": mix, "precision": "fp32 storage, TF32 tensor-core contractions (RN-rounded "
                                                            "operands), fp32 accumulate"},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": (kernels_per_iter or 0) * args.steps, "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
