"""Data-parallel adversarial-training step under torchrun (NCCL): every rank attacks and back-propagates its own batch
shard, ONE all-reduce of the flat gradient buffer, fused clip + Adam; checks that all ranks end with identical weights
and reports the device time of the attack part and of the update part (reported separately).
  torchrun --nproc-per-node N scripts/train_step_ddp.py [batch_per_rank] [crop] [attack_steps]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import models as pm  # noqa: E402
from imagecompression_adversarial_b200 import training as ptr  # noqa: E402
from imagecompression_adversarial_b200.attack import attack_  # noqa: E402
from types import SimpleNamespace  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl")
dev = torch.device("cuda", local)
bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 8
crop = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
torch.manual_seed(0)                                   # identical initial weights on every rank
net = pm.init_model("hyper", 1, "ms-ssim", pretrained=False).to(dev)
args = SimpleNamespace(model="hyper", metric="mse", quality=1, steps=steps, random=1, noise=1e-4, lr_attack=0.01,
                       att_metric="L2", epsilon=16.0, clamp=True, adv=True, lr_train=1e-5)
crit = ptr.RateDistortionLoss("mse", ptr.LAMBDA_MSE[1])
opt, aux = ptr.configure_optimizers(net, args)
g = torch.Generator(device=dev).manual_seed(100 + rank)  # rank-dependent data shard
x = torch.rand(bsz, 3, crop, crop, device=dev, generator=g)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(2):                                     # one warm-up, one timed
    torch.cuda.synchronize()
    ev[0].record()
    adv = attack_(x, net, args)[0].detach()
    ev[1].record()
    net.train()
    out = crit(net(adv), adv)
    opt.zero_grad(); aux.zero_grad()
    out["loss"].backward()
    opt.step()
    a = net.aux_loss(); a.backward(); aux.step(None)
    ev[2].record()
    torch.cuda.synchronize()
digest = torch.stack([opt.flat.double().sum(), opt.flat.double().abs().sum()])
same = True
if world > 1:
    all_d = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    same = all(bool(torch.equal(all_d[0], d)) for d in all_d)
if rank == 0:
    print(json.dumps({"world": world, "batch_per_rank": bsz, "crop": crop, "attack_steps": steps,
                      "attack_ms": round(ev[0].elapsed_time(ev[1]), 2), "update_ms": round(ev[1].elapsed_time(ev[2]), 2),
                      "loss": float(out["loss"]), "weights_identical_across_ranks": same,
                      "flat_params": opt.flat.numel()}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
assert same
