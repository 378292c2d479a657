"""Host-side multi-rank logic on CPU (gloo, world_size 2): image sharding + final metric gather."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from imagecompression_adversarial_b200.distributed import gather_metrics, shard_indices
    mine = shard_indices(n_items, rank, world)
    local = torch.tensor([[float(i), float(i) * 10.0] for i in mine]).reshape(len(mine), 2)
    table = gather_metrics(local, n_items, rank, world)
    q.put((rank, mine, table.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 8])
def test_shard_and_gather_world2(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_items
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = sorted(sum((r[1] for r in res), []))
    assert shards == list(range(n_items))           # disjoint and complete
    want = [[float(i), float(i) * 10.0] for i in range(n_items)]
    for r in res:
        assert r[2] == want                          # same table, image order, on every rank


def test_shard_indices_properties():
    from imagecompression_adversarial_b200.distributed import shard_indices
    for n in (1, 5, 64):
        for w in (1, 2, 4, 8):
            parts = [shard_indices(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _grad_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from imagecompression_adversarial_b200.training import exchange_gradients
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)        # rank-dependent "gradients"
    scale = exchange_gradients(flat)                                   # one SUM all-reduce of the flat buffer
    q.put((rank, scale, flat.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_training_gradient_exchange_world2():
    """The one exchange step of data-parallel adversarial training (train.py --adv): a SUM all-reduce
    of the flat gradient buffer and the 1/world factor the fused clip + Adam kernel applies."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [float(i) * 3.0 for i in range(10)]                         # (1 + 2) * i on both ranks
    for rank, scale, flat in res:
        assert scale == 0.5 and flat == want


def test_training_gradient_exchange_single_process():
    from imagecompression_adversarial_b200.training import exchange_gradients
    flat = torch.ones(4)
    assert exchange_gradients(flat) == 1.0 and flat.tolist() == [1.0] * 4
