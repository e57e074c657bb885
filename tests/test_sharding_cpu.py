"""Host-side sharding logic, incl. a world_size-2 gloo run on CPU: the gathered result of a sharded
job equals the unsharded one bit for bit, and the per-image arg-min matches the reference's rule."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sharding


def fake_result(pair, width):
    """Deterministic stand-in for one trajectory's [loss, key(4), alpha(3)] row."""
    i, g = pair
    gen = torch.Generator().manual_seed(1000 * i + g)
    return torch.randn(width, generator=gen)


def run_rank(rank, world, port, total_imgs, n, width, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pairs = sharding.trajectory_list(total_imgs, n)
    mine = sharding.partition(len(pairs), rank, world)
    rows = []
    for b in sharding.batches(mine, 3):
        rows += [fake_result(pairs[t], width) for t in b]
    local = torch.stack(rows) if rows else torch.empty(0, width)
    full = sharding.gather_rows(local, len(pairs), rank, world)
    if rank == 0:
        torch.save(full, os.path.join(out_dir, "gathered.pt"))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("imgs,n", [(5, 3), (4, 4)])
def test_two_rank_gloo_gather_equals_unsharded(tmp_path, imgs, n):
    width = 8
    mp.spawn(run_rank, args=(2, free_port(), imgs, n, width, str(tmp_path)), nprocs=2, join=True)
    full = torch.load(os.path.join(str(tmp_path), "gathered.pt"))
    ref = torch.stack([fake_result(p, width) for p in sharding.trajectory_list(imgs, n)])
    assert torch.equal(full, ref)
    best, keys, alphas = sharding.select_best(full, n, key_len=4)
    for i in range(imgs):
        losses = [float(ref[i * n + g, 0]) for g in range(n)]
        assert int(best[i]) == losses.index(min(losses))
        assert torch.equal(keys[i], ref[i * n + int(best[i]), 1:5])
        assert torch.equal(alphas[i], ref[i * n + int(best[i]), 5:])


def test_partition_is_balanced_and_complete():
    for total in (0, 1, 7, 2000):
        for world in (1, 2, 3, 8):
            parts = [sharding.partition(total, r, world) for r in range(world)]
            assert sum(len(p) for p in parts) == total
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            flat = [t for p in parts for t in p]
            assert flat == list(range(total))
    assert [len(sharding.partition(2000, r, 8)) for r in range(8)] == [250] * 8


def test_bit_accuracy():
    logits = torch.tensor([[3.0, -2.0, 0.1, -0.1]])
    assert float(sharding.bit_accuracy(logits, torch.tensor([[1, 0, 1, 1]]))) == 0.75
