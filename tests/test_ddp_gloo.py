"""CPU, world_size 2 over gloo: the bucketed gradient exchange (egm-unet_b200/ddp.py) averages exactly like DDP, launches
buckets in reverse-registration order as soon as they complete, and tolerates parameters written in several pieces."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import egm_unet_b200  # noqa: F401
from egm_unet_b200.ddp import BucketReducer, make_buckets


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = [40, 8, 8, 120, 16, 64, 8, 200, 24]
    ranges, off = [], 0
    for n in sizes:
        ranges.append((off, n))
        off += n
    flat = torch.zeros(off)
    red = BucketReducer(flat, ranges, num_buckets=3)
    g = torch.Generator().manual_seed(100 + rank)
    local = [torch.randn(n, generator=g) for n in sizes]
    launched_after = []
    for i in range(len(sizes) - 1, -1, -1):           # backward order
        red.mark(i)
        o, n = ranges[i]
        flat[o:o + n] = local[i]                      # "kernel" writes the gradient
        red.mark(i)                                   # a second piece of the same parameter
        red.flush_ready()
        launched_after.append((i, list(red.order)))
    order = red.finish()
    ret[rank] = (flat.clone(), order, launched_after)
    dist.destroy_process_group()


def test_bucket_reducer_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    sizes = [40, 8, 8, 120, 16, 64, 8, 200, 24]
    expect = []
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        expect.append(torch.cat([torch.randn(n, generator=g) for n in sizes]))
    total = expect[0] + expect[1]
    for r in range(world):
        flat, order, launched_after = ret[r]
        assert torch.allclose(flat, total, atol=1e-6)           # SUM; the 1/world factor is applied by the SGD kernel
        assert order == [0, 1, 2]                               # buckets launched last-parameters-first
        assert launched_after[0][1] in ([], [0])                # nothing is reduced before its bucket is complete
    mean = total / world
    assert torch.allclose(mean, (expect[0] + expect[1]) / 2)


def test_make_buckets_cover_everything_once():
    ranges = [(0, 16), (16, 8), (24, 104), (128, 8), (136, 64)]
    b = make_buckets(ranges, 2)
    seen = sorted(i for _, _, idx in b for i in idx)
    assert seen == [0, 1, 2, 3, 4]
    assert b[0][2][-1] == 4 and b[0][0] >= b[-1][0]            # first bucket holds the LAST parameters
    covered = sorted((lo, hi) for lo, hi, _ in b)
    assert covered[0][0] == 0 and covered[-1][1] == 200 and all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
