"""N > 1 host-side logic on CPU: bucket construction over the gradient arena and the sum all-reduce of its slices,
world_size 2 over gloo (rendezvous on 127.0.0.1)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from touhouimageclassification_b200.model import ViTConfig, ViTForImageClassification
from touhouimageclassification_b200.parallel import GradBucketer, make_buckets, stage_grad_ranges

TINY = dict(hidden_size=128, num_hidden_layers=4, num_attention_heads=2, intermediate_size=256, image_size=32, num_labels=10)


def test_buckets_partition_the_arena():
    m = ViTForImageClassification(ViTConfig(**TINY))
    ranges = stage_grad_ranges(m)
    assert len(ranges) == TINY["num_hidden_layers"] + 2
    for target in (1, 200_000, 400_000, 10**9):
        buckets = make_buckets(ranges, target)
        # stages are covered once, in backward order
        assert buckets[0][0] == 0 and buckets[-1][1] == len(ranges)
        for a, b in zip(buckets, buckets[1:]):
            assert a[1] == b[0]
        spans = sorted((b, e) for _, _, b, e in buckets)
        assert spans[0][0] == 0 and spans[-1][1] == m._total
        for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
            assert e0 == b1
        # every bucket is exactly the union of its stages' ranges
        for s0, s1, b, e in buckets:
            assert sum(r[1] - r[0] for r in ranges[s0:s1]) == e - b
    assert len(make_buckets(ranges, 1)) == len(ranges)
    assert len(make_buckets(ranges, 10**9)) < len(ranges)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, buckets, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.arange(total, dtype=torch.float32) * (rank + 1)
        bucketer = GradBucketer()
        assert bucketer.world_size == world
        for _, _, b, e in buckets:
            bucketer.all_reduce(g, b, e)
        expect = torch.arange(total, dtype=torch.float32) * sum(r + 1 for r in range(world))
        ok = torch.equal(g, expect)
        # identical AdamW inputs on every rank -> parameters stay bit-identical
        gathered = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(gathered, g)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        if rank == 0:
            out.put((ok, same))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    m = ViTForImageClassification(ViTConfig(**TINY))
    buckets = make_buckets(stage_grad_ranges(m), 300_000)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, m._total, buckets, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, same = out.get(timeout=10)
    assert ok and same
