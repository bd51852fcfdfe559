"""Host-side multi-GPU logic on CPU: pair sharding and the one metrics all-reduce (gloo, world_size 2)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ogmm_b200 import pipeline


def test_shard_range_partitions_every_pair_once():
    for total in (0, 1, 7, 256, 1000, 8192):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = pipeline.shard_range(total, r, world)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi))
            assert seen == list(range(total))
            sizes = [pipeline.shard_range(total, r, world) for r in range(world)]
            assert max(h - l for l, h in sizes) <= -(-total // world)


def test_local_metrics_known_answer():
    eye = torch.eye(3).expand(4, 3, 3).contiguous()
    rot = eye.clone()
    c, s = torch.cos(torch.tensor(0.5 * torch.pi / 180)), torch.sin(torch.tensor(0.5 * torch.pi / 180))
    rot[1] = torch.tensor([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    t = torch.zeros(4, 3)
    t[2, 0] = 0.5
    v = pipeline.local_metrics(rot, t, eye, torch.zeros(4, 3))
    assert abs(float(v[0]) - 0.5) < 1e-2          # one pair off by 0.5 degree
    assert abs(float(v[1]) - 0.5) < 1e-6          # one pair off by 0.5 in translation
    assert float(v[2]) == 3.0 and float(v[3]) == 4.0


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 10
    lo, hi = pipeline.shard_range(total, rank, world)
    eye = torch.eye(3).expand(hi - lo, 3, 3).contiguous()
    trans = torch.full((hi - lo, 3), 0.01 * (rank + 1))
    vec = pipeline.local_metrics(eye, trans, eye, torch.zeros(hi - lo, 3))
    m = pipeline.reduce_metrics(vec)
    if rank == 0:
        out.put(m)
    dist.destroy_process_group()


def test_metrics_allreduce_world_size_2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    m = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert m["count"] == 10.0 and m["recall"] == 1.0
    expect = (5 * 0.01 * 3 ** 0.5 + 5 * 0.02 * 3 ** 0.5) / 10
    assert abs(m["mean_err_t"] - expect) < 1e-6 and m["mean_err_r_deg"] < 1e-3
