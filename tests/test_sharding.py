"""The multi-GPU path shards independent pairs with no data-path collective.  world_size 2 on
gloo: every rank takes its contiguous shard of the synthetic pair list, the union is the whole list, and the
max-over-ranks reduction bench.py uses for its timing works."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from template_switch_aligner_b200 import workloads


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pairs = workloads.read_pairs(batch, start=rank * batch, length=40)   # the shard rule of bench.py
    cells = torch.tensor([sum(len(r) * len(q) for r, q in pairs)], dtype=torch.float64)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(cells, op=dist.ReduceOp.SUM)
    out[rank] = (pairs[0], pairs[-1], float(t.item()), float(cells.item()))
    dist.destroy_process_group()


def test_two_rank_sharding():
    world, batch = 2, 5
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, batch, out), nprocs=world, join=True)
        whole = workloads.read_pairs(world * batch, length=40)
        assert out[0][0] == whole[0] and out[0][1] == whole[batch - 1]
        assert out[1][0] == whole[batch] and out[1][1] == whole[-1]
        assert out[0][2] == out[1][2] == 2.0                      # max over ranks
        assert out[0][3] == sum(len(r) * len(q) for r, q in whole)  # whole-job work


def test_workload_is_seeded():
    a = workloads.read_pairs(4, start=7)
    b = workloads.read_pairs(4, start=7)
    assert a == b and a != workloads.read_pairs(4, start=8)
    r, q = workloads.read_pair(0)
    assert len(r) == 150 and set(r) <= set("ACGT")
