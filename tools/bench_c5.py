#!/usr/bin/env python3
"""BASELINE.json configs[4]: one ~230 kb pair without template switches, column-banded over the GPUs of one box under a memory
limit (csrc/tsa_band.cuh, tsa_long.cu).  One JSON line on stdout (rank 0).

    python tools/bench_c5.py --gpus N [--length L] [--memory-limit BYTES] [--steps K] [--check]        # one process, N devices
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_c5.py --gpus N ...               # one process per GPU

Single process: `tsa_align_long` drives all bands (peer access between neighbouring devices).  torchrun: every rank owns one band
(`tsa_long_*`), the 64-byte CUDA IPC handles of the boundary buffers are exchanged with torch.distributed, the boundary columns
then stream with plain stores over NVLink; the traceback state is handed from rank to rank (broadcast of 40 bytes per hand-over).
`--check`: also aligns the pair on ONE device the ordinary way (codes of the whole matrix resident, `tsa_align_batch`) and
compares cost and alignment cost; for lengths the CPU oracle finishes it compares with the oracle too."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import template_switch_aligner_b200 as tsa  # noqa: E402
from template_switch_aligner_b200 import api, workloads  # noqa: E402
from bench_shapes import c5_pair  # noqa: E402


def line_of(args, world, mode, r, q, res, stats, wall_s, steps):
    cells = len(r) * len(q)
    fwd = max(s["forward_ms"] for s in stats)
    trace = sum(s["trace_ms"] for s in stats)
    nvlink = sum(s["boundary_bytes_out"] for s in stats)
    return {"shape": f"c5: one {len(r)} x {len(q)} pair, --no-ts, alignment, column bands", "n_gpus": world, "mode": mode, "steps": steps,
            "cost": res.cost, "runs": len(res.ops or []), "value_gcups": cells / wall_s / 1e9, "ms_per_step": wall_s * 1e3,
            "forward_ms_max_over_bands": fwd, "forward_gcups": cells / (fwd * 1e-3) / 1e9 if fwd else None, "trace_ms": trace,
            "tiles": sum(s["tiles"] for s in stats), "tile_cells": sum(s["tile_cells"] for s in stats), "recomputed_fraction": sum(s["tile_cells"] for s in stats) / max(1, cells),
            "nvlink_bytes_per_step": nvlink, "nvlink_bytes_algorithmic": 8 * (len(r) + 1) * (world - 1),
            "resident_bytes_max": max(s["resident_bytes"] for s in stats), "memory_limit": args.memory_limit,
            "interval": stats[0]["interval"], "group": stats[0]["group"], "per_band": stats}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--length", type=int, default=230147)
    ap.add_argument("--memory-limit", type=int, default=64_000_000_000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--interval", type=int, default=0)
    ap.add_argument("--group", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    text = workloads.sample_config_text()
    r, q = c5_pair(args.length)

    if world == 1:
        aligner = tsa.Aligner(costs=text, no_ts=True, device=0)
        devices = list(range(args.gpus))
        best, res, stats = None, None, None
        for _ in range(args.steps + 1):
            t = time.perf_counter()
            res, stats = api.align_long(aligner, r, q, devices=devices, interval=args.interval, group=args.group, memory_limit=args.memory_limit)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        assert res.found, res
        line = line_of(args, args.gpus, "one process, peer access", r, q, res, stats, best, args.steps)
    else:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        aligner = tsa.Aligner(costs=text, no_ts=True, device=local)
        best, res, all_stats = None, None, None
        for _ in range(args.steps + 1):
            band = api.LongBand(aligner, r, q, rank, world, interval=args.interval, group=args.group, memory_limit=args.memory_limit)
            handles = [None] * world
            dist.all_gather_object(handles, band.export_handle())
            if rank + 1 < world:
                band.connect(handles[rank + 1])
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t = time.perf_counter()
            band.forward()
            # traceback: the walk state travels by broadcast from the rank that holds it
            state = [None]
            if rank == world - 1:
                cost, kind = band.cost()
                assert kind == "FoundTarget"
                state = [((band.rows, band.columns, 0, 1, cost), 0, world - 1, cost)]
            dist.broadcast_object_list(state, src=world - 1)
            (st, status, owner, cost), ops = state[0], b""
            segments = []
            while status == 0:
                if rank == owner:
                    st, status, seg = band.walk(st)
                    segments.append(seg)
                    nxt = band.owner(st[1]) if status == 0 else owner
                    state = [(st, status, nxt, cost)]
                src = owner
                dist.broadcast_object_list(state, src=src)
                st, status, owner, cost = state[0]
            assert status == 1, status
            gathered = [None] * world
            dist.all_gather_object(gathered, b"".join(segments))
            torch.cuda.synchronize(); dist.barrier()
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
            all_stats = [None] * world
            dist.all_gather_object(all_stats, band.stats())
            if rank == 0:
                # every band is visited at most once, from the right to the left
                ops = b"".join(gathered[k] for k in range(world - 1, -1, -1))
                res = band.result(cost, ops)
            band.close()
        if rank != 0:
            dist.destroy_process_group()
            return 0
        line = line_of(args, world, "one process per GPU (torchrun), CUDA IPC", r, q, res, all_stats, best, args.steps)
        dist.destroy_process_group()

    if args.check:
        from oracle import oracle, tsa_config
        flat = oracle.FlatConfig(tsa_config.parse(text, "dna-n"))
        cost, er, eq, ok = oracle.rescore(flat, r, q, [oracle.Op(*o) for o in res.ops], 0, 0)
        line["rescored_cost"] = cost
        assert ok and cost == res.cost and (er, eq) == (len(r), len(q)), (cost, res.cost, er, eq)
        one = tsa.Aligner(costs=text, no_ts=True, device=0).align_batch([(r, q)])[0]
        line["one_gpu_resident_codes_cost"] = one.cost
        assert one.found and one.cost == res.cost
        if len(r) <= 20000:
            want = oracle.dp_align(flat, r, q, no_ts=True)
            line["oracle_cost"] = want.cost
            assert want.cost == res.cost
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
