#!/usr/bin/env python3
"""Benchmark of the hot path on BASELINE.json's configs[1]: synthetic 150 bp read pairs with one planted template
switch, default `tsalign align` cost model (sample_tsa_config), batched per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch PAIRS_PER_GPU] [--impl ours|reference]

One "step" = one pass of the hot path (layer loop: primary fill + TS jump kernels) over one batch of pairs that is
already resident in HBM.  `value` = GCUPS (sum |R||Q| over all ranks' pairs / max-over-ranks time / 1e9), `e2e` = the
same metric through the C ABI call `tsa_align_batch` with host buffers (encode + H2D + kernels + D2H + result assembly
per step).
Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference's A* (oracle/, the
reference is Rust and cannot be built in this image) on the host cores for the same workload.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GCUPS (TS-aware cells/s), 150 bp read pairs with planted TSMs"
UNIT = "GCUPS"
READ_LEN = 150


def workload_config(batch, n_gpus):
    return {"workload": "configs[1]: synthetic 150 bp read pairs, one planted reverse TSM each, sample_tsa_config costs, dna-n",
            "pairs_per_step_per_gpu": batch, "pairs_per_step": batch * n_gpus, "read_length": READ_LEN,
            "l2_policy": "working set (D/seed matrices, 274 kB per pair) is far larger than the 126 MB L2",
            "parallelism": f"dp{n_gpus} over independent pairs, no data-path collective"}


# ------------------------------------------------------------------------------------------------ CPU arms
_FLAT = None


def _cpu_init():
    global _FLAT
    from oracle import oracle, tsa_config
    from template_switch_aligner_b200 import workloads
    _FLAT = oracle.FlatConfig(tsa_config.parse(workloads.sample_config_text(), "dna-n"))


def _cpu_astar(index):
    from oracle import oracle
    from template_switch_aligner_b200 import workloads
    r, q = workloads.read_pair(index, READ_LEN)
    t = time.perf_counter()
    res = oracle.astar_align(_FLAT, r, q)  # CLI defaults: lookahead min-length, maximise total TS length
    return index, len(r) * len(q), res.cost, time.perf_counter() - t


def cpu_astar_sample(first_index, budget_s, cores):
    """Run the restated reference A* on pairs first_index, first_index+1, ... on `cores` processes for about
    `budget_s` seconds.  Returns (pairs finished, cells finished, elapsed, costs by index, pairs started)."""
    ctx = mp.get_context("fork")
    started = 4 * cores
    t0 = time.perf_counter()
    done, cells, costs = 0, 0, {}
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        it = pool.imap_unordered(_cpu_astar, range(first_index, first_index + started))
        while True:
            remaining = budget_s - (time.perf_counter() - t0)
            if remaining <= 0:
                break
            try:
                idx, c, cost, _dt = it.next(timeout=remaining)
            except mp.TimeoutError:
                break
            except StopIteration:
                break
            done += 1
            cells += c
            costs[idx] = cost
        pool.terminate()
    return done, cells, time.perf_counter() - t0, costs, started


def run_reference(args):
    """--impl reference: the reference's own algorithm (restated A*, oracle/) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__  # builds oracle/ if needed
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    cores = os.cpu_count() or 1
    per_step_budget = max(5.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
    times, cells_total, pairs_total = [], 0, 0
    idx = 0
    for step in range(args.warmup + args.steps):
        done, cells, elapsed, _costs, started = cpu_astar_sample(idx, per_step_budget, cores)
        idx += started
        if step >= args.warmup:
            times.append(elapsed)
            cells_total += cells
            pairs_total += done
    total = sum(times)
    value = cells_total / total / 1e9 if total > 0 else 0.0
    sample = (f"{pairs_total} pairs finished in {args.steps} steps of {per_step_budget:.0f} s on {cores} processes "
              f"(4x{cores} pairs started per step; unfinished pairs count as no work)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": workload_config(args.batch, args.gpus),
            "pairs_per_s": pairs_total / total if total > 0 else 0.0,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.rows:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import template_switch_aligner_b200 as tsa
    from template_switch_aligner_b200 import _lib, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this benchmark has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.default()

    text = workloads.sample_config_text()
    batch = args.batch
    # weak scaling: every rank aligns its own contiguous shard of the pair list
    pairs = workloads.read_pairs(batch, start=rank * batch, length=READ_LEN)
    cells = sum(len(r) * len(q) for r, q in pairs)
    aligner = tsa.Aligner(costs=text, alphabet="dna-n", device=local, lib=lib, first_threshold=args.first_threshold, scout=args.scout)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm ------------------------------------------------------------------------------
    staged = tsa.StagedBatch(aligner, pairs)
    for _ in range(args.warmup):
        staged.run()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    jump_ms = fill_ms = 0.0
    launches = jump_launches = 0
    work = {"chains_started": 0, "chains_run": 0, "rows_filled": 0, "rows_jumped": 0}
    for _ in range(args.steps):
        staged.run()
        tm, st = staged.timing(), staged.stats()
        jump_ms += tm["jump_ms"]; fill_ms += tm["fill_ms"]
        launches += st["launches"]; jump_launches += st["jump_launches"]
        for key in work:
            work[key] += tm[key]
    barrier()
    elapsed = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    results = staged.fetch()
    stats = staged.stats()
    staged.close()
    experiment = bool(os.environ.get("TSA_BENCH_EXPERIMENT"))   # developer knob: timing of deliberately broken kernel variants
    bad = [] if experiment else [r for r in results if not r.found]
    if bad:
        raise SystemExit(f"bench.py: {len(bad)} pairs did not produce an alignment cost: {bad[0]}")

    # ---- end-to-end arm: public API with host buffers, every step ------------------------------------------
    # The timed call is the C ABI entry `tsa_align_batch` (include/tsalign_b200.h) on host buffers: an array of tsa_pair
    # pointing at the ASCII sequences in host memory in, an array of tsa_result (costs + run-length encoded alignments
    # in host memory) out.  Encoding, H2D, all kernels, D2H and result assembly are inside the call; building the
    # ctypes view of the inputs and turning the results into Python objects (the Python mirror's job) are outside.
    import ctypes as C
    from template_switch_aligner_b200 import api
    arr, keep = api._make_pairs(pairs)
    # like `tsalign align`: extension beyond the range and equal-cost ranges of every switch (a_star_aligner.rs:238-253)
    opt = api._options(False, local, None, None, first_threshold=args.first_threshold, traceback=True, scout=args.scout,
                       postprocess=api.POST_EXTEND_BEYOND_RANGE | api.POST_EQUAL_COST_RANGES)
    err = C.create_string_buffer(512)

    def abi_call():
        res = (_lib.TsaResult * max(1, batch))()
        rc = lib.tsa_align_batch(aligner.config._h, C.byref(opt), arr, batch, res, err, len(err))
        if rc != 0:
            raise SystemExit(f"bench.py: tsa_align_batch failed: {err.value!r}")
        return res

    lib.tsa_results_free(abi_call(), batch)  # warm the allocator (device buffers of the engine are sized once)
    barrier()
    t1 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for step in range(e2e_steps):
        res = abi_call()
        if step + 1 < e2e_steps:
            lib.tsa_results_free(res, batch)
    barrier()
    e2e_elapsed = max_over_ranks(time.perf_counter() - t1)
    e2e_costs = [res[i].cost for i in range(batch)]
    e2e_ops = sum(res[i].n_ops for i in range(batch))
    lib.tsa_results_free(res, batch)
    del keep
    assert experiment or (e2e_costs == [r.cost for r in results] and e2e_ops == sum(len(r.ops) for r in results))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = cells * world * args.steps / elapsed / 1e9
    e2e_value = cells * world * e2e_steps / e2e_elapsed / 1e9
    # algorithmic work of SURVEY.md 8(d) for the dominant kernel (k_ts_jump): everything but the primary-fill term
    w_total = sum(workloads.algorithmic_work(len(r), len(q), res.template_switches) for (r, q), res in zip(pairs, results))
    w_fill = sum((res.template_switches + 1) * 7.0 * len(r) * len(q) for (r, q), res in zip(pairs, results))
    w_jump = w_total - w_fill
    s16 = __import__("ctypes").c_double()
    s32 = __import__("ctypes").c_double()
    lib.tsa_measure_addmin_peak(local, s16, s32)
    jl = max(1, jump_launches)
    achieved = (w_jump * args.steps / jl) / ((jump_ms / jl) * 1e-3) / 1e12 if jump_ms > 0 else None
    peak = s16.value / 1e12
    # DRAM traffic of the dominant kernel from the committed ncu capture (profiles/): bytes per pair of the layer-0 launch,
    # scaled to the pairs of one launch here; nothing is measured under a profiler in this run
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_k_ts_jump5_traffic.json")) as fh:
            tr = json.load(fh)
        traffic = (tr["dram_read_bytes"] + tr["dram_write_bytes"]) / tr["pairs_in_launch"] * batch
        traffic_src = tr["source"]
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "integer (DPX add-min, packed s16x2 lanes)", "kernel": "k_ts_jump<5,false>", "achieved": achieved, "peak": peak, "unit": "Tadd-min/s",
                "frac": (achieved / peak) if achieved and peak else None, "traffic": traffic, "traffic_unit": "bytes per layer-0 launch (dram read + write)",
                "traffic_source": traffic_src,
                "peak_source": "measured in this run by tsa_measure_addmin_peak (back-to-back __viaddmin_s16x2 on all SMs); s32 rate %.2f T/s" % (s32.value / 1e12),
                "algorithmic_ops_per_launch": w_jump * args.steps / jl, "avg_launch_ms": jump_ms / jl,
                "kernel_share_of_step": {"jump_ms": jump_ms / args.steps, "fill_ms": fill_ms / args.steps, "step_ms": 1e3 * elapsed / args.steps},
                "note": "HBM is not the bound of this path (min-plus on-chip); MEASURED_PEAKS.json hbm_gbs is used only for the traceback traffic of later rounds"}

    cores = os.cpu_count() or 1
    done, ccells, celapsed, costs, started = cpu_astar_sample(0, args.cpu_budget, cores)
    by_index = {i: res.cost for i, res in enumerate(results)}
    mism = [i for i, c in costs.items() if i in by_index and by_index[i] != c]
    if mism and not experiment:
        raise SystemExit(f"bench.py: GPU cost differs from the CPU A* on pairs {mism}")
    cpu = {"value": ccells / celapsed / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
           "pairs_per_s": done / celapsed,
           "sample": f"pairs 0..{started - 1} of the same workload started on {cores} processes, {done} finished within {args.cpu_budget:.0f} s "
                     f"(restated reference A*, oracle/astar_oracle.cpp; costs equal to the GPU's on all finished pairs)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s16x2 (jump kernel) / s32 (primary fill)", "data": "synthetic", "config": workload_config(batch, world),
            "pairs_per_s": batch * world * args.steps / elapsed,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": stats["h2d_bytes"], "d2h_bytes_per_step": stats["d2h_bytes"],
                    "pairs_per_s": batch * world * e2e_steps / e2e_elapsed, "steps": e2e_steps},
            "gpu_launches": launches, "layers_per_step": stats["layers"], "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "work": {**{k: v / args.steps for k, v in work.items()},
                     "note": "per step: chain pairs started / surviving chain-level pruning, chain rows filled (2 chains x 160 columns each), "
                             "rows whose jump-in/jump-out was evaluated; the dense formula of SURVEY 8(d) assumes 99 rows per chain"},
            "alignments": "every pair returns its run-length encoded alignment (traceback kernel inside the timed step)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16384, help="pairs per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--first-threshold", type=int, default=0, help="tuning knob of the exact pruning (0 = library default)")
    ap.add_argument("--scout", action="store_true", help="tuning knob: enable the reverse-kinds scouting round")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU A* for the cpu_baseline object")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
