#!/usr/bin/env python3
"""Pins the layered DP (and through it the CUDA path) to the reference's own algorithm at BASELINE config 2 size.

Runs the restated reference A* (oracle/astar_oracle.cpp: generic_a_star/src/lib.rs:316-552 over
lib_tsalign/.../template_switch_distance/context.rs:125-729, default strategies of `tsalign align`) TO COMPLETION on a fixed
seeded subsample of the bench workload -- pairs 0..N-1 of workloads.read_pairs (150 bp read pairs, one planted reverse
template switch, sample_tsa_config) -- and records per pair: optimal cost, total TS length of the returned alignment, opened
nodes, seconds, nodes/s.  Hard pairs take minutes and gigabytes (one 67 M-node pair: 52 s, 7 GB), which is why this runs
offline and its result is committed:

    python tests/golden/make_astar_c2.py [--pairs 32] [--workers 3]   ->  tests/golden/astar_c2.json

A pair whose search exceeds --max-nodes opened + closed nodes is recorded as {"result": "ExceededMemoryLimit"} (the
reference's own --memory-limit outcome) and is not used as a cost vector.  Consumers: tests/test_oracle_kat.py (DP == A* on
these pairs, CPU), tests/test_gpu_parity.py (GPU == A*), bench.py --impl reference (replays the recorded node counts and
seconds as the CPU arm's fixed sample).
"""
import argparse
import json
import multiprocessing as mp
import os
import platform
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def work(task):
    idx, max_nodes = task
    from oracle import oracle, tsa_config
    from template_switch_aligner_b200 import workloads
    flat = oracle.FlatConfig(tsa_config.parse(workloads.sample_config_text(), "dna-n"))
    r, q = workloads.read_pair(idx, 150)
    t0 = time.time()
    # memory_limit is in the reference's bytes-per-node accounting (160 B x 2.3, see astar_oracle.cpp)
    a = oracle.astar_align(flat, r, q, memory_limit=int(max_nodes * 160 * 2.3))
    dt = time.time() - t0
    d = oracle.dp_align(flat, r, q)
    rec = {"index": idx, "result": a.result_type, "opened_nodes": a.opened_nodes, "closed_nodes": a.closed_nodes,
           "seconds": round(dt, 3), "nodes_per_s": round(a.opened_nodes / max(dt, 1e-9)), "dp_cost": d.cost,
           "reference_len": len(r), "query_len": len(q)}
    if a.found:
        rec.update({"cost": a.cost, "ts_total_length": a.ts_total_length, "cigar": a.cigar(),
                    "template_switches": sum(1 for o in a.ops if o.type == oracle.OP_TS_EXIT)})
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=32)
    ap.add_argument("--workers", type=int, default=3)
    ap.add_argument("--max-nodes", type=float, default=150e6)
    ap.add_argument("--out", default=os.path.join(HERE, "astar_c2.json"))
    args = ap.parse_args()
    recs = []
    t0 = time.time()
    with mp.Pool(args.workers, maxtasksperchild=1) as pool:
        for rec in pool.imap_unordered(work, [(i, args.max_nodes) for i in range(args.pairs)]):
            recs.append(rec)
            print(rec["index"], rec["result"], rec.get("cost"), rec["dp_cost"], rec["opened_nodes"], rec["seconds"], flush=True)
            recs.sort(key=lambda r: r["index"])
            json.dump({"partial": True, "pairs": recs}, open(args.out + ".partial", "w"), indent=1)
    out = {
        "what": "restated reference A* run to completion on pairs 0..%d of workloads.read_pairs(length=150), sample_tsa_config, dna-n" % (args.pairs - 1),
        "generator": "tests/golden/make_astar_c2.py", "host": platform.processor() or platform.machine(), "workers": args.workers,
        "wall_seconds": round(time.time() - t0, 1), "max_nodes": args.max_nodes, "pairs": recs,
    }
    json.dump(out, open(args.out, "w"), indent=1)
    os.remove(args.out + ".partial")


if __name__ == "__main__":
    main()
