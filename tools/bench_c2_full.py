#!/usr/bin/env python3
"""BASELINE config 2 at its full size through ONE tsa_align_batch call: python tools/bench_c2_full.py [--pairs 1048576] [--unique 65536]
The batch repeats `unique` seeded read pairs (generating 10^6 pairs in Python would dominate the run); the call chunks, encodes,
aligns on two engines, post-processes and returns every alignment."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import template_switch_aligner_b200 as tsa
from template_switch_aligner_b200 import _lib, api, workloads


def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    n, unique = arg("--pairs", 1 << 20), arg("--unique", 1 << 16)
    base = workloads.read_pairs(min(unique, n))
    pairs = (base * ((n + len(base) - 1) // len(base)))[:n]
    aligner = tsa.Aligner(costs=workloads.sample_config_text())
    lib = aligner._lib
    arr, keep = api._make_pairs(pairs)
    opt = api._options(False, 0, None, None, traceback=True, postprocess=api.POST_EXTEND_BEYOND_RANGE | api.POST_EQUAL_COST_RANGES)
    err = C.create_string_buffer(512)
    cells = sum(len(r) * len(q) for r, q in pairs)
    times = []
    for rep in range(2):
        res = (_lib.TsaResult * n)()
        t = time.perf_counter()
        rc = lib.tsa_align_batch(aligner.config._h, C.byref(opt), arr, n, res, err, len(err))
        times.append(time.perf_counter() - t)
        assert rc == 0, err.value
        bad = sum(1 for i in range(0, n, 997) if res[i].status != 0 or res[i].result_type != 0)
        same = all(res[i].cost == res[i % len(base)].cost for i in range(0, n, 1013))
        switches = sum(res[i].template_switches for i in range(0, n, 101))
        lib.tsa_results_free(res, n)
    best = min(times)
    print(json.dumps({"workload": f"configs[1] at full size: {n} read pairs of 150 bp ({len(base)} distinct, repeated) in ONE tsa_align_batch call, alignments + post-processing",
                      "pairs": n, "seconds": best, "first_call_seconds": times[0], "pairs_per_s": n / best, "e2e_gcups": cells / best / 1e9,
                      "sampled_not_found": bad, "repeats_agree": same, "sampled_switches": switches}))


if __name__ == "__main__":
    main()
