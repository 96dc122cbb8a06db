#!/usr/bin/env python3
"""Developer probe (not part of the product or of bench.py): times the hot path on the other BASELINE.json shapes so
that the next optimisation target is chosen from measurements.

    python tools/probe.py e2e [pairs]        phase breakdown of tsa_align_batch on C2 read pairs (TSA_B200_DEBUG=1)
    python tools/probe.py c4 [pairs] [len]   --no-ts long pairs: costs only, then with traceback
    python tools/probe.py c3 [pairs] [len]   1 kb pairs with 5 planted TSMs, flank lengths 50/50
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import template_switch_aligner_b200 as tsa  # noqa: E402
from template_switch_aligner_b200 import _lib, api, workloads  # noqa: E402


def timed_align(aligner, pairs, reps=3, label=""):
    arr, keep = api._make_pairs(pairs)
    n = len(pairs)
    cells = sum(len(r) * len(q) for r, q in pairs)
    err = C.create_string_buffer(512)
    opt = api._options(aligner.no_ts, aligner.device, None, None, first_threshold=aligner.first_threshold, traceback=aligner.traceback, scout=aligner.scout)
    best = None
    for rep in range(reps):
        res = (_lib.TsaResult * max(1, n))()
        t = time.perf_counter()
        rc = aligner._lib.tsa_align_batch(aligner.config._h, C.byref(opt), arr, n, res, err, len(err))
        dt = time.perf_counter() - t
        assert rc == 0, err.value
        bad = [i for i in range(n) if res[i].status != 0]
        cost0 = res[0].cost
        aligner._lib.tsa_results_free(res, n)
        best = dt if best is None else min(best, dt)
        print(f"{label} rep {rep}: {dt * 1e3:.1f} ms, {cells / dt / 1e9:.3f} GCUPS, {n / dt:.0f} pairs/s, bad={len(bad)}, cost[0]={cost0}", flush=True)
    return best


def main():
    mode = sys.argv[1]
    if mode == "e2e":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
        pairs = workloads.read_pairs(n)
        os.environ["TSA_B200_DEBUG"] = "1"
        aligner = tsa.Aligner(costs=workloads.sample_config_text(), alphabet="dna-n")
        timed_align(aligner, pairs, 3, "c2 e2e")
    elif mode == "c4":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
        length = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
        pairs = [workloads.long_pair(i, length) for i in range(n)]
        for tb in (False, True):
            aligner = tsa.Aligner(costs=workloads.sample_config_text(), alphabet="dna-n", no_ts=True, traceback=tb)
            timed_align(aligner, pairs, 3, f"c4 no-ts len={length} traceback={tb}")
    elif mode == "c3":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
        length = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
        fl = int(sys.argv[4]) if len(sys.argv) > 4 else 50
        text = workloads.sample_config_text().replace("left_flank_length = 0", f"left_flank_length = {fl}").replace("right_flank_length = 0", f"right_flank_length = {fl}")
        pairs = [workloads.long_pair(i, length, indel_rate=0.0, n_tsm=5) for i in range(n)]
        os.environ["TSA_B200_DEBUG"] = "1"
        aligner = tsa.Aligner(costs=text, alphabet="dna-n")
        timed_align(aligner, pairs, 2, f"c3 len={length} flanks={fl}")
    elif mode == "c5":
        # BASELINE config 5 shape without template switches: one 230 147 x 236 216 pair on one GPU
        n_len = int(sys.argv[2]) if len(sys.argv) > 2 else 230147
        import random
        rnd = random.Random(5)
        ref = "".join(rnd.choice("ACGT") for _ in range(n_len))
        out = []
        i = 0
        while i < n_len:
            x = rnd.random()
            if x < 0.0015:
                i += 1 + int(rnd.expovariate(1 / 4.0)); continue
            if x < 0.003:
                out.extend(rnd.choice("ACGT") for _ in range(1 + int(rnd.expovariate(1 / 4.0))))
            c = ref[i]
            if rnd.random() < 0.012:
                c = rnd.choice([b for b in "ACGT" if b != c])
            out.append(c); i += 1
        qry = "".join(out)
        print("lengths", len(ref), len(qry), flush=True)
        os.environ["TSA_B200_DEBUG"] = "1"
        for tb in (False, True):
            aligner = tsa.Aligner(costs=workloads.sample_config_text(), alphabet="dna-n", no_ts=True, traceback=tb)
            timed_align(aligner, [(ref, qry)], 2, f"c5 no-ts traceback={tb}")
    elif mode == "longts":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
        length = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
        pairs = [workloads.long_pair(i, length, sub_rate=0.004, indel_rate=0.002, n_tsm=4) for i in range(n)]
        os.environ["TSA_B200_DEBUG"] = "1"
        aligner = tsa.Aligner(costs=workloads.sample_config_text(), alphabet="dna-n")
        timed_align(aligner, pairs, 2, f"long ts len={length}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
