#!/usr/bin/env python3
"""Measured lines for the BASELINE.json shapes that are not the headline of bench.py (configs[2..4]); one JSON object per
shape on stdout.  `value` = device-resident kernels (StagedBatch.run), `e2e` = the C ABI call on host buffers.

    python tools/bench_shapes.py [c4] [c5] [c3] [--pairs N]
C4: 10 kb pairs, --no-ts, with alignments.  C5: one 230 147 bp pair (human/chimp-like divergence), --no-ts, with alignment.
C3: 1 kb pairs, five planted template switches, flank lengths 50 / 50."""
import ctypes as C
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import template_switch_aligner_b200 as tsa  # noqa: E402
from template_switch_aligner_b200 import _lib, api, workloads  # noqa: E402


def c5_pair(n_len=230147, seed=5):
    rnd = random.Random(seed)
    ref = "".join(rnd.choice("ACGT") for _ in range(n_len))
    out, i = [], 0
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
    return ref, "".join(out)


def measure(name, aligner, pairs, steps=3, extra=None):
    lib = aligner._lib
    cells = sum(len(r) * len(q) for r, q in pairs)
    staged = tsa.StagedBatch(aligner, pairs)
    staged.run()
    t = time.perf_counter()
    fill_ms = jump_ms = 0.0
    for _ in range(steps):
        staged.run()
        tm = staged.timing()
        fill_ms += tm["fill_ms"]; jump_ms += tm["jump_ms"]
    dt = (time.perf_counter() - t) / steps
    res = staged.fetch()
    staged.close()
    bad = [r for r in res if not r.found]
    arr, keep = api._make_pairs(pairs)
    opt = api._options(aligner.no_ts, aligner.device, None, None, traceback=aligner.traceback, postprocess=0, dev_flags=aligner.dev_flags)
    err = C.create_string_buffer(512)
    best = None
    for _ in range(steps):
        out = (_lib.TsaResult * len(pairs))()
        t = time.perf_counter()
        rc = lib.tsa_align_batch(aligner.config._h, C.byref(opt), arr, len(pairs), out, err, len(err))
        e = time.perf_counter() - t
        assert rc == 0, err.value
        lib.tsa_results_free(out, len(pairs))
        best = e if best is None else min(best, e)
    s16, s32 = C.c_double(), C.c_double()
    lib.tsa_measure_addmin_peak(aligner.device, s16, s32)
    line = {"shape": name, "pairs": len(pairs), "cells": cells, "value_gcups": cells / dt / 1e9, "ms_per_step": dt * 1e3,
            "e2e_gcups": cells / best / 1e9, "e2e_ms": best * 1e3, "not_found": len(bad), "cost0": res[0].cost,
            "template_switches": sum(r.template_switches for r in res), "fill_ms": fill_ms / steps, "jump_ms": jump_ms / steps,
            "addmin_peak_s32_T": s32.value / 1e12, "addmin_peak_s16x2_T": s16.value / 1e12}
    if aligner.no_ts and fill_ms > 0:
        # 7 add-min lane operations per cell for the gap-affine fill
        line["roofline"] = {"kernel": "k_affine_wave", "achieved_Taddmin_s": 7.0 * cells / (fill_ms / steps * 1e-3) / 1e12, "peak_Taddmin_s": s32.value / 1e12,
                            "frac": 7.0 * cells / (fill_ms / steps * 1e-3) / s32.value}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--") and not a.isdigit()]
    shapes = args or ["c4", "c5", "c3"]
    n_pairs = int(sys.argv[sys.argv.index("--pairs") + 1]) if "--pairs" in sys.argv else None
    text = workloads.sample_config_text()
    dev_flags = int(sys.argv[sys.argv.index("--dev-flags") + 1]) if "--dev-flags" in sys.argv else 0   # 128: one warp per pair in the primary fill; 32: fused first window stage; 256: --no-ts strips in (pair, strip) ticket order
    if "c4" in shapes:
        pairs = [workloads.long_pair(i, 10000) for i in range(n_pairs or 1024)]
        for tb in (False, True):
            measure(f"c4: 10 kb pairs, --no-ts, alignments={tb}", tsa.Aligner(costs=text, no_ts=True, traceback=tb, dev_flags=dev_flags), pairs, extra={"dev_flags": dev_flags})
    if "c5" in shapes:
        pair = c5_pair()
        for tb in (False, True):
            measure(f"c5 shape: one {len(pair[0])} x {len(pair[1])} pair, --no-ts, alignment={tb}, one GPU", tsa.Aligner(costs=text, no_ts=True, traceback=tb), [pair])
    if "medium" in shapes:
        # pairs of the "medium" jump class (545 .. 1055 characters): column windows of 544 columns, template switches on, no flanks
        pairs = [workloads.long_pair(200 + i, 800, sub_rate=0.004, indel_rate=0.002, n_tsm=3) for i in range(n_pairs or 256)]
        measure("medium: 800 bp pairs, 3 planted TSMs, sample config", tsa.Aligner(costs=text, dev_flags=dev_flags), pairs, steps=2, extra={"dev_flags": dev_flags})
    if "c3" in shapes:
        fl = 50
        ftext = text.replace("left_flank_length = 0", f"left_flank_length = {fl}").replace("right_flank_length = 0", f"right_flank_length = {fl}")
        pairs = [workloads.long_pair(i, 1000, indel_rate=0.0, n_tsm=5) for i in range(n_pairs or 16)]
        measure("c3: 1 kb pairs, 5 planted TSMs, flanks 50/50", tsa.Aligner(costs=ftext, dev_flags=dev_flags), pairs, steps=2, extra={"dev_flags": dev_flags})
    return 0


if __name__ == "__main__":
    sys.exit(main())
