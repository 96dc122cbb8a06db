#!/usr/bin/env python3
"""The plain wavefront kernel on the C4 shape, costs only (profiling target): python tools/time_c4_costs.py [pairs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import template_switch_aligner_b200 as tsa
from template_switch_aligner_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pairs = [workloads.long_pair(i, 10000) for i in range(n)]
staged = tsa.StagedBatch(tsa.Aligner(costs=workloads.sample_config_text(), no_ts=True, traceback=False), pairs)
for _ in range(3):
    t = time.perf_counter()
    staged.run()
    dt = time.perf_counter() - t
    print(f"{n} pairs: {dt * 1e3:.1f} ms, {sum(len(r) * len(q) for r, q in pairs) / dt / 1e9:.0f} GCUPS", flush=True)
staged.close()
