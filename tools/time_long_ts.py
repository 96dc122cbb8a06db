#!/usr/bin/env python3
"""Times one long pair WITH template switches through the C ABI (column windows + tiled stage): python tools/time_long_ts.py [length] [n_tsm]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import template_switch_aligner_b200 as tsa
from template_switch_aligner_b200 import workloads

length = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
n_tsm = int(sys.argv[2]) if len(sys.argv) > 2 else 6
text = workloads.sample_config_text()
pair = workloads.long_pair(90, length, sub_rate=0.002, indel_rate=0.001, n_tsm=n_tsm)
al = tsa.Aligner(costs=text)
for rep in range(2):
    t0 = time.time()
    g = al.align_batch([pair])[0]
    print(f"length {length}: status {g.status} cost {g.cost} switches {g.template_switches} {time.time() - t0:.2f} s {g.message}", flush=True)
h = tsa.Aligner(costs=text, no_ts=True).align_batch([pair])[0]
print("no-ts cost", h.cost)
