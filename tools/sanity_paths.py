#!/usr/bin/env python3
"""One small call through every kernel family (meant to run under compute-sanitizer on a B200): python tools/sanity_paths.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import template_switch_aligner_b200 as tsa
from template_switch_aligner_b200 import workloads

text = workloads.sample_config_text()
ftext = text.replace("left_flank_length = 0", "left_flank_length = 5").replace("right_flank_length = 0", "right_flank_length = 4")


def run(label, aligner, pairs):
    res = aligner.align_batch(pairs)
    print(label, [(g.status, g.cost, g.template_switches) for g in res], flush=True)
    assert all(g.status == 0 for g in res), label


reads = workloads.read_pairs(6)
run("reads", tsa.Aligner(costs=text), reads)
run("reads, tiled stage only", tsa.Aligner(costs=text, dev_flags=64), reads)
run("reads, flanks", tsa.Aligner(costs=ftext), reads[:3])
medium = [workloads.long_pair(20 + k, 600 + 90 * k, sub_rate=0.004, indel_rate=0.002, n_tsm=2) for k in range(2)]
run("medium (windows of 544, row queue)", tsa.Aligner(costs=text), medium)
run("medium, fused windows", tsa.Aligner(costs=text, dev_flags=32), medium)
run("medium, flanks (multi-warp fill + flank tiles)", tsa.Aligner(costs=ftext), medium[:1])
long_ = [workloads.long_pair(60, 1500, sub_rate=0.0005, indel_rate=0.0, n_tsm=1), workloads.long_pair(7, 1300, sub_rate=0.004, indel_rate=0.002, n_tsm=4)]
run("long (windows of 544 / 1056 / tiled)", tsa.Aligner(costs=text), long_)
nots = [workloads.long_pair(300 + k, 900 + 300 * k) for k in range(4)]
run("--no-ts, code matrix", tsa.Aligner(costs=text, no_ts=True, dev_flags=16), nots)
run("--no-ts, checkpoints", tsa.Aligner(costs=text, no_ts=True, dev_flags=8), nots)
run("--no-ts, costs only", tsa.Aligner(costs=text, no_ts=True, traceback=False), nots)
print("sanity ok")
