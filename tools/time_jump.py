#!/usr/bin/env python3
"""Developer tool: device time of the jump / fill kernels on C2 read pairs through the staged ABI (no result checks; used to time
experimental builds selected with TSA_B200_LIB).    python tools/time_jump.py [pairs] [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import template_switch_aligner_b200 as tsa  # noqa: E402
from template_switch_aligner_b200 import api, workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 0
aligner = tsa.Aligner(costs=workloads.sample_config_text(), alphabet="dna-n", traceback=False, first_threshold=thr)
b = api.StagedBatch(aligner, workloads.read_pairs(n))
out = []
for _ in range(reps):
    b.run()
    out.append(b.timing())
print(json.dumps({"lib": os.environ.get("TSA_B200_LIB", "default"), "pairs": n, "first_threshold": thr, "last": out[-1], "jump_ms": [round(o["jump_ms"], 2) for o in out], "stats": b.stats()}))
