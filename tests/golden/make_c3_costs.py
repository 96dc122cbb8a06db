#!/usr/bin/env python3
"""Pre-computes the CPU oracle's optimal costs (scalar layered DP with flank planes, oracle/dp_oracle.cpp) for BASELINE config 3
at its named shape: tests/golden/c3_costs.json {"<length>|<index>|<n_tsm>": cost}.  Pairs:
template_switch_aligner_b200.workloads.long_pair(index, length, n_tsm=n_tsm) (1 % substitutions, 0.5 % indels, planted reverse
template switches), cost model = sample_tsa_config with left_flank_length = right_flank_length = 50.  1 kb pairs take tens of
minutes on one core; the 300-500 bp pairs straddle several 64 x 64 tiles of k_flank_fused:
    python tests/golden/make_c3_costs.py
"""
import json
import multiprocessing as mp
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = [(300, 100, 2), (330, 101, 2), (420, 102, 3), (500, 103, 3), (190, 104, 1), (257, 105, 2)] + [(1000, i, 5) for i in range(8)]
OUT = os.path.join(HERE, "c3_costs.json")


def config_text():
    from template_switch_aligner_b200 import workloads
    return workloads.sample_config_text().replace("left_flank_length = 0", "left_flank_length = 50").replace("right_flank_length = 0", "right_flank_length = 50")


def work(case):
    from oracle import oracle, tsa_config
    from template_switch_aligner_b200 import workloads
    length, index, n_tsm = case
    flat = oracle.FlatConfig(tsa_config.parse(config_text(), "dna-n"))
    r, q = workloads.long_pair(index, length, n_tsm=n_tsm)
    res = oracle.dp_align(flat, r, q)
    return case, (res.cost if res.found else None)


def main():
    done = json.load(open(OUT)) if os.path.exists(OUT) else {}
    todo = [c for c in CASES if "%d|%d|%d" % c not in done]
    with mp.Pool(min(int(os.environ.get("WORKERS", "3")), max(1, len(todo)))) as pool:
        for case, cost in pool.imap_unordered(work, todo):
            done["%d|%d|%d" % case] = cost
            json.dump(done, open(OUT, "w"), indent=1, sort_keys=True)
            print(case, cost, flush=True)


if __name__ == "__main__":
    main()
