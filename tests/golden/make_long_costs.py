#!/usr/bin/env python3
"""Pre-computes the CPU oracle's optimal costs (scalar layered DP, oracle/dp_oracle.cpp) for the long synthetic pairs of
the column-window GPU tests: tests/golden/long_costs.json {"<length>|<index>|<n_tsm>": cost}.  The pairs come from
template_switch_aligner_b200.workloads.long_pair(index, length, sub_rate=0.004, indel_rate=0.002, n_tsm=n_tsm) with the
sample cost model; each takes minutes to tens of minutes on one core:    python tests/golden/make_long_costs.py
"""
import json
import multiprocessing as mp
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = [(600, 3, 2), (700, 4, 3), (900, 5, 3), (1100, 6, 3), (1300, 7, 4), (1600, 8, 4), (2200, 9, 4), (3000, 41, 4)]
OUT = os.path.join(HERE, "long_costs.json")


def work(case):
    from oracle import oracle, tsa_config
    from template_switch_aligner_b200 import workloads
    length, index, n_tsm = case
    flat = oracle.FlatConfig(tsa_config.parse(workloads.sample_config_text(), "dna-n"))
    r, q = workloads.long_pair(index, length, sub_rate=0.004, indel_rate=0.002, n_tsm=n_tsm)
    res = oracle.dp_align(flat, r, q)
    return case, (res.cost if res.found else None)


def main():
    done = json.load(open(OUT)) if os.path.exists(OUT) else {}
    todo = [c for c in CASES if "%d|%d|%d" % c not in done]
    with mp.Pool(min(6, max(1, len(todo)))) as pool:
        for case, cost in pool.imap_unordered(work, todo):
            done["%d|%d|%d" % case] = cost
            json.dump(done, open(OUT, "w"), indent=1, sort_keys=True)
            print(case, cost, flush=True)


if __name__ == "__main__":
    main()
