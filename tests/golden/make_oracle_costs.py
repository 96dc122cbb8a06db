#!/usr/bin/env python3
"""Pre-computes the CPU oracle's answers for the larger parity cases so that the GPU tests do not spend their box
time in the scalar oracle:  tests/golden/oracle_costs.json  {"<config>|<ts|nots>|<file>": cost or null}.
Generated here by oracle.dp_align (the scalar layered DP, itself proven equal to the restated reference A* in
tests/test_dp_vs_astar.py):    python tests/golden/make_oracle_costs.py
"""
import json
import multiprocessing as mp
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = ["sample", "bench", "experiments", "small", "no_intra_forward_jump", "range"]


def work(task):
    from oracle import oracle
    from helpers import clean_record, parse_config_any
    cfg_name, no_ts, name = task
    configs = json.load(open(os.path.join(HERE, "configs.json")))
    pairs = json.load(open(os.path.join(HERE, "pairs.json")))
    ocfg = parse_config_any(configs[cfg_name])
    p = pairs[name]
    r = clean_record(p["records"][0][1].replace("|", ""))
    q = clean_record(p["records"][1][1].replace("|", ""))
    try:
        oracle.alphabets.encode(ocfg.alphabet, r), oracle.alphabets.encode(ocfg.alphabet, q)
    except ValueError:
        return task, "skip"
    flat = oracle.FlatConfig(ocfg)
    res = oracle.dp_align(flat, r, q, no_ts=no_ts)
    return task, (res.cost if res.found else None)


def main():
    pairs = json.load(open(os.path.join(HERE, "pairs.json")))
    tasks = []
    for cfg in CONFIGS:
        for name, p in pairs.items():
            L = max(len(p["records"][0][1]), len(p["records"][1][1]))
            if L > 1150:
                continue
            if cfg != "sample" and L > 560:
                continue
            if cfg == "range" and L > 130:
                continue   # 11 flank planes in the scalar oracle
            tasks.append((cfg, False, name))
            tasks.append((cfg, True, name))
    out = {}
    path = os.path.join(HERE, "oracle_costs.json")
    if os.path.exists(path):   # incremental: keep what is already there
        out = json.load(open(path))
        tasks = [t for t in tasks if f"{t[0]}|{'nots' if t[1] else 'ts'}|{t[2]}" not in out]
    with mp.Pool(os.cpu_count()) as pool:
        for (cfg, no_ts, name), cost in pool.imap_unordered(work, tasks):
            if cost != "skip":
                out[f"{cfg}|{'nots' if no_ts else 'ts'}|{name}"] = cost
    json.dump(out, open(os.path.join(HERE, "oracle_costs.json"), "w"), indent=0, sort_keys=True)
    print(len(out), "entries")


if __name__ == "__main__":
    main()
