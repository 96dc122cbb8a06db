#!/usr/bin/env python3
"""Optimal costs (scalar DP oracle, oracle/dp_oracle.cpp) of the reference's result files whose recorded alignment is not
optimal under the cost model that rescoring assigns to them: tests/golden/toml_optima.json {"<file>": cost}.
The files are fixtures of `tsalign show`; nothing in the reference says which cost model produced them.
    python tests/golden/make_toml_optima.py        (about two minutes)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FILES = ["twin_ari_chrX_146823507_146823598.toml"]


def main():
    from oracle import oracle
    from helpers import ops_from_toml, parse_config_any
    golden = json.load(open(os.path.join(HERE, "toml_golden.json")))
    configs = json.load(open(os.path.join(HERE, "configs.json")))
    out = {}
    for name in FILES:
        g = golden[name]
        p = g["parsed"]
        seqs = p["sequences"]
        flat = oracle.FlatConfig(parse_config_any(configs[g["config"]]))
        cost, er, eq, ok = oracle.rescore(flat, seqs["reference"], seqs["query"], ops_from_toml(p["alignment"]), p["reference_offset"], p["query_offset"])
        assert ok and cost == int(p["cost"])
        res = oracle.dp_align(flat, seqs["reference"], seqs["query"], (p["reference_offset"], er, p["query_offset"], eq))
        assert res.found and res.cost <= cost
        out[name] = res.cost
        print(name, "recorded", cost, "optimum", res.cost, flush=True)
    json.dump(out, open(os.path.join(HERE, "toml_optima.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
