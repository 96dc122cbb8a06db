#!/usr/bin/env python3
"""Developer probe: align one golden TOML case of tests/golden/toml_golden.json on the GPU and dump the alignment."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle
from helpers import ops_from_toml, parse_config_any
import template_switch_aligner_b200 as tsa
name = sys.argv[1]
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
G = json.load(open(os.path.join(ROOT, "tests/golden/toml_golden.json")))
configs = json.load(open(os.path.join(ROOT, "tests/golden/configs.json")))
g = G[name]; p = g["parsed"]; seqs = p["sequences"]
ocfg = parse_config_any(configs[g["config"]]); flat = oracle.FlatConfig(ocfg)
cost, er, eq, ok = oracle.rescore(flat, seqs["reference"], seqs["query"], ops_from_toml(p["alignment"]), p["reference_offset"], p["query_offset"])
rng = (p["reference_offset"], er, p["query_offset"], eq)
a = tsa.Aligner(costs=configs[g["config"]], alphabet=ocfg.alphabet, dev_flags=flags)
res = a.align_batch([(seqs["reference"], seqs["query"], rng)])[0]
print(json.dumps({"name": name, "config": g["config"], "range": rng, "status": res.status, "message": res.message, "cost": res.cost, "want": p["cost"], "ops": res.ops, "cigar": tsa.cigar_of(res.ops) if res.ops else None,
                  "n": len(seqs["reference"]), "m": len(seqs["query"])}))
