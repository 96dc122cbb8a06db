"""The dense layered DP (oracle/dp_oracle.cpp, the statement the CUDA kernels implement) against the
restated reference A* (oracle/astar_oracle.cpp) -- same optimal cost on the reference's test files and on
random sequences / cost models that exercise the quirks of SURVEY.md appendix A.6.  CPU only."""
import random

import pytest

from oracle import oracle, tsa_config
from helpers import clean_record, parse_config_any
import randcfg


def _pair(p):
    r = clean_record(p["records"][0][1].replace("|", ""))
    q = clean_record(p["records"][1][1].replace("|", ""))
    return r, q


@pytest.mark.parametrize("cfg_name,max_len", [("sample", 45), ("bench", 45), ("experiments", 35), ("small", 45), ("range", 45)])
def test_test_files_small(configs, pairs, cfg_name, max_len):
    flat = oracle.FlatConfig(parse_config_any(configs[cfg_name]))
    n = 0
    for name, p in pairs.items():
        r, q = _pair(p)
        if max(len(r), len(q)) > max_len:
            continue
        try:
            a = oracle.astar_align(flat, r, q)
        except ValueError:
            continue  # characters outside dna-n
        d = oracle.dp_align(flat, r, q)
        assert (a.result_type, a.cost) == (d.result_type, d.cost), name
        for res in (a, d):
            cost, er, eq, ok = oracle.rescore(flat, r, q, res.ops)
            assert ok and (er, eq) == (len(r), len(q)), (name, res.cigar())
            if flat.cfg.left_flank_length == 0 and flat.cfg.right_flank_length == 0:
                assert cost == res.cost, (name, res.cigar())  # flanks: RLE is ambiguous, see test_random_models
        n += 1
    assert n >= 15


def test_no_ts_test_files(configs, pairs):
    flat = oracle.FlatConfig(tsa_config.parse(configs["sample"], "dna-n"))
    for name, p in pairs.items():
        r, q = _pair(p)
        if max(len(r), len(q)) > 130:
            continue
        try:
            a = oracle.astar_align(flat, r, q, no_ts=True)
        except ValueError:
            continue
        d = oracle.dp_align(flat, r, q, no_ts=True)
        assert (a.result_type, a.cost) == (d.result_type, d.cost), name
        assert oracle.rescore(flat, r, q, d.ops)[0] == d.cost


@pytest.mark.parametrize("flanks,seed0,count", [(False, 0, 120), (True, 1000, 80)])
def test_random_models(flanks, seed0, count):
    with_ts = 0
    for seed in range(seed0, seed0 + count):
        rng = random.Random(seed)
        flat = oracle.FlatConfig(randcfg.random_config(rng, flanks=flanks))
        for _ in range(4):
            r, q = randcfg.random_pair(rng, max_len=11)
            rg = randcfg.random_range(rng, r, q)
            no_ts = rng.random() < 0.1
            a = oracle.astar_align(flat, r, q, rg, no_ts=no_ts, min_length_lookahead=rng.random() < 0.5,
                                   total_length_maximise=rng.random() < 0.5)
            d = oracle.dp_align(flat, r, q, rg, no_ts=no_ts)
            assert (a.result_type, a.cost) == (d.result_type, d.cost), (seed, r, q, rg, a.cigar(), d.cigar())
            if d.found and not flanks:
                # (with flank lengths > 0 the RLE merges flank and non-flank ops, alignment_type.rs:101-121, and the
                # reference's compute_cost is todo!() there -- only costs and end points are compared)
                for res in (a, d):
                    cost, er, eq, ok = oracle.rescore(flat, r, q, res.ops, rg[0], rg[2], as_searched=True)
                    assert ok and cost == res.cost and (er, eq) == (rg[1], rg[3]), (seed, res.cigar())
                with_ts += any(o.type == oracle.OP_TS_ENTRANCE for o in d.ops)
    if not flanks:
        assert with_ts > 10


def test_cost_limit(configs):
    flat = oracle.FlatConfig(tsa_config.parse(configs["sample"], "dna-n"))
    r, q = "ACGTACGTAC", "ACGTTCGTAC"
    for fn in (oracle.astar_align, oracle.dp_align):
        assert fn(flat, r, q, cost_limit=1).result_type == "ExceededCostLimit"
        assert fn(flat, r, q, cost_limit=1).cost == 1
        # label-setting search (--ts-total-length-strategy none): limit == optimum is found
        res = fn(flat, r, q, cost_limit=2, total_length_maximise=False)
        assert res.found and res.cost == 2
    # Reference quirk (generic_a_star/src/lib.rs:349-363): the label-correcting search (default `maximise`) keeps
    # draining the open list after the first target and reports ExceededCostLimit if any successor was ever cut
    # by the limit, even though a target within the limit exists.  The DP reports FoundTarget.
    assert oracle.astar_align(flat, r, q, cost_limit=2).result_type == "ExceededCostLimit"
    assert oracle.dp_align(flat, r, q, cost_limit=2).found


def test_c2_pairs_against_completed_astar():
    """BASELINE config 2 at its named size: the restated reference A* was run to completion offline on seeded 150 bp read pairs
    (tests/golden/make_astar_c2.py; the hard ones open 10^8 nodes) -- the layered DP must return the same optimal costs."""
    from conftest import load_golden
    from template_switch_aligner_b200 import workloads
    golden = load_golden("astar_c2.json")
    flat = oracle.FlatConfig(tsa_config.parse(workloads.sample_config_text(), "dna-n"))
    checked = 0
    for rec in golden["pairs"]:
        if rec["result"] != "FoundTarget":
            continue
        r, q = workloads.read_pair(rec["index"], 150)
        assert (len(r), len(q)) == (rec["reference_len"], rec["query_len"])
        d = oracle.dp_align(flat, r, q)
        assert d.found and d.cost == rec["cost"], (rec["index"], d.cost, rec["cost"])
        checked += 1
    assert checked >= 32 and max(rec.get("opened_nodes", 0) for rec in golden["pairs"]) > 5e7   # hard pairs included
