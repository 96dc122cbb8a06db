"""Pins the CPU oracle (oracle/) against the reference's own known answers.  CPU only."""
import pytest

from oracle import oracle, tsa_config
from helpers import config_from_dict, ops_from_json, ops_from_toml, parse_config_any


def test_tsnax_disc1_473(configs, kats):
    # lib_tsalign/src/tests.rs:38-194: cost 10 under every min-length strategy.
    k = kats["tsnax_disc1_473"]
    flat = oracle.FlatConfig(tsa_config.parse(configs[k["config"]], k["alphabet"]))
    for lookahead in (False, True):
        r = oracle.astar_align(flat, k["reference"], k["query"], tuple(k["range"]), total_length_maximise=False,
                               min_length_lookahead=lookahead)
        assert r.found and r.cost == k["cost"]
        cost, er, eq, ok = oracle.rescore(flat, k["reference"], k["query"], r.ops, k["range"][0], k["range"][2])
        assert ok and cost == k["cost"] and (er, eq) == (k["range"][1], k["range"][3])
    # The comment at tests.rs:87 records the alignment the reference found (M/S notation).
    ours = r.cigar().replace("[-]:[-]", "[0,0]:[0,0]").replace("=", "M").replace("X", "S")
    # (the comment is after extend_beyond_range: 3M -> 165M on the left, one more M on the right)
    core = k["sample_cigar_comment"][len("165M"):-len("1M")]
    assert ours == "3M" + core


def test_tsnax_dp_matches(configs, kats):
    k = kats["tsnax_disc1_473"]
    flat = oracle.FlatConfig(tsa_config.parse(configs[k["config"]], k["alphabet"]))
    r = oracle.dp_align(flat, k["reference"], k["query"], tuple(k["range"]))
    assert r.found and r.cost == k["cost"]


def test_match_overtakes_gap(kats):
    # lib_tsalign/src/a_star_aligner/tests.rs:10-29 (plain gap-affine A*): the same graph is the TS graph with
    # template switches disabled and a base-agnostic primary table.
    k = kats["match_overtakes_gap"]
    cfg = tsa_config.rust_default(k["alphabet"])
    cfg.tables[0] = tsa_config.base_agnostic(k["alphabet"], "Primary Edit Costs", k["match"], k["substitution"], k["gap_open"], k["gap_extend"])
    flat = oracle.FlatConfig(cfg)
    for fn in (oracle.astar_align, oracle.dp_align):
        r = fn(flat, k["reference"], k["query"], None, no_ts=True)
        assert r.found and r.cost == k["cost"]
        assert oracle.rescore(flat, k["reference"], k["query"], r.ops)[0] == k["cost"]
    r = oracle.astar_align(flat, k["reference"], k["query"], None, no_ts=True)
    assert r.cigar() == k["cigar"]


@pytest.mark.parametrize("which", ["start", "end"])
def test_compute_cost_vectors(kats, which):
    # template_switch_specifics.rs:863-1410: compute_cost == closed-form sum of table entries.
    k = kats["compute_cost"]
    flat = oracle.FlatConfig(config_from_dict(k["config"]))
    blk = k[which]
    for v in blk["vectors"]:
        ops = ops_from_json(v["alignment"])
        cost, er, eq, ok = oracle.rescore(flat, blk["reference"], blk["query"], ops, *blk["offsets"])
        assert ok and cost == v["cost"], (v, cost)


def test_golden_toml_rescoring(configs, toml_golden):
    # The 8 committed result files: alignment + sequences rescore to the recorded cost under the config named
    # in SURVEY.md section 4; the walk ends exactly at the end of both sequences or the recorded range.
    for name, g in toml_golden.items():
        p = g["parsed"]
        seqs = p["sequences"]
        flat = oracle.FlatConfig(parse_config_any(configs[g["config"]]))
        if p["type"] != "WithTarget":
            assert p["result"]["astar_result_type"] == "ExceededCostLimit"
            continue
        ops = ops_from_toml(p["alignment"])
        cost, er, eq, ok = oracle.rescore(flat, seqs["reference"], seqs["query"], ops, p["reference_offset"], p["query_offset"])
        assert ok, name
        assert cost == int(p["cost"]) == p["result"]["cost"], (name, cost, p["cost"])
        # reverse complements recorded in the file pin the ACGT complement map
        from oracle import alphabets
        assert alphabets.reverse_complement("dna-n", seqs["reference"]) == seqs["reference_rc"]
        assert alphabets.reverse_complement("dna-n", seqs["query"]) == seqs["query_rc"]


def test_config_parser_rejects_stale_format(configs):
    # test_files/config/indel/config.tsa is a pre-v2 file (rr_cost ...) that the reference parser rejects.
    with pytest.raises(tsa_config.ConfigError):
        tsa_config.parse(configs["indel"], "dna-n")
    for name in ("sample", "bench", "experiments", "small", "range"):
        cfg = parse_config_any(configs[name])
        assert len(cfg.fns) == 6 and len(cfg.tables) == 5


def test_sample_config_values(configs):
    # SURVEY.md appendix B, sample_tsa_config/config.tsa.
    cfg = tsa_config.parse(configs["sample"], "dna-n")
    assert cfg.base == [3, 2, 2, 3, 3, 2, 2, 3]
    assert cfg.min_length == 5
    assert [cfg.evaluate(2, x) for x in (5, 6, 7, 8, 99)] == [5, 3, 1, 0, 0]
    assert cfg.evaluate(2, 100) == tsa_config.INF and cfg.evaluate(2, 4) == tsa_config.INF
    assert cfg.evaluate(4, 0) == 0 and cfg.evaluate(4, 1) == tsa_config.INF
    assert cfg.evaluate(5, -10**9) == 0
