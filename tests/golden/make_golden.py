#!/usr/bin/env python3
"""Generates the committed golden fixtures under tests/golden/ from /root/reference.

Run once in the build container (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

Outputs (all small JSON):
  configs.json ....... the config.tsa texts shipped with the reference (sample + test_files/config/*)
  pairs.json ......... every test_files/*.fa with exactly two records and < 8 kB (raw record text, not yet
                       skip-character-filtered), keyed by file name
  kats.json .......... the reference's own known-answer tests for the alignment path:
                       - lib_tsalign/src/tests.rs:38-194          (TSNAX-DISC1_473, bench config, cost 10)
                       - lib_tsalign/src/a_star_aligner/tests.rs:10-29   (AGT/GTCC, 1D2=2I, cost 9)
                       - .../alignment/template_switch_specifics.rs:863-1247 (10 alignments + closed-form costs)
                       - tsalign-tests/tests/integration.rs:6-29  (CLI smoke lines)
  toml_golden.json ... the 8 committed result files test_files/*.toml (raw text + parsed)
"""
import json
import os
import re
import tomllib

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def read(p):
    with open(os.path.join(REF, p)) as f:
        return f.read()


def parse_fasta(text):
    recs = []
    for ln in text.splitlines():
        if ln.startswith(">"):
            recs.append([ln[1:], ""])
        elif recs:
            recs[-1][1] += ln.strip()
    return recs


def main():
    # ---- configs -------------------------------------------------------------------------------------------
    configs = {"sample": read("sample_tsa_config/config.tsa")}
    cdir = os.path.join(REF, "test_files/config")
    for name in sorted(os.listdir(cdir)):
        p = os.path.join(cdir, name, "config.tsa")
        if os.path.exists(p):
            configs[name] = open(p).read()
    json.dump(configs, open(os.path.join(OUT, "configs.json"), "w"), indent=0, sort_keys=True)

    # ---- pairs ---------------------------------------------------------------------------------------------
    pairs = {}
    tdir = os.path.join(REF, "test_files")
    for name in sorted(os.listdir(tdir)):
        if not name.endswith(".fa"):
            continue
        text = open(os.path.join(tdir, name)).read()
        if len(text) > 8000:
            continue
        recs = parse_fasta(text)
        if len(recs) != 2:
            continue
        pairs[name] = {"raw": text, "records": recs}
    json.dump(pairs, open(os.path.join(OUT, "pairs.json"), "w"), indent=0, sort_keys=True)

    # ---- KATs ----------------------------------------------------------------------------------------------
    tests_rs = read("lib_tsalign/src/tests.rs")
    m_ref = re.search(r'let reference = VectorGenome::<DnaAlphabet>::from_slice_u8\(\s*b"([ACGT]+)"', tests_rs)
    m_qry = re.search(r'let query = VectorGenome::from_slice_u8\(\s*b"([ACGT]+)"', tests_rs)
    assert m_ref and m_qry
    assert "AlignmentCoordinates::new(196, 196)" in tests_rs and "AlignmentCoordinates::new(219, 212)" in tests_rs
    assert "assert_eq!(result.statistics().cost, r64(10.0));" in tests_rs
    kats = {
        "tsnax_disc1_473": {
            "source": "lib_tsalign/src/tests.rs:38-194",
            "alphabet": "dna", "config": "bench",
            "reference": m_ref.group(1), "query": m_qry.group(1),
            "range": [196, 219, 196, 212],  # ref offset, ref limit, qry offset, qry limit
            "cost": 10,
            "total_length_strategy": "none",
            "sample_cigar_comment": "165M[TSQQR:[0,0]:[0,0]:21:5M1D1M1I3M:17]2M1S1M",
        },
        "match_overtakes_gap": {
            "source": "lib_tsalign/src/a_star_aligner/tests.rs:10-29",
            "alphabet": "dna", "reference": "AGT", "query": "GTCC",
            "match": 0, "substitution": 2, "gap_open": 4, "gap_extend": 1,
            "cigar": "1D2=2I", "cost": 9,
        },
    }

    # compute_cost vectors (template_switch_specifics.rs:863-1247).  The config there is built in code:
    # base costs 10^k, secondary fwd sub 3 / open [3,6,6,6] / ext 1, secondary rev sub 5 / open [3,7,7,7] / ext 1,
    # primary sub 2 / open 3 / ext 1, flanks zero, and six step functions f(i) = mult * (i + 21) on -20..=20
    # (Length on 0..=20) with multipliers 17,17,19,23,29,31.
    def fn(mult, lo):
        return [[i, mult * (i + 21)] for i in range(lo, 21)]

    cc_cfg = {
        "alphabet": "dna", "left_flank_length": 0, "right_flank_length": 0,
        "base": [10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000],  # rrf rqf qrf qqf rrr rqr qrr qqr
        "fns": [fn(17, -20), fn(17, -20), fn(19, 0), fn(23, -20), fn(29, -20), fn(31, -20)],
        "tables": [
            {"match": 0, "sub": 2, "open": [3, 3, 3, 3], "ext": [1, 1, 1, 1]},
            {"match": 0, "sub": 3, "open": [3, 6, 6, 6], "ext": [1, 1, 1, 1]},
            {"match": 0, "sub": 5, "open": [3, 7, 7, 7], "ext": [1, 1, 1, 1]},
            {"match": 0, "sub": 0, "open": [0, 0, 0, 0], "ext": [0, 0, 0, 0]},
            {"match": 0, "sub": 0, "open": [0, 0, 0, 0], "ext": [0, 0, 0, 0]},
        ],
    }
    rqr = 1000000
    oc = lambda o: 17 * (o + 21)
    apg = lambda a: 31 * (a + 21)
    lc = lambda l: 19 * (l + 21)
    ldc = lambda d: 23 * (d + 21)
    sub_r = 5

    def ts(first_offset, inner, apg_v):
        ops = [[1, "TemplateSwitchEntrance", {"first_offset": first_offset, "primary": "Reference", "secondary": "Query", "direction": "Reverse"}]]
        ops += inner
        ops.append([1, "TemplateSwitchExit", {"anti_primary_gap": apg_v}])
        return ops

    M, SM, SS = "PrimaryMatch", "SecondaryMatch", "SecondarySubstitution"
    start = [
        ([[6, M]] + ts(-6, [[2, SM]], 2) + [[2, M]], rqr + oc(-6) + apg(2) + lc(2) + ldc(0)),
        ([[5, M]] + ts(-4, [[3, SM]], 3) + [[2, M]], rqr + oc(-4) + apg(3) + lc(3) + ldc(0)),
        ([[4, M]] + ts(-2, [[4, SM]], 4) + [[2, M]], rqr + oc(-2) + apg(4) + lc(4) + ldc(0)),
        ([[3, M]] + ts(0, [[1, SS], [4, SM]], 5) + [[2, M]], rqr + oc(0) + sub_r + apg(5) + lc(5) + ldc(0)),
        ([[2, M]] + ts(2, [[2, SS], [4, SM]], 6) + [[2, M]], rqr + oc(2) + 2 * sub_r + apg(6) + lc(6) + ldc(0)),
    ]
    end = [
        ([[1, M]] + ts(10, [[2, SM]], 2) + [[6, M]], rqr + oc(10) + apg(2) + lc(2) + ldc(0)),
        ([[1, M]] + ts(10, [[3, SM]], 3) + [[5, M]], rqr + oc(10) + apg(3) + lc(3) + ldc(0)),
        ([[1, M]] + ts(10, [[4, SM]], 4) + [[4, M]], rqr + oc(10) + apg(4) + lc(4) + ldc(0)),
        ([[1, M]] + ts(10, [[4, SM], [1, SS]], 5) + [[3, M]], rqr + oc(10) + sub_r + apg(5) + lc(5) + ldc(0)),
        ([[1, M]] + ts(10, [[4, SM], [2, SS]], 6) + [[2, M]], rqr + oc(10) + 2 * sub_r + apg(6) + lc(6) + ldc(0)),
    ]
    src = read("lib_tsalign/src/a_star_aligner/alignment_result/alignment/template_switch_specifics.rs")
    assert 'static START_REFERENCE: &[u8] = b"AGAGAGCTCTAA";' in src and 'static START_QUERY: &[u8] = b"AGAGAGCTTTAA";' in src
    assert 'static END_REFERENCE: &[u8] = b"AACTCTAGAGAG";' in src and 'static END_QUERY: &[u8] = b"AATTCTAGAGAG";' in src
    kats["compute_cost"] = {
        "source": "lib_tsalign/src/a_star_aligner/alignment_result/alignment/template_switch_specifics.rs:863-1410",
        "config": cc_cfg,
        "start": {"offsets": [2, 2], "reference": "AGAGAGCTCTAA", "query": "AGAGAGCTTTAA", "vectors": [{"alignment": a, "cost": c} for a, c in start]},
        "end": {"offsets": [1, 1], "reference": "AACTCTAGAGAG", "query": "AATTCTAGAGAG", "vectors": [{"alignment": a, "cost": c} for a, c in end]},
    }
    kats["cli_smoke"] = {
        "source": "tsalign-tests/tests/integration.rs:6-29",
        "lines": [
            ["align", "-p", "test_files/twin_a.fa"],
            ["align", "-r", "test_files/reference_a.fa", "-q", "test_files/query_a.fa"],
            ["align", "-p", "test_files/twin_100_0.01.fa", "--cost-limit", "0"],
            ["align", "-p", "test_files/twin_100_0.01.fa", "--memory-limit", "1000"],
            ["align", "-p", "test_files/twin_embedded.fa", "--use-embedded-rq-ranges"],
        ],
        "reference_a.fa": read("test_files/reference_a.fa"),
        "query_a.fa": read("test_files/query_a.fa"),
    }
    integ = read("tsalign-tests/tests/integration.rs")
    for line in kats["cli_smoke"]["lines"]:
        for tok in line[1:]:
            assert tok in integ, tok
    json.dump(kats, open(os.path.join(OUT, "kats.json"), "w"), indent=0, sort_keys=True)

    # ---- golden TOML results -------------------------------------------------------------------------------
    # Which config each file was produced with (re-derived by rescoring).
    toml_cfg = {
        "twin_ari_chrX_146823507_146823598.toml": "sample",
        "twin_ari_chrX_146823507_146823598_no_ts.toml": "sample",
        "twin_heli_MDC1-AS1_10.toml": "experiments",
        "twin_heli_MDC1-AS1_10_no_ts.toml": "experiments",
        "twin_heli_linc01237_114.toml": "experiments",
        "twin_heli_linc01237_114_no_ts.toml": "experiments",
        "underscore.toml": "experiments",
        "underscore_no_ts.toml": "experiments",
    }
    golden = {}
    for name, cfg in toml_cfg.items():
        text = read("test_files/" + name)
        golden[name] = {"config": cfg, "raw": text, "parsed": tomllib.loads(text)}
    json.dump(golden, open(os.path.join(OUT, "toml_golden.json"), "w"), indent=0, sort_keys=True)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
