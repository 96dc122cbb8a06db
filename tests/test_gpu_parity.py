"""Parity of the CUDA path (through the C ABI of libtsalign_b200.so) against the CPU oracle.  Needs a B200."""
import pytest

from oracle import oracle, tsa_config
from helpers import parse_config_any
import parity
import template_switch_aligner_b200 as tsa
from template_switch_aligner_b200 import _lib, api, workloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle_costs():
    # tests/golden/make_oracle_costs.py: the scalar oracle's answers, computed once in the build container
    from conftest import load_golden
    return load_golden("oracle_costs.json")


def _expected(oracle_costs, cfg_name, no_ts, items):
    return [oracle_costs[f"{cfg_name}|{'nots' if no_ts else 'ts'}|{name}"] for name, _, _ in items]


@pytest.fixture(scope="module")
def lib():
    lib = _lib.default()
    assert lib.tsa_device_count() >= 1, "no CUDA device: the GPU tests cannot run"
    assert b"sm_100a" in lib.tsa_version()
    return lib


def test_random_models_gpu(lib):
    n_ts = parity.random_model_batches(lib, range(0, 150), max_len=20, pairs_per_model=8)
    assert n_ts > 30


def test_random_models_with_flanks_gpu(lib):
    n_ts = parity.random_model_batches(lib, range(1000, 1100), max_len=20, pairs_per_model=8, flanks=True)
    assert n_ts > 20


def test_flank_config_gpu(lib, configs, pairs, oracle_costs):
    # test_files/config/range (flank lengths 5 / 5) on the test files; BASELINE config 3 shape (1 kb, flanks 50) at 2 pairs
    ocfg = parse_config_any(configs["range"])
    flat = oracle.FlatConfig(ocfg)
    items = parity.test_file_pairs(pairs, ocfg.alphabet, 130)
    aligner = tsa.Aligner(costs=configs["range"], alphabet=ocfg.alphabet, lib=lib)
    parity.check_batch(aligner, flat, [(r, q) for _, r, q in items], label="range", expected=_expected(oracle_costs, "range", False, items))
    text = workloads.sample_config_text().replace("left_flank_length = 0", "left_flank_length = 50").replace("right_flank_length = 0", "right_flank_length = 50")
    big = [workloads.long_pair(i, 1000, n_tsm=5) for i in range(2)]
    a = tsa.Aligner(costs=text, lib=lib)
    b = tsa.Aligner(costs=workloads.sample_config_text(), lib=lib)
    ra, rb = a.align_batch(big), b.align_batch(big)
    for x, y in zip(ra, rb):
        assert x.status == 0 and x.found and y.found and x.ops is not None
        assert x.cost >= y.cost    # sample flank tables cost at least the primary table: flanks can only add cost


@pytest.mark.parametrize("cfg_name,max_len", [("sample", 130), ("bench", 130), ("experiments", 110), ("small", 130), ("no_intra_forward_jump", 130)])
def test_test_files_gpu(lib, configs, pairs, oracle_costs, cfg_name, max_len):
    ocfg = parse_config_any(configs[cfg_name])
    flat = oracle.FlatConfig(ocfg)
    items = parity.test_file_pairs(pairs, ocfg.alphabet, max_len)
    assert len(items) >= 30
    for no_ts in (False, True):
        aligner = tsa.Aligner(costs=configs[cfg_name], alphabet=ocfg.alphabet, no_ts=no_ts, lib=lib)
        parity.check_batch(aligner, flat, [(r, q) for _, r, q in items], no_ts=no_ts, label=cfg_name,
                           expected=_expected(oracle_costs, cfg_name, no_ts, items))


def test_test_files_long_gpu(lib, configs, pairs, oracle_costs):
    # 200 .. 1055 bp pairs of test_files (jump kernel classes C = 9, 17, 33), sample config
    ocfg = parse_config_any(configs["sample"])
    flat = oracle.FlatConfig(ocfg)
    items = parity.test_file_pairs(pairs, ocfg.alphabet, 1055, min_len=131)
    assert len(items) >= 15
    for no_ts in (False, True):
        aligner = tsa.Aligner(costs=configs["sample"], alphabet=ocfg.alphabet, no_ts=no_ts, lib=lib)
        parity.check_batch(aligner, flat, [(r, q) for _, r, q in items], no_ts=no_ts, label="long",
                           expected=_expected(oracle_costs, "sample", no_ts, items))


def test_reference_kats_gpu(lib, configs, kats):
    k = kats["tsnax_disc1_473"]  # lib_tsalign/src/tests.rs:38-194
    aligner = tsa.Aligner(costs=configs[k["config"]], alphabet=k["alphabet"], lib=lib)
    res = aligner.align_batch([(k["reference"], k["query"], tuple(k["range"]))])[0]
    assert res.found and res.cost == k["cost"]
    flat = oracle.FlatConfig(parse_config_any(configs[k["config"]]))
    parity.check_alignment(flat, (k["reference"], k["query"], tuple(k["range"])), res, "tsnax")
    k = kats["match_overtakes_gap"]  # lib_tsalign/src/a_star_aligner/tests.rs:10-29
    cfg = tsa_config.rust_default(k["alphabet"])
    cfg.tables[0] = tsa_config.base_agnostic(k["alphabet"], "Primary Edit Costs", k["match"], k["substitution"], k["gap_open"], k["gap_extend"])
    from helpers import config_to_text
    aligner = tsa.Aligner(costs=config_to_text(cfg), alphabet=k["alphabet"], no_ts=True, lib=lib)
    res = aligner.align_batch([(k["reference"], k["query"])])[0]
    assert res.found and res.cost == k["cost"]
    assert tsa.cigar_of(res.ops) == k["cigar"]   # 1D2=2I: the reference's own tie-break happens to agree here


def _toml_optima():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "toml_optima.json")))


def test_golden_toml_costs_gpu(lib, configs, toml_golden):
    # The committed result files of the reference (test_files/*.toml): same sequences, same range (the offsets in
    # the file and the end point its alignment reaches) -> exactly the recorded optimal cost.
    from helpers import ops_from_toml
    checked = 0
    optima = _toml_optima()
    for name, g in toml_golden.items():
        p = g["parsed"]
        if p["type"] != "WithTarget":
            continue
        seqs = p["sequences"]
        ocfg = parse_config_any(configs[g["config"]])
        flat = oracle.FlatConfig(ocfg)
        cost, er, eq, ok = oracle.rescore(flat, seqs["reference"], seqs["query"], ops_from_toml(p["alignment"]), p["reference_offset"], p["query_offset"])
        assert ok and cost == int(p["cost"])
        rng = (p["reference_offset"], er, p["query_offset"], eq)
        aligner = tsa.Aligner(costs=configs[g["config"]], alphabet=ocfg.alphabet, no_ts="no_ts" in name, lib=lib)
        res = aligner.align_batch([(seqs["reference"], seqs["query"], rng)])[0]
        assert res.status == 0, (name, res.status, res.message)
        # The files are fixtures of `tsalign show`; nothing in the reference says which cost model produced them.  Under
        # the sample model the recorded alignment of the 1.1 kb pair (three switches, cost 6) is not optimal: two switches
        # with a free length difference of 100 cost 4 (rescored below with the compute_cost restatement, and equal to the
        # scalar DP oracle's optimum, tests/golden/toml_optima.json written by tests/golden/make_toml_optima.py).
        want = optima.get(name, int(p["cost"]))
        assert res.found and res.cost == want and want <= int(p["cost"]), (name, res.cost, p["cost"])
        parity.check_alignment(flat, (seqs["reference"], seqs["query"], rng), res, name)
        checked += 1
    assert checked >= 5


def test_read_pair_batch_gpu(lib):
    # C2-shaped workload: 150 bp read pairs with a planted template switch, default `tsalign align` cost model
    text = workloads.sample_config_text()
    flat = oracle.FlatConfig(tsa_config.parse(text, "dna-n"))
    pairs = workloads.read_pairs(96)
    aligner = tsa.Aligner(costs=text, lib=lib)
    n_ts = parity.check_batch(aligner, flat, pairs, label="reads")
    assert n_ts >= 48  # the planted TSMs are found


def test_c2_pairs_against_completed_astar_gpu(lib):
    # GPU == the reference's own algorithm at config 2 size: costs of the restated A* run to completion offline
    # (tests/golden/astar_c2.json, make_astar_c2.py; hard pairs open 10^8 nodes), alignments rescored
    from conftest import load_golden
    golden = load_golden("astar_c2.json")
    text = workloads.sample_config_text()
    flat = oracle.FlatConfig(tsa_config.parse(text, "dna-n"))
    recs = [rec for rec in golden["pairs"] if rec["result"] == "FoundTarget"]
    pairs = [workloads.read_pair(rec["index"], 150) for rec in recs]
    assert len(pairs) >= 32
    got = tsa.Aligner(costs=text, lib=lib).align_batch(pairs)
    for rec, p, g in zip(recs, pairs, got):
        assert g.status == 0 and g.found and g.cost == rec["cost"], (rec["index"], g.cost, rec["cost"])
        parity.check_alignment(flat, p, g, "astar c2")


def test_c3_shape_against_oracle_gpu(lib):
    # BASELINE config 3 at its named shape: 1 kb pairs with 5 planted switches (and 190-500 bp pairs that straddle several
    # 64 x 64 tiles of k_flank_fused) under flank lengths 50 / 50 -- 101 planes per layer.  Costs of the scalar DP oracle
    # computed offline (tests/golden/c3_costs.json, make_c3_costs.py); flank alignments rescored (TSA_FLAG_KEEP_FLANK_RUNS)
    from conftest import load_golden
    golden = load_golden("c3_costs.json")
    text = workloads.sample_config_text().replace("left_flank_length = 0", "left_flank_length = 50").replace("right_flank_length = 0", "right_flank_length = 50")
    flat = oracle.FlatConfig(tsa_config.parse(text, "dna-n"))
    cases, want = [], []
    for key, cost in sorted(golden.items()):
        length, index, n_tsm = map(int, key.split("|"))
        cases.append(workloads.long_pair(index, length, n_tsm=n_tsm)); want.append(cost)
    assert sum(len(c[0]) == 1000 for c in cases) >= 8 and len(cases) >= 12
    aligner = tsa.Aligner(costs=text, lib=lib)
    n_ts = parity.check_batch(aligner, flat, cases, label="c3", expected=want)
    assert n_ts >= 8


def test_full_size_properties_gpu(lib):
    # BASELINE config 2 at batch scale (no oracle): size-independent properties of the optimum
    text = workloads.sample_config_text()
    aligner = tsa.Aligner(costs=text, lib=lib)
    no_ts = tsa.Aligner(costs=text, no_ts=True, lib=lib)
    pairs = workloads.read_pairs(4096, start=10_000)
    a = aligner.align_batch(pairs)
    b = no_ts.align_batch(pairs)
    again = aligner.align_batch(list(reversed(pairs)))
    for x, y, z, (r, q) in zip(a, b, reversed(again), pairs):
        assert x.found and y.found
        assert x.cost <= y.cost                       # template switches can only help
        assert x.cost == z.cost                       # independent of batch position
        assert (x.cost == 0) == (r == q)              # zero cost iff identical (sample model: every edit costs)
    # identical pairs cost 0; a pair against itself reversed-complemented once is one template switch
    same = aligner.align_batch([(r, r) for r, _ in pairs[:64]])
    assert all(s.found and s.cost == 0 for s in same)
    # swapping reference and query keeps the optimum (the sample cost model is symmetric in R/Q)
    sw = aligner.align_batch([(q, r) for r, q in pairs[:256]])
    assert [s.cost for s in sw] == [x.cost for x in a[:256]]


def test_edge_cases_gpu(lib, configs):
    ocfg = parse_config_any(configs["sample"])
    flat = oracle.FlatConfig(ocfg)
    aligner = tsa.Aligner(costs=configs["sample"], lib=lib)
    cases = [("", ""), ("A", ""), ("", "ACGT"), ("ACGTN", "NNNNN"), ("ACGT" * 8, "ACGT" * 8, (3, 3, 5, 5)), ("ACGTACGTAC", "ACGTTCGTAC", (0, 10, 0, 10))]
    parity.check_batch(aligner, flat, cases, label="edge")
    res = aligner.align_batch([("ACGX", "ACGT"), ("ACGT", "ACGT", (3, 2, 0, 4)), ("ACGT", "ACGT")])
    assert [r.status for r in res] == [7, 8, 0] and res[2].cost == 0
    assert aligner.align_batch([]) == []
    r, q = workloads.long_pair(0, 1500)
    nots = tsa.Aligner(costs=configs["sample"], no_ts=True, lib=lib)
    want = oracle.dp_align(flat, r, q, no_ts=True)
    assert nots.align_batch([(r, q)])[0].cost == want.cost


def test_no_ts_long_gpu(lib):
    # BASELINE config 4 shape (--no-ts, long pairs): k_affine_wave with many strips per pair and many pairs in flight
    # (strips of one pair on different SMs, boundary columns through L2), ragged lengths, ranges; alignments rescored
    text = workloads.sample_config_text()
    flat = oracle.FlatConfig(tsa_config.parse(text, "dna-n"))
    pairs = [workloads.long_pair(i, 700 + 37 * i, sub_rate=0.02, indel_rate=0.01) for i in range(48)]
    pairs += [workloads.long_pair(100 + i, 2500) for i in range(3)]
    r, q = workloads.long_pair(200, 3000)
    pairs.append((r, q, (100, len(r) - 7, 321, len(q) - 1)))
    expected = [oracle.dp_align(flat, p[0], p[1], p[2] if len(p) > 2 else None, no_ts=True).cost for p in pairs]
    for tb in (True, False):
        nots = tsa.Aligner(costs=text, no_ts=True, traceback=tb, lib=lib)
        parity.check_batch(nots, flat, pairs, no_ts=True, label="wave", expected=expected)
    # one 10 kb pair (40 strips): cost against the oracle, alignment rescored
    r, q = workloads.long_pair(0, 10000)
    nots = tsa.Aligner(costs=text, no_ts=True, lib=lib)
    parity.check_batch(nots, flat, [(r, q)], no_ts=True, label="wave10k")


def _long_case(length, index, n_tsm):
    return workloads.long_pair(index, length, sub_rate=0.004, indel_rate=0.002, n_tsm=n_tsm)


def test_column_windows_gpu(lib):
    # Pairs wider than 544 characters run the jump kernel on column windows (k_ts_jump<17, true>, second stage whole
    # sequences up to 1055 characters, 1056-column windows beyond).  (1) medium pairs: the windowed path must return the
    # same optimum as the whole-sequence path (dev_flags=2), which the oracle tests pin; (2) the precomputed oracle costs
    # of tests/golden/long_costs.json (scalar DP, minutes per pair on a CPU); (3) every alignment rescored.
    import json
    import os
    text = workloads.sample_config_text()
    flat = oracle.FlatConfig(tsa_config.parse(text, "dna-n"))
    win = tsa.Aligner(costs=text, lib=lib)
    plain = tsa.Aligner(costs=text, dev_flags=2, lib=lib)
    medium = [_long_case(600 + 45 * k, 20 + k, 2 + k % 3) for k in range(10)]
    a = win.align_batch(medium)
    b = plain.align_batch(medium)
    assert all(x.status == 0 and y.status == 0 for x, y in zip(a, b)), [(x.status, x.message) for x in a]
    assert [x.cost for x in a] == [y.cost for y in b]
    assert sum(x.template_switches > 0 for x in a) >= 5
    for p, g in zip(medium, a):
        parity.check_alignment(flat, p, g, "windows medium")
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "long_costs.json")))
    cases, want = [], []
    for key, cost in sorted(golden.items()):
        length, index, n_tsm = map(int, key.split("|"))
        cases.append(_long_case(length, index, n_tsm)); want.append(cost)
    assert len(cases) >= 4
    # every pair gets an answer: up to 1055 characters the second stage is the whole sequences; beyond, windows of 1056 columns
    # and then the tiled stage (sub-ranges of the entrance columns).  The answer must be the oracle's optimum.
    got = win.align_batch(cases)
    for p, g, w in zip(cases, got, want):
        assert g.status == 0, (len(p[0]), g.status, g.message)
        assert g.found and g.cost == w, (len(p[0]), g.cost, w)
        parity.check_alignment(flat, p, g, "windows golden")
    # long pairs beyond any whole-sequence class.  With the sample cost model a template switch may shift the diagonal by
    # +-100 columns for free, so the band of cheap cells grows with every layer: pairs whose optimum needs few switches
    # fit the 1056-column windows, the others reach the tiled stage.
    easy = [workloads.long_pair(60 + k, 1500 + 900 * k, sub_rate=0.0005, indel_rate=0.0, n_tsm=1) for k in range(3)]
    res = win.align_batch(easy)
    nots = tsa.Aligner(costs=text, no_ts=True, lib=lib).align_batch(easy)
    for p, g, h in zip(easy, res, nots):
        assert g.status == 0 and g.found, (g.status, g.message)
        assert g.cost <= h.cost
        parity.check_alignment(flat, p, g, "windows long")
    assert sum(g.template_switches > 0 for g in res) >= 2
    hard = [_long_case(3000, 41, 4)]
    g = win.align_batch(hard)[0]
    assert g.status == 0 and g.found, (g.status, g.message)
    if "3000|41|4" in golden:
        assert g.cost == golden["3000|41|4"]
    parity.check_alignment(flat, hard[0], g, "windows hard")


def test_tiled_windows_gpu(lib):
    # The tiled window stage on its own (dev_flags=64: every pair wider than 31 runs only that stage, sub-ranges of at most 24
    # entrance columns): random cost models against the oracle, the read pairs of config 2 against the completed A* costs, and
    # a 4 kb pair with template switches (no pair is refused for its width; tools/time_long_ts.py runs 10 kb), rescored.
    n_ts = parity.random_model_batches(lib, range(0, 40), max_len=64, pairs_per_model=6, dev_flags=64, min_len=33)
    assert n_ts > 15
    from conftest import load_golden
    golden = load_golden("astar_c2.json")
    text = workloads.sample_config_text()
    flat = oracle.FlatConfig(tsa_config.parse(text, "dna-n"))
    recs = [rec for rec in golden["pairs"] if rec["result"] == "FoundTarget"]
    pairs = [workloads.read_pair(rec["index"], 150) for rec in recs]
    got = tsa.Aligner(costs=text, dev_flags=64, lib=lib).align_batch(pairs)
    for rec, p, g in zip(recs, pairs, got):
        assert g.status == 0 and g.found and g.cost == rec["cost"], (rec["index"], g.cost, rec["cost"], g.message)
        parity.check_alignment(flat, p, g, "tiled c2")
    big = [workloads.long_pair(90, 4000, sub_rate=0.002, indel_rate=0.001, n_tsm=6)]
    g = tsa.Aligner(costs=text, lib=lib).align_batch(big)[0]
    h = tsa.Aligner(costs=text, no_ts=True, lib=lib).align_batch(big)[0]
    assert g.status == 0 and g.found, (g.status, g.message)
    assert g.cost < h.cost and g.template_switches >= 3
    parity.check_alignment(flat, big[0], g, "4 kb with template switches")


def test_single_long_pair_no_ts_gpu(lib):
    # BASELINE config 5 shape at reduced size (--no-ts).  A 12 kb pair (47 strips, all in flight at once) against the scalar
    # oracle; a 60 kb pair (235 strips; too large for the oracle's matrices) through size-independent properties: the alignment
    # rescoring to the cost, symmetry under swapping the sequences (the sample model is symmetric), and additivity when the
    # pair is cut at a match of the returned alignment (two ranged alignments of the same pair).
    text = workloads.sample_config_text()
    flat = oracle.FlatConfig(tsa_config.parse(text, "dna-n"))
    nots = tsa.Aligner(costs=text, no_ts=True, lib=lib)
    r, q = workloads.long_pair(5, 12000, sub_rate=0.012, indel_rate=0.003)
    parity.check_batch(nots, flat, [(r, q), (r, q, (1234, len(r) - 77, 999, len(q) - 5))], no_ts=True, label="wave12k")
    r, q = workloads.long_pair(6, 60000, sub_rate=0.012, indel_rate=0.003)
    whole, swapped = nots.align_batch([(r, q), (q, r)])
    assert whole.found and swapped.found and whole.cost == swapped.cost
    parity.check_alignment(flat, (r, q), whole, "wave60k")
    # walk to the first match at or after the middle of the reference
    i = j = 0
    cut = None
    for count, t, *_ in whole.ops:
        for _ in range(count):
            if t == 3 and i >= len(r) // 2 and cut is None:
                cut = (i, j)
            i += t in (1, 2, 3)
            j += t in (0, 2, 3)
    assert cut is not None
    left, right = nots.align_batch([(r, q, (0, cut[0], 0, cut[1])), (r, q, (cut[0], len(r), cut[1], len(q)))])
    assert left.found and right.found and left.cost + right.cost == whole.cost


def test_postprocess_through_the_abi_gpu(lib):
    # tsa_options.postprocess on the GPU path == the host-only entry applied to the searched alignment; with the extension the
    # reported range grows over the matching flanks of a ranged pair and the cost does not change
    text = workloads.sample_config_text()
    pairs = workloads.read_pairs(64, start=700)
    ranged = [(r, q, (10, len(r) - 10, 10, len(q) - 10)) for r, q in pairs[:16] if r[:10] == q[:10]]
    plain = tsa.Aligner(costs=text, lib=lib)
    post = tsa.Aligner(costs=text, postprocess=api.POST_EXTEND_BEYOND_RANGE | api.POST_EQUAL_COST_RANGES, lib=lib)
    a = plain.align_batch(pairs + ranged)
    b = post.align_batch(pairs + ranged)
    n_ranges = 0
    for p, x, y in zip(pairs + ranged, a, b):
        assert x.found and y.found
        ops, ranges, rng, cost = api.postprocess(plain.config, p[0], p[1], x.ops, x.range, api.POST_EXTEND_BEYOND_RANGE | api.POST_EQUAL_COST_RANGES)
        assert (ops, ranges, rng) == (y.ops, y.equal_cost_ranges, y.range)
        assert cost <= x.cost and y.cost == x.cost
        n_ranges += sum(e is not None for e in y.equal_cost_ranges)
        if len(p) > 2:
            assert y.range[0] < p[2][0] or p[0][p[2][0] - 1] != p[1][p[2][2] - 1]
    assert n_ranges >= 32


def test_no_ts_random_models_long_gpu(lib):
    # k_affine_wave under random cost models (asymmetric tables, infinite entries, N characters), pairs of several strips with
    # random ranges: optimum against the scalar oracle, alignments rescored
    import random
    import randcfg
    from helpers import config_to_text
    checked = 0
    for seed in range(300, 312):
        rng = random.Random(seed)
        cfg = randcfg.random_config(rng)
        flat = oracle.FlatConfig(cfg)
        cases = []
        for k in range(4):
            r, q = randcfg.random_pair(rng, max_len=rng.choice([300, 520, 900]), n_weight=0.02)
            cases.append((r, q, randcfg.random_range(rng, r, q)))
        aligner = tsa.Aligner(costs=config_to_text(cfg), no_ts=True, lib=lib)
        parity.check_batch(aligner, flat, cases, no_ts=True, label=f"wave seed {seed}")
        checked += len(cases)
    assert checked == 48


def test_cli_binary_batch_gpu(lib, tmp_path):
    # the real `tsalign-b200` binary (linked against libtsalign_b200.so) on a B200: batch front-end and single-pair mode against
    # the committed A* costs of the read-pair sample and the library's own answers
    import json
    import os
    import subprocess
    from conftest import load_golden
    cli = os.path.join(os.path.dirname(os.path.abspath(_lib.LIB_PATH)), "tsalign-b200")
    assert os.path.exists(cli), "template_switch_aligner_b200/tsalign-b200 is not built (python -c 'import __graft_entry__ as g; g.build()')"
    os.makedirs(tmp_path / "cfg")
    (tmp_path / "cfg" / "config.tsa").write_text(workloads.sample_config_text())
    gold = {p["index"]: p for p in load_golden("astar_c2.json")["pairs"] if p["result"] == "FoundTarget"}
    ids = sorted(gold)
    with open(tmp_path / "reads.fa", "w") as fh:
        for i in ids:
            r, q = workloads.read_pair(i, 150)
            fh.write(f">r{i} reference\n{r}\n>q{i} query\n{q}\n")
    out = subprocess.run([cli, "align", "--pairs", "reads.fa", "-c", "cfg", "--output-jsonl", "reads.jsonl"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    recs = [json.loads(ln) for ln in open(tmp_path / "reads.jsonl")]
    assert [r["cost"] for r in recs] == [gold[i]["cost"] for i in ids] and all(r["result"] == "FoundTarget" for r in recs)
    want = tsa.Aligner(costs=workloads.sample_config_text(), device=0, postprocess=api.POST_EXTEND_BEYOND_RANGE | api.POST_EQUAL_COST_RANGES).align_batch([workloads.read_pair(i, 150) for i in ids])
    assert [r["template_switches"] for r in recs] == [w.template_switches for w in want]
    # single-pair mode, TOML file
    r, q = workloads.read_pair(ids[0], 150)
    (tmp_path / "one.fa").write_text(f">ref\n{r}\n>qry\n{q}\n")
    one = subprocess.run([cli, "align", "-p", "one.fa", "-c", "cfg", "-o", "one.toml"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert one.returncode == 0 and f"Reached target with cost {gold[ids[0]]['cost']}" in one.stdout, one.stderr
    assert "astar_result_type = \"FoundTarget\"" in (tmp_path / "one.toml").read_text()


def test_cli_toml_bytes_against_golden_gpu(lib, configs, toml_golden, tmp_path):
    # Byte-level layout of the TOML result file against the reference's committed result files (test_files/*.toml, written by
    # toml 0.9 / noisy_float through align/template_switch_distance_type_selectors.rs:442-449): the real CLI binary is run on the
    # file's sequences and range; every line that does not depend on the search statistics must be identical text, and where the
    # returned alignment equals the recorded one the `alignment = [...]` line (inline tables of the entrances / exits) as well.
    import os
    import subprocess
    import tomllib
    from helpers import ops_from_toml
    cli = os.path.join(os.path.dirname(os.path.abspath(_lib.LIB_PATH)), "tsalign-b200")
    assert os.path.exists(cli)
    volatile = ("duration_seconds", "opened_nodes", "closed_nodes", "suboptimal_opened_nodes", "suboptimal_opened_nodes_ratio", "runtime", "memory")

    def fmt_alignment(al):
        out = []
        for count, op in al:
            if isinstance(op, str):
                out.append(f'[{count}, "{op}"]')
            elif "TemplateSwitchEntrance" in op:
                e = op["TemplateSwitchEntrance"]
                r = e["equal_cost_range"]
                out.append(f'[{count}, {{ TemplateSwitchEntrance = {{ first_offset = {e["first_offset"]}, equal_cost_range = {{ min_start = {r["min_start"]}, max_start = {r["max_start"]}, '
                           f'min_end = {r["min_end"]}, max_end = {r["max_end"]} }}, primary = "{e["primary"]}", secondary = "{e["secondary"]}", direction = "{e["direction"]}" }} }}]')
            else:
                out.append(f'[{count}, {{ TemplateSwitchExit = {{ anti_primary_gap = {op["TemplateSwitchExit"]["anti_primary_gap"]} }} }}]')
        return "alignment = [" + ", ".join(out) + "]"

    compared_alignments = compared_files = 0
    for name, g in toml_golden.items():
        p = g["parsed"]
        if p["type"] != "WithTarget" or name.startswith("twin_ari"):      # (twin_ari: see DESIGN.md section 2, not an optimum of the sample model)
            continue
        seqs = p["sequences"]
        ocfg = parse_config_any(configs[g["config"]])
        flat = oracle.FlatConfig(ocfg)
        _, er, eq, ok = oracle.rescore(flat, seqs["reference"], seqs["query"], ops_from_toml(p["alignment"]), p["reference_offset"], p["query_offset"])
        assert ok
        d = tmp_path / name.replace(".toml", "")
        os.makedirs(d / "cfg")
        (d / "cfg" / "config.tsa").write_text(configs[g["config"]])
        rn, qn = seqs["reference_name"], seqs["query_name"]       # "<id> <comment>" (align.rs:418-419)
        (d / "pair.fa").write_text(f">{rn.rstrip(' ') if rn.endswith(' ') and ' ' not in rn[:-1] else rn}\n{seqs['reference']}\n>{qn.rstrip(' ') if qn.endswith(' ') and ' ' not in qn[:-1] else qn}\n{seqs['query']}\n")
        args = [cli, "align", "-p", "pair.fa", "-c", "cfg", "-a", ocfg.alphabet, "-o", "out.toml", "--dont-extend-beyond-range",
                "--rq-ranges", f"R{p['reference_offset']}..{er}Q{p['query_offset']}..{eq}"]
        if "no_ts" in name:
            args.append("--no-ts")
        run = subprocess.run(args, cwd=d, capture_output=True, text=True, timeout=900)
        assert run.returncode == 0, (name, run.stderr)
        ours_raw = (d / "out.toml").read_text()
        ours = tomllib.loads(ours_raw)
        if int(ours["cost"]) != int(p["cost"]):
            continue                                                # (an optimum below the recorded cost: covered by test_golden_toml_costs_gpu)
        ours_lines, gold_lines = ours_raw.splitlines(), g["raw"].splitlines()
        assert [ln.split(" =")[0] for ln in ours_lines] == [ln.split(" =")[0] for ln in gold_lines], name      # same keys, same order, same sections
        for a, b in zip(ours_lines, gold_lines):
            key = a.split(" =")[0]
            if key in volatile:
                continue
            if key == "alignment":
                # one formatter must reproduce both lines byte for byte from their parsed values: same inline-table layout
                assert a == fmt_alignment(ours["alignment"]) and b == fmt_alignment(p["alignment"]), (name, a[:200], b[:200])
                compared_alignments += 1
                if ours["alignment"] == p["alignment"]:
                    assert a == b
                continue
            assert a == b, (name, a[:120], b[:120])
        compared_files += 1
    assert compared_files >= 4 and compared_alignments >= 1, (compared_files, compared_alignments)
