"""Kernel logic on the CPU: the product sources compiled against the SIMT emulator (tests/emul) must agree with the
oracle.  These tests exercise exactly the code the GPU runs (same .cuh), minus the hardware."""
import pytest

from oracle import oracle
from helpers import parse_config_any
from emul_lib import emul
import parity
import template_switch_aligner_b200 as tsa


def test_random_models_emulated():
    n_ts = parity.random_model_batches(emul(), range(0, 60), max_len=14)
    assert n_ts > 10


def test_random_models_with_flanks_emulated():
    # left / right flank lengths 0..2: flank planes, three-state seeds of the ordinary plane, plane-aware traceback
    n_ts = parity.random_model_batches(emul(), range(1000, 1040), max_len=12, flanks=True)
    assert n_ts > 5


def test_range_config_emulated(configs, pairs):
    # test_files/config/range: flank lengths 5 / 5 (the only shipped config with flanks)
    ocfg = parse_config_any(configs["range"])
    assert (ocfg.left_flank_length, ocfg.right_flank_length) == (5, 5)
    flat = oracle.FlatConfig(ocfg)
    items = parity.test_file_pairs(pairs, ocfg.alphabet, 24)
    aligner = tsa.Aligner(costs=configs["range"], alphabet=ocfg.alphabet, lib=emul())
    parity.check_batch(aligner, flat, [(r, q) for _, r, q in items], label="range")


@pytest.mark.parametrize("cfg_name,max_len", [("sample", 45), ("bench", 40), ("small", 40)])
def test_test_files_emulated(configs, pairs, cfg_name, max_len):
    ocfg = parse_config_any(configs[cfg_name])
    flat = oracle.FlatConfig(ocfg)
    items = parity.test_file_pairs(pairs, ocfg.alphabet, max_len)
    assert len(items) >= 15
    for no_ts in (False, True):
        aligner = tsa.Aligner(costs=configs[cfg_name], alphabet=ocfg.alphabet, no_ts=no_ts, lib=emul())
        parity.check_batch(aligner, flat, [(r, q) for _, r, q in items], no_ts=no_ts, label=cfg_name)


def test_tsnax_kat_emulated(configs, kats):
    # lib_tsalign/src/tests.rs:38-194 through the emulated kernels: 416 x 416 bp (jump kernel class C = 17)
    k = kats["tsnax_disc1_473"]
    aligner = tsa.Aligner(costs=configs[k["config"]], alphabet=k["alphabet"], lib=emul())
    res = aligner.align_batch([(k["reference"], k["query"], tuple(k["range"]))])[0]
    assert res.found and res.cost == k["cost"]


def test_edge_cases_emulated(configs):
    ocfg = parse_config_any(configs["sample"])
    flat = oracle.FlatConfig(ocfg)
    aligner = tsa.Aligner(costs=configs["sample"], lib=emul())
    cases = [("", ""), ("A", ""), ("", "ACGT"), ("ACGTN", "NNNNN"), ("ACGT" * 8, "ACGT" * 8, (3, 3, 5, 5)), ("ACGTACGTAC", "ACGTTCGTAC", (0, 10, 0, 10))]
    parity.check_batch(aligner, flat, cases, label="edge")
    # invalid characters / ranges are per-pair input errors (align.rs:389-405), the rest of the batch still runs
    res = aligner.align_batch([("ACGX", "ACGT"), ("ACGT", "ACGT", (3, 2, 0, 4)), ("ACGT", "ACGT")])
    assert [r.status for r in res] == [7, 8, 0] and res[2].cost == 0
    # cost limit semantics of the exact fill
    res = aligner.align_batch([("ACGTACGTAC", "ACGTTCGTAC")], cost_limit=1)[0]
    assert res.result_type == "ExceededCostLimit" and res.cost == 1
    res = aligner.align_batch([("ACGTACGTAC", "ACGTTCGTAC")], cost_limit=2)[0]
    assert res.found and res.cost == 2


def test_python_surface_emulated(configs):
    # mirror of python_bindings/python/tsalign/__init__.py: align() returns None when no target within limits
    a = tsa.align("ACGTACGTAC", "ACGTTCGTAC", costs=configs["sample"], lib=emul())
    assert a is not None and a.cost == 2 and a.stats()["result"]["astar_result_type"] == "FoundTarget"
    assert tsa.align("ACGTACGTAC", "ACGTTCGTAC", costs=configs["sample"], cost_limit=0, lib=emul()) is None
    with pytest.raises(ValueError):
        tsa.Aligner(costs="x", costs_file="y", lib=emul())
    with pytest.raises(ValueError):
        tsa.Aligner(min_length_strategy="bogus", lib=emul())


def test_no_ts_multi_strip_emulated(configs):
    # k_affine_wave: pairs wider than one 256-column strip (boundary column + progress flags between warps), with
    # ranges that do not start at a strip boundary, ragged lengths, and a batch that mixes one- and many-strip pairs
    from template_switch_aligner_b200 import workloads
    ocfg = parse_config_any(configs["sample"])
    flat = oracle.FlatConfig(ocfg)
    cases = []
    for idx, length in enumerate((255, 256, 257, 300, 520, 700)):
        r, q = workloads.long_pair(idx, length, sub_rate=0.03, indel_rate=0.02)
        cases.append((r, q))
    r, q = workloads.long_pair(7, 600, sub_rate=0.02, indel_rate=0.01)
    cases.append((r, q, (13, len(r) - 5, 40, len(q))))
    cases.append((r[:40], q, (0, 40, 0, len(q))))          # short reference, three strips of query
    cases.append((r, q[:33]))
    cases.append(("ACGT", "ACGA"))
    for tb in (True, False):
        aligner = tsa.Aligner(costs=configs["sample"], no_ts=True, traceback=tb, lib=emul())
        parity.check_batch(aligner, flat, cases, no_ts=True, label="wave")


def _narrow_model(text, off, ld, lmax):
    """The sample cost model with small offset / length-difference / length hulls, so that column windows fit 96 columns."""
    text = text.replace("RQQROffset\n -inf -100 101", f"RQQROffset\n -inf -{off} {off + 1}").replace("RRQQOffset\n -inf -100 101", f"RRQQOffset\n -inf -{off} {off + 1}")
    text = text.replace("LengthDifference\n -inf -100 101", f"LengthDifference\n -inf -{ld} {ld + 1}")
    text = text.replace("   0 5 6 7 8 100\n", f"   0 5 6 7 8 {lmax + 1}\n")
    assert f"-{off} {off + 1}" in text and f"8 {lmax + 1}" in text
    return text


def test_column_windows_emulated():
    # k_ts_jump<C, true> / k_traceback<C, true>: the emulator build runs pairs wider than 48 characters on 96-column windows
    # (second stage 160 columns) so that the window arithmetic -- offsets, band vectors of the fill, staging of table
    # windows, overflow flag and second stage, windowed traceback -- is exercised at sizes the oracle finishes quickly
    from template_switch_aligner_b200 import workloads
    from oracle import tsa_config
    base = workloads.sample_config_text()
    total_ts = 0
    for off, ld, lmax, count, length, thr in ((10, 8, 20, 4, 130, 0), (20, 5, 16, 3, 170, 5)):
        text = _narrow_model(base, off, ld, lmax)
        flat = oracle.FlatConfig(tsa_config.parse(text, "dna-n"))
        pairs = []
        for k in range(count):
            r, q = workloads.read_pair(900 + 31 * k + off, length)
            # plant a short reverse-complement copy within reach of the narrow offsets
            p0 = 40 + 7 * k
            q = q[:p0] + workloads.revcomp(r[p0 + 2:p0 + 2 + 12]) + q[p0 + 12:]
            pairs.append((r, q))
        pairs.append((pairs[0][0], pairs[0][1], (10, len(pairs[0][0]) - 3, 12, len(pairs[0][1]))))
        aligner = tsa.Aligner(costs=text, alphabet="dna-n", dev_flags=4, first_threshold=thr, lib=emul())
        total_ts += parity.check_batch(aligner, flat, pairs, label=f"windows {off}/{ld}/{lmax}")
    assert total_ts >= 3
    # a cost model whose offset / length-difference hulls alone exceed the lane grid (+-100 against the 160 columns of this
    # build's widest windows) is refused loudly, never answered wrongly
    wide = tsa.Aligner(costs=base, alphabet="dna-n", dev_flags=4, lib=emul())
    res = wide.align_batch(workloads.read_pairs(1, start=500, length=200))[0]
    assert res.status == 9 and "windows" in res.message


def test_tiled_windows_emulated():
    # Third window stage: the entrance columns of a chain pair are cut into sub-ranges, each evaluated on its own windows, the
    # seeds combined through atomicMin (no pair is refused for its width).  (1) dev_flags=64: every pair wider than 31 runs ONLY
    # that stage with sub-ranges of at most 24 columns, under random cost models; (2) pairs whose bands overflow the 96- and the
    # 160-column windows of this build and reach the stage the regular way.
    from template_switch_aligner_b200 import workloads
    from oracle import tsa_config
    n_ts = parity.random_model_batches(emul(), range(0, 7), max_len=56, pairs_per_model=3, dev_flags=64, min_len=33)
    assert n_ts >= 3
    text = _narrow_model(workloads.sample_config_text(), 20, 5, 16)
    flat = oracle.FlatConfig(tsa_config.parse(text, "dna-n"))
    pairs = []
    for k in range(1):
        r, q = workloads.long_pair(70 + k, 420, sub_rate=0.01, indel_rate=0.004, n_tsm=0)
        for t in range(3):
            p0 = 40 + 90 * t + 7 * k
            q = q[:p0] + workloads.revcomp(r[p0 + 2:p0 + 2 + 12]) + q[p0 + 12:]
        pairs.append((r, q))
    aligner = tsa.Aligner(costs=text, alphabet="dna-n", dev_flags=4, lib=emul())
    assert parity.check_batch(aligner, flat, pairs, label="tiled windows") == 1
    # ragged shapes through the pipelined primary fill (32-column blocks in this build: 1 .. 5 blocks, fewer chunks than warps,
    # more blocks than warps) and the tiled stage
    r, q = workloads.read_pair(77, 150)
    ragged = [(r[:5], q), (r, q[:5]), ("", q[:40]), (r[:40], ""), (r[:33], q[:33]), (r[:16], q[:140]), (r[:140], q[:17])]
    aligner = tsa.Aligner(costs=text, alphabet="dna-n", dev_flags=64, lib=emul())
    parity.check_batch(aligner, flat, ragged, label="ragged")


def test_flank_tiles_emulated():
    # k_flank_fused: pairs longer than one 64 x 64 tile (halo exchange through the planes in global memory) and flank lengths
    # above the 8 planes of one launch; costs against the oracle, alignments walked through the per-plane codes
    from template_switch_aligner_b200 import workloads
    from oracle import tsa_config
    base = workloads.sample_config_text()
    for lf, rf, length, count in ((5, 5, 100, 2), (12, 9, 140, 1)):
        text = base.replace("left_flank_length = 0", f"left_flank_length = {lf}").replace("right_flank_length = 0", f"right_flank_length = {rf}")
        flat = oracle.FlatConfig(tsa_config.parse(text, "dna-n"))
        pairs = workloads.read_pairs(count, start=40, length=length)
        pairs.append((pairs[0][0], pairs[0][1], (3, length - 2, 5, len(pairs[0][1]))))
        n_ts = parity.check_batch(tsa.Aligner(costs=text, lib=emul()), flat, pairs, label=f"flank tiles {lf}/{rf}")
        assert n_ts >= 1


def test_layer_cap_mixed_batch_emulated(configs):
    # max_template_switches below what some pairs of the batch need: those pairs are refused (status 9), every answered pair
    # still carries the proven optimum -- the cap of one pair must not end the deepening rounds of its batch siblings
    from template_switch_aligner_b200 import workloads
    ocfg = parse_config_any(configs["sample"])
    flat = oracle.FlatConfig(ocfg)
    pairs = workloads.read_pairs(5, start=11, length=44) + [("ACGTTGCAAGGCTA" * 3, "ACGTTGCAAGGCTA" * 3), ("ACGTACGTACGTAAGT", "ACGTACGTTCGTAAGT")]
    want = [oracle.dp_align(flat, r, q).cost for r, q in pairs]
    seen = set()
    for mts in (1, 2, 3, 64):
        for thr in (3, 12):
            aligner = tsa.Aligner(costs=configs["sample"], lib=emul(), max_template_switches=mts, first_threshold=thr)
            res = aligner.align_batch(pairs)
            for g, w in zip(res, want):
                assert g.status in (0, 9), g.message
                if g.status == 0:
                    assert g.found and g.cost == w, (mts, thr, g.cost, w)
                seen.add((mts, g.status))
            if mts == 64:
                assert all(g.status == 0 for g in res)
            assert res[5].status == 0 and res[5].cost == 0
    assert (1, 9) in seen and (1, 0) in seen


def test_descendant_strategy_emulated(configs, pairs):
    # --ts-descendant-strategy allow-only-all-equal (strategies/descendant.rs:22-104): all template switches of an alignment share
    # their primary.  Expected = the better of the oracle's optima under the cost model with only reference-primary kinds and with
    # only query-primary kinds (base cost of the other kinds = inf).
    import re
    text = configs["sample"]

    def only(primary):
        other = "q" if primary == "r" else "r"
        return re.sub(rf"^({other}[rq][fr]_cost\s*=\s*)\S+", r"\1inf", text, flags=re.M)

    ocfg = parse_config_any(text)
    flats = [oracle.FlatConfig(parse_config_any(only(p))) for p in "rq"]
    items = parity.test_file_pairs(pairs, ocfg.alphabet, 40)[:12]
    items += [("two", "ACGTTGCATGCAAGTCCGATAGGCTTACGATC", "ACGTTGCAACTTGCATCCGATAGAAGCCTATC")]
    aligner = tsa.Aligner(costs=text, alphabet=ocfg.alphabet, lib=emul(), descendant_strategy="allow-only-all-equal")
    free = tsa.Aligner(costs=text, alphabet=ocfg.alphabet, lib=emul())
    got = aligner.align_batch([(r, q) for _, r, q in items])
    unconstrained = free.align_batch([(r, q) for _, r, q in items])
    differs = 0
    for (name, r, q), g, u in zip(items, got, unconstrained):
        want = min(oracle.dp_align(f, r, q).cost for f in flats)
        assert g.status == 0 and g.found and g.cost == want, (name, g.cost, want)
        assert g.cost >= u.cost
        differs += g.cost != u.cost
        primaries = {op[2] for op in g.ops if op[1] == oracle.OP_TS_ENTRANCE}
        assert len(primaries) <= 1, (name, g.ops)
        flat = oracle.FlatConfig(ocfg)
        cost, er, eq, ok = oracle.rescore(flat, r, q, [oracle.Op(*o) for o in g.ops], g.range[0], g.range[2])
        assert ok and cost == g.cost
