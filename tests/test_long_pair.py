"""One long pair without template switches as column bands with checkpoints (csrc/tsa_band.cuh, tsa_long.cu): the emulated kernels
against the oracle -- several bands, several column groups and row blocks, tiny memory limits, ranges, edge shapes."""
import pytest

from oracle import oracle
from helpers import parse_config_any
from emul_lib import emul
import template_switch_aligner_b200 as tsa
from template_switch_aligner_b200 import api, workloads


def _check(flat, r, q, res, rng=None):
    ro, rl, qo, ql = rng if rng else (0, len(r), 0, len(q))
    want = oracle.dp_align(flat, r, q, no_ts=True, rng=(ro, rl, qo, ql)) if rng else oracle.dp_align(flat, r, q, no_ts=True)
    assert res.status == 0 and res.found and res.cost == want.cost, (res, want.cost)
    cost, er, eq, ok = oracle.rescore(flat, r, q, [oracle.Op(*o) for o in res.ops], ro, qo)
    assert ok and cost == res.cost and (er, eq) == (rl, ql)


@pytest.mark.parametrize("world,interval,group", [(1, 64, 1), (1, 100, 2), (2, 64, 1), (3, 37, 1), (2, 1000, 2), (1, 0, 0)])
def test_bands_emulated(configs, world, interval, group):
    flat = oracle.FlatConfig(parse_config_any(configs["sample"]))
    aligner = tsa.Aligner(costs=configs["sample"], no_ts=True, lib=emul())
    for seed, (n, sub, indel) in enumerate([(900, 0.05, 0.03), (1100, 0.2, 0.1), (800, 0.01, 0.0)]):
        r, q = workloads.long_pair(100 + seed, n, sub_rate=sub, indel_rate=indel)
        res, stats = api.align_long(aligner, r, q, devices=[0] * world, interval=interval, group=group)
        _check(flat, r, q, res)
        assert len(stats) == world and sum(s["tiles"] + s["speculated_used"] for s in stats) >= 1
        if world > 1:
            assert all(s["boundary_bytes_out"] == 8 * (len(r) + 1) for s in stats[:-1]) and stats[-1]["boundary_bytes_out"] == 0


def test_band_protocol_emulated(configs):
    # the per-rank protocol a multi-process launch uses (handles, forward, walk hand-over)
    flat = oracle.FlatConfig(parse_config_any(configs["sample"]))
    aligner = tsa.Aligner(costs=configs["sample"], no_ts=True, lib=emul())
    r, q = workloads.long_pair(7, 1000, sub_rate=0.1, indel_rate=0.05)
    bands = [api.LongBand(aligner, r, q, k, 3, interval=128, group=1) for k in range(3)]
    res, kind = api.run_long_bands(bands)
    assert kind == "FoundTarget"
    _check(flat, r, q, res)
    assert bands[0].owner(0) == 0 and bands[0].owner(len(q)) == 2


def test_band_shapes_and_limits_emulated(configs):
    flat = oracle.FlatConfig(parse_config_any(configs["sample"]))
    aligner = tsa.Aligner(costs=configs["sample"], no_ts=True, lib=emul())
    # unequal lengths, a gap-only alignment, an embedded range, a pair narrower than the number of devices
    cases = [("ACGT" * 100, "ACGT" * 170), ("A" * 300, "C" * 10), ("ACGTTGCA" * 80, "ACGTTGCA" * 80), ("", "ACGT" * 70), ("ACGT" * 70, "")]
    for r, q in cases:
        res, _ = api.align_long(aligner, r, q, devices=[0, 0], interval=50, group=1)
        _check(flat, r, q, res)
    r, q = workloads.long_pair(3, 700, sub_rate=0.05, indel_rate=0.02)
    rng = (100, 650, 50, 600)
    res, _ = api.align_long(aligner, r, q, devices=[0, 0], interval=90, group=1, rng=rng)
    _check(flat, r, q, res, rng)
    assert res.range == rng
    # memory limit: the planner picks the spacing; a limit nothing fits gives ExceededMemoryLimit (generic_a_star/src/lib.rs:380-389)
    res, stats = api.align_long(aligner, r, q, devices=[0], memory_limit=200_000)
    _check(flat, r, q, res)
    assert stats[0]["resident_bytes"] <= 200_000 + 65536
    res, _ = api.align_long(aligner, r, q, devices=[0], memory_limit=1000)
    assert res.status == 0 and res.result_type == "ExceededMemoryLimit"
    # costs only
    costs = tsa.Aligner(costs=configs["sample"], no_ts=True, traceback=False, lib=emul())
    res, _ = api.align_long(costs, r, q, devices=[0, 0, 0])
    assert res.found and res.cost == oracle.dp_align(flat, r, q, no_ts=True).cost and res.ops is None


# ---- the same through the CUDA library on a B200 ----------------------------------------------------------------------------------
@pytest.mark.gpu
def test_bands_gpu(configs):
    from template_switch_aligner_b200 import _lib
    lib = _lib.default()
    assert b"sm_100a" in lib.tsa_version()
    flat = oracle.FlatConfig(parse_config_any(configs["sample"]))
    aligner = tsa.Aligner(costs=configs["sample"], no_ts=True, device=0)
    n_dev = lib.tsa_device_count()
    for seed, (n, sub, indel, interval, group) in enumerate([(3000, 0.05, 0.03, 256, 1), (5000, 0.02, 0.01, 512, 2), (2500, 0.3, 0.2, 100, 1), (6000, 0.012, 0.003, 0, 0)]):
        r, q = workloads.long_pair(200 + seed, n, sub_rate=sub, indel_rate=indel)
        for world in sorted({1, 2, min(4, max(1, n_dev))}):
            devices = [k % n_dev for k in range(world)]          # one box with fewer GPUs: several bands share a device
            res, stats = api.align_long(aligner, r, q, devices=devices, interval=interval, group=group)
            _check(flat, r, q, res)


@pytest.mark.gpu
def test_long_pair_matches_resident_codes_gpu(configs):
    # 40 kb: checkpointed traceback under a small memory limit against the ordinary path with the whole code matrix resident
    flat = oracle.FlatConfig(parse_config_any(configs["sample"]))
    aligner = tsa.Aligner(costs=configs["sample"], no_ts=True, device=0)
    r, q = workloads.long_pair(9, 40000, sub_rate=0.012, indel_rate=0.003)
    want = aligner.align_batch([(r, q)])[0]
    res, stats = api.align_long(aligner, r, q, devices=[0], memory_limit=64_000_000)
    assert res.found and want.found and res.cost == want.cost
    assert stats[0]["resident_bytes"] <= 64_000_000 + (1 << 20)
    cost, er, eq, ok = oracle.rescore(flat, r, q, [oracle.Op(*o) for o in res.ops], 0, 0)
    assert ok and cost == res.cost and (er, eq) == (len(r), len(q))
    bands = [api.LongBand(aligner, r, q, k, 2, interval=2048, group=4) for k in range(2)]   # the per-rank protocol (CUDA IPC handles need two processes: bench_c5.py)
    del bands


def test_batch_memory_limit_emulated(configs):
    # --memory-limit on the batch entry point: a --no-ts pair whose code matrix does not fit is aligned by the checkpointed path
    # (same cost, alignment rescored); with template switches the pair is reported as ExceededMemoryLimit (generic_a_star/src/lib.rs:380-389)
    flat = oracle.FlatConfig(parse_config_any(configs["sample"]))
    r, q = workloads.long_pair(21, 1300, sub_rate=0.05, indel_rate=0.02)
    small = ("ACGTTGCA" * 10, "ACGTTGCA" * 10)
    nots = tsa.Aligner(costs=configs["sample"], no_ts=True, lib=emul())
    got = nots.align_batch([small, (r, q), small], memory_limit=1 << 20)
    assert [g.result_type for g in got] == ["FoundTarget"] * 3
    _check(flat, r, q, got[1])
    _check(flat, small[0], small[1], got[0])
    ts = tsa.Aligner(costs=configs["sample"], lib=emul())
    got = ts.align_batch([small, (r[:400], q[:400]), small], memory_limit=1 << 20)
    assert [g.result_type for g in got] == ["FoundTarget", "ExceededMemoryLimit", "FoundTarget"] and got[1].status == 0 and got[1].ops is None


def test_batch_checkpoints_emulated(configs):
    # --no-ts batches with alignments through checkpoint rows + recomputed tiles (k_band_batch_*, what chunks of long pairs use on
    # the GPU), forced here on short pairs: several tiles per pair, edge shapes, a range; and the code-matrix path on the same pairs
    flat = oracle.FlatConfig(parse_config_any(configs["sample"]))
    pairs = [workloads.long_pair(300 + k, n, sub_rate=s, indel_rate=d) for k, (n, s, d) in enumerate([(700, 0.05, 0.03), (1300, 0.1, 0.05), (300, 0.02, 0.01), (900, 0.3, 0.2)])]
    pairs += [("ACGT" * 100, "ACGT" * 170), ("A" * 300, "C" * 10), ("", "ACGT" * 70), ("ACGT" * 70, ""), ("", "")]
    ck = tsa.Aligner(costs=configs["sample"], no_ts=True, lib=emul(), dev_flags=8).align_batch(pairs)
    codes = tsa.Aligner(costs=configs["sample"], no_ts=True, lib=emul(), dev_flags=16).align_batch(pairs)
    for (r, q), g, c in zip(pairs, ck, codes):
        _check(flat, r, q, g)
        assert c.cost == g.cost
    r, q = pairs[1]
    rng = (100, 1200, 50, 1150)
    res = tsa.Aligner(costs=configs["sample"], no_ts=True, lib=emul(), dev_flags=8).align_batch([(r, q, rng)])[0]
    _check(flat, r, q, res, rng)
