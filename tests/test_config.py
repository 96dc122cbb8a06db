"""Host cost-model layer of the product (csrc/tsa_config.cpp) through the C ABI -- no GPU needed: parsing,
error kinds (lib_tsalign/src/error.rs:5-49), Display -> parse round trip (config/io.rs:277-293)."""
import ctypes
import os

import pytest

from oracle import tsa_config
from template_switch_aligner_b200 import _lib, workloads
import template_switch_aligner_b200 as tsa
from helpers import parse_config_any


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.default()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(os.path.dirname(_lib._HERE), "include", "tsalign_b200.h")).read()
    import re
    declared = set(re.findall(r"\b(tsa_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.tsa_version()


def test_parse_matches_oracle_parser(lib, configs):
    for name in ("sample", "bench", "experiments", "small", "range", "no_intra_forward_jump"):
        ocfg = parse_config_any(configs[name])
        cfg = tsa.Config(configs[name], ocfg.alphabet, lib=lib)
        # the product's Display output parses back to the same model in the (independent) oracle parser
        again = tsa_config.parse(cfg.text(), ocfg.alphabet)
        assert again.base == ocfg.base and again.fns == ocfg.fns
        assert (again.left_flank_length, again.right_flank_length) == (ocfg.left_flank_length, ocfg.right_flank_length)
        for a, b in zip(again.tables, ocfg.tables):
            assert (a.sub, a.open, a.ext) == (b.sub, b.open, b.ext)


def test_default_config_round_trip(lib):
    # config/io.rs:277-293: Default -> Display -> parse is the identity
    cfg = tsa.Config(None, "dna-n", lib=lib)
    ref = tsa_config.rust_default("dna-n")
    again = tsa_config.parse(cfg.text(), "dna-n")
    assert again.base == ref.base and again.fns == ref.fns
    for a, b in zip(again.tables, ref.tables):
        assert (a.sub, a.open, a.ext) == (b.sub, b.open, b.ext)
    tsa.Config(cfg.text(), "dna-n", lib=lib)


def test_shipped_sample_config_equals_reference_sample(configs):
    a = tsa_config.parse(workloads.sample_config_text(), "dna-n")
    b = tsa_config.parse(configs["sample"], "dna-n")
    assert a.base == b.base and a.fns == b.fns
    for x, y in zip(a.tables, b.tables):
        assert (x.sub, x.open, x.ext) == (y.sub, y.open, y.ext)


def test_error_kinds(lib, configs):
    with pytest.raises(tsa.TsaError) as e:
        tsa.Config(configs["indel"], "dna-n", lib=lib)  # stale pre-v2 file
    assert e.value.status == 2
    with pytest.raises(tsa.TsaError) as e:
        tsa.Config(configs["bench"], "dna-n", lib=lib)  # 4-letter tables, 5-letter alphabet
    assert e.value.status == 2
    sample = configs["sample"]
    bad = sample.replace("RQQROffset\n -inf -100 101\n  inf    0 inf", "RQQROffset\n -inf -100 0 101\n  inf    0 3 1")
    assert bad != sample
    with pytest.raises(tsa.TsaError) as e:
        tsa.Config(bad, "dna-n", lib=lib)
    assert e.value.status == 3  # RQQROffsetCostsNotVShaped
    bad = sample.replace("LengthDifference\n -inf -100 101\n  inf    0 inf", "LengthDifference\n -inf -100 -5 101\n  inf    0 2 inf")
    with pytest.raises(tsa.TsaError) as e:
        tsa.Config(bad, "dna-n", lib=lib)
    assert e.value.status == 5
    # cost functions must start at the type minimum (cost_function/io.rs:103-110)
    bad = sample.replace("Length\n   0 5 6 7 8 100", "Length\n   1 5 6 7 8 100")
    with pytest.raises(tsa.TsaError) as e:
        tsa.Config(bad, "dna-n", lib=lib)
    assert e.value.status == 2
    # '+' signs and inf literals are accepted (config/io.rs:181-221)
    ok = sample.replace("rrf_cost = 3", "rrf_cost = +3").replace("rqf_cost = 2", "rqf_cost = inf")
    tsa.Config(ok, "dna-n", lib=lib)


def test_no_device_fails_loudly(lib):
    # This container has no GPU: the compute entry must refuse, not fall back to anything.
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    aligner = tsa.Aligner(costs=workloads.sample_config_text(), lib=lib)
    with pytest.raises(tsa.TsaError) as e:
        aligner.align_batch([("ACGT", "ACGT")])
    assert e.value.status == 1
    assert lib.tsa_device_count() == 0
