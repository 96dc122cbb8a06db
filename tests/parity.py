"""Parity checks shared by the emulator tests (CPU) and the GPU tests: product path vs. the CPU oracle."""
import random

from oracle import oracle
from helpers import clean_record, config_to_text, parse_config_any
import randcfg
import template_switch_aligner_b200 as tsa


def check_batch(aligner, flat, pairs, no_ts=False, label=""):
    """pairs: [(r, q) | (r, q, range)].  The product's costs must equal the scalar DP oracle's, bit for bit."""
    got = aligner.align_batch(pairs)
    n_ts = 0
    for p, g in zip(pairs, got):
        rng = p[2] if len(p) > 2 else None
        want = oracle.dp_align(flat, p[0], p[1], rng, no_ts=no_ts)
        assert g.status == 0, (label, p, g.message)
        assert g.result_type == want.result_type, (label, p, g.result_type, want.result_type)
        if want.found:
            assert g.cost == want.cost, (label, p, g.cost, want.cost, want.cigar())
            n_ts += g.template_switches > 0
    return n_ts


def test_file_pairs(pairs, alphabet, max_len, min_len=0):
    out = []
    for name, p in sorted(pairs.items()):
        r = clean_record(p["records"][0][1].replace("|", ""))
        q = clean_record(p["records"][1][1].replace("|", ""))
        if not (min_len <= max(len(r), len(q)) <= max_len):
            continue
        try:
            oracle.alphabets.encode(alphabet, r), oracle.alphabets.encode(alphabet, q)
        except ValueError:
            continue
        out.append((name, r, q))
    return out


def random_model_batches(lib, seeds, max_len, pairs_per_model=6, device=0, first_threshold=0):
    """Random cost models (pieces, quirks, infinite entries) x random pairs and ranges."""
    total_ts = 0
    for seed in seeds:
        rng = random.Random(seed)
        cfg = randcfg.random_config(rng, flanks=False)
        flat = oracle.FlatConfig(cfg)
        ps = []
        for _ in range(pairs_per_model):
            r, q = randcfg.random_pair(rng, max_len=max_len)
            ps.append((r, q, randcfg.random_range(rng, r, q)))
        for no_ts in (False, True):
            aligner = tsa.Aligner(costs=config_to_text(cfg), no_ts=no_ts, lib=lib, device=device, first_threshold=first_threshold or (1 + seed % 7))
            total_ts += check_batch(aligner, flat, ps, no_ts=no_ts, label=f"seed {seed}")
    return total_ts
