"""Parity checks shared by the emulator tests (CPU) and the GPU tests: product path vs. the CPU oracle."""
import random

from oracle import oracle
from helpers import clean_record, config_to_text, parse_config_any
import randcfg
import template_switch_aligner_b200 as tsa
from template_switch_aligner_b200 import api


def end_points(ops, ri, qi):
    """Reference / query coordinates after the run-length encoded operations (alignment/iter.rs:62-90 semantics)."""
    primary, pi = 0, 0
    for op in ops:
        t, count = op.type, op.count
        if t < 8:
            kind = t & 3
            ri += count if kind in (1, 2, 3) else 0
            qi += count if kind in (0, 2, 3) else 0
        elif t in (8, 10, 11):
            pi += count          # secondary insertion / substitution / match consume the primary sequence
        elif t == 12:
            primary, pi = op.primary, (ri if op.primary == 0 else qi)
        elif t == 13:
            if primary == 0:
                ri, qi = pi, qi + op.value
            else:
                qi, ri = pi, ri + op.value
    return ri, qi


def merge_flank_runs(ops):
    """The reference's run-length rule (alignment_type.rs:101-121, a_star_aligner.rs:100-122): flank and non-flank variants of a
    primary operation repeat each other, a merged run carries the label of its last operation."""
    out = []
    for o in ops:
        if out and o.type < 8 and out[-1].type < 8 and (o.type & 3) == (out[-1].type & 3):
            out[-1] = oracle.Op(out[-1].count + o.count, o.type)
        else:
            out.append(oracle.Op(o.count, o.type, o.primary, o.secondary, o.direction, o.value))
    return out


def check_alignment(flat, p, g, label="", merged=None):
    """The returned alignment must rescore to the returned cost under the reference cost function
    (compute_cost restatement, template_switch_specifics.rs:591-835) and span exactly the requested range.
    With flank lengths > 0 the reference's run-length encoding merges flank and non-flank operations
    (alignment_type.rs:101-121) and its compute_cost is todo!() there: the alignment is requested with
    TSA_FLAG_KEEP_FLANK_RUNS (flank runs kept apart) and scored as the search charges flank moves (context.rs:225-353);
    `merged` (the same alignment requested without the flag) must be the reference's merge of it."""
    r, q = p[0], p[1]
    rng = p[2] if len(p) > 2 and p[2] is not None else (0, len(r), 0, len(q))
    assert g.ops is not None, (label, p)
    ops = [oracle.Op(*o) for o in g.ops]
    flanks = flat.cfg.left_flank_length != 0 or flat.cfg.right_flank_length != 0
    cost, er, eq, ok = oracle.rescore(flat, r, q, ops, rng[0], rng[2], as_searched=True)
    assert ok and (er, eq) == (rng[1], rng[3]), (label, p, tsa.cigar_of(g.ops), cost, g.cost, er, eq)
    assert cost == g.cost, (label, p, tsa.cigar_of(g.ops), cost, g.cost)
    if flanks and merged is not None:
        mops = [oracle.Op(*o) for o in merged.ops]
        assert merged.cost == g.cost and end_points(mops, rng[0], rng[2]) == (rng[1], rng[3]), (label, p)
        want = merge_flank_runs(ops)
        assert [(o.count, o.type, o.value) for o in mops] == [(o.count, o.type, o.value) for o in want], (label, p, tsa.cigar_of(merged.ops))
    assert sum(1 for o in ops if o.type == oracle.OP_TS_EXIT) == g.template_switches


def check_batch(aligner, flat, pairs, no_ts=False, label="", expected=None):
    """pairs: [(r, q) | (r, q, range)].  The product's costs must equal the scalar DP oracle's, bit for bit, and its
    alignments must rescore to them.  expected: optional precomputed oracle costs (None = no target)."""
    flanks = flat.cfg.left_flank_length != 0 or flat.cfg.right_flank_length != 0
    merged = None
    if flanks and aligner.traceback:
        merged = aligner.align_batch(pairs)          # the reference's run-length encoding (flank runs merged)
        saved = aligner.flags
        aligner.flags |= api.FLAG_KEEP_FLANK_RUNS
        got = aligner.align_batch(pairs)
        aligner.flags = saved
    else:
        got = aligner.align_batch(pairs)
    n_ts = 0
    for idx, (p, g) in enumerate(zip(pairs, got)):
        rng = p[2] if len(p) > 2 else None
        assert g.status == 0, (label, p, g.message)
        if expected is not None:
            want_found, want_cost = expected[idx] is not None, expected[idx]
        else:
            want = oracle.dp_align(flat, p[0], p[1], rng, no_ts=no_ts)
            want_found, want_cost = want.found, want.cost
        assert g.found == want_found, (label, p, g.result_type)
        if want_found:
            assert g.cost == want_cost, (label, p, g.cost, want_cost)
            n_ts += g.template_switches > 0
            if aligner.traceback:
                check_alignment(flat, p, g, label, merged[idx] if merged is not None else None)
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


def random_model_batches(lib, seeds, max_len, pairs_per_model=6, device=0, first_threshold=0, flanks=False, dev_flags=0, min_len=0):
    """Random cost models (pieces, quirks, infinite entries) x random pairs and ranges."""
    total_ts = 0
    for seed in seeds:
        rng = random.Random(seed)
        cfg = randcfg.random_config(rng, flanks=flanks)
        flat = oracle.FlatConfig(cfg)
        ps = []
        for _ in range(pairs_per_model):
            r, q = randcfg.random_pair(rng, max_len=max_len)
            while max(len(r), len(q)) < min_len:
                r, q = randcfg.random_pair(rng, max_len=max_len)
            ps.append((r, q, randcfg.random_range(rng, r, q)))
        for no_ts in (False, True):
            aligner = tsa.Aligner(costs=config_to_text(cfg), no_ts=no_ts, lib=lib, device=device, first_threshold=first_threshold or (1 + seed % 7), dev_flags=dev_flags)
            total_ts += check_batch(aligner, flat, ps, no_ts=no_ts, label=f"seed {seed}")
    return total_ts
