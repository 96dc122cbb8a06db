"""Parity checks shared by the emulator tests (CPU) and the GPU tests: product path vs. the CPU oracle."""
import random

from oracle import oracle
from helpers import clean_record, config_to_text, parse_config_any
import randcfg
import template_switch_aligner_b200 as tsa


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


def check_alignment(flat, p, g, label=""):
    """The returned alignment must rescore to the returned cost under the reference cost function
    (compute_cost restatement, template_switch_specifics.rs:591-835) and span exactly the requested range.
    With flank lengths > 0 the reference's run-length encoding merges flank and non-flank operations
    (alignment_type.rs:101-121) and its compute_cost is todo!() there: only the end points are checked (the traceback
    kernel itself verifies that the edge costs of its path sum to the optimum, else status TSA_ERR_INTERNAL)."""
    r, q = p[0], p[1]
    rng = p[2] if len(p) > 2 and p[2] is not None else (0, len(r), 0, len(q))
    assert g.ops is not None, (label, p)
    ops = [oracle.Op(*o) for o in g.ops]
    if flat.cfg.left_flank_length == 0 and flat.cfg.right_flank_length == 0:
        cost, er, eq, ok = oracle.rescore(flat, r, q, ops, rng[0], rng[2], as_searched=True)
        assert ok and (er, eq) == (rng[1], rng[3]), (label, p, tsa.cigar_of(g.ops), cost, g.cost, er, eq)
        assert cost == g.cost, (label, p, tsa.cigar_of(g.ops), cost, g.cost)
    else:
        # merged runs carry the label of their last operation (a_star_aligner.rs:100-122), so the flank table of a unit
        # operation cannot be recovered from the run-length encoding: walk the coordinates only
        assert end_points(ops, rng[0], rng[2]) == (rng[1], rng[3]), (label, p, tsa.cigar_of(g.ops))
    assert sum(1 for o in ops if o.type == oracle.OP_TS_EXIT) == g.template_switches


def check_batch(aligner, flat, pairs, no_ts=False, label="", expected=None):
    """pairs: [(r, q) | (r, q, range)].  The product's costs must equal the scalar DP oracle's, bit for bit, and its
    alignments must rescore to them.  expected: optional precomputed oracle costs (None = no target)."""
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
                check_alignment(flat, p, g, label)
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


def random_model_batches(lib, seeds, max_len, pairs_per_model=6, device=0, first_threshold=0, flanks=False):
    """Random cost models (pieces, quirks, infinite entries) x random pairs and ranges."""
    total_ts = 0
    for seed in seeds:
        rng = random.Random(seed)
        cfg = randcfg.random_config(rng, flanks=flanks)
        flat = oracle.FlatConfig(cfg)
        ps = []
        for _ in range(pairs_per_model):
            r, q = randcfg.random_pair(rng, max_len=max_len)
            ps.append((r, q, randcfg.random_range(rng, r, q)))
        for no_ts in (False, True):
            aligner = tsa.Aligner(costs=config_to_text(cfg), no_ts=no_ts, lib=lib, device=device, first_threshold=first_threshold or (1 + seed % 7))
            total_ts += check_batch(aligner, flat, ps, no_ts=no_ts, label=f"seed {seed}")
    return total_ts
