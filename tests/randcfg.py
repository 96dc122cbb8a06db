"""Random cost models / sequences for the oracle-vs-oracle and GPU-vs-oracle parity tests."""
import random

from oracle import tsa_config, alphabets

INF = tsa_config.INF
I64_MIN = tsa_config.I64_MIN


def v_shaped(rng, lo, hi, max_cost=6, left_inf=True, right_inf=True):
    """Step function finite on [lo, hi] (lo <= 0 <= hi), non-increasing towards 0 and non-decreasing after."""
    pts = []
    neg = sorted(rng.sample(range(lo, 0), min(len(range(lo, 0)), rng.randint(0, 3)))) if lo < 0 else []
    pos = sorted(rng.sample(range(1, hi + 1), min(len(range(1, hi + 1)), rng.randint(0, 3)))) if hi > 0 else []
    c0 = rng.randint(0, 2)
    # costs left of zero: decreasing towards zero
    lc = sorted([rng.randint(c0, c0 + max_cost) for _ in range(len(neg) + 1)], reverse=True)
    rc = sorted([rng.randint(c0, c0 + max_cost) for _ in range(len(pos))])
    pts.append((I64_MIN, INF if left_inf else lc[0]))
    if left_inf:
        pts.append((lo, lc[0]))
    for x, c in zip(neg, lc[1:]):
        if x > lo:
            pts.append((x, c))
    # ensure the piece containing 0 has the minimum cost
    if pts[-1][0] != 0:
        pts.append((0, c0)) if rng.random() < 0.5 and pts[-1][1] != INF else None
    if pts[-1][1] == INF or pts[-1][1] > (rc[0] if rc else pts[-1][1]):
        if pts[-1][0] < 0 or pts[-1][0] == I64_MIN:
            pts.append((0, c0))
    for x, c in zip(pos, rc):
        if c >= pts[-1][1]:
            pts.append((x, c))
    if right_inf:
        pts.append((hi + 1, INF))
    # dedupe equal x
    out = []
    for x, c in pts:
        if out and out[-1][0] == x:
            out[-1] = (x, c)
        else:
            out.append((x, c))
    assert tsa_config.is_v_shaped(out), out
    return out


def any_shape(rng, lo, hi, max_cost=5):
    xs = sorted(set([lo] + rng.sample(range(lo, hi + 1), min(hi - lo + 1, rng.randint(0, 4)))))
    pts = [(I64_MIN, INF if rng.random() < 0.7 else rng.randint(0, max_cost))]
    for x in xs:
        pts.append((x, rng.randint(0, max_cost)))
    if rng.random() < 0.7:
        pts.append((hi + 1, INF))
    return pts


def length_fn(rng, max_len):
    mn = rng.randint(1, 4)
    pts = [(0, INF), (mn, rng.randint(0, 4))]
    x = mn
    for _ in range(rng.randint(0, 2)):
        x += rng.randint(1, 3)
        pts.append((x, rng.randint(0, 3)))
    if rng.random() < 0.7:
        pts.append((max(x + 1, rng.randint(mn + 1, max_len)), INF))
    return pts


def table(rng, alphabet, name, lo=0, hi=5, allow_inf=False):
    A = len(alphabets.chars(alphabet))
    if rng.random() < 0.5:
        t = tsa_config.base_agnostic(alphabet, name, rng.randint(0, 1) if rng.random() < 0.3 else 0, rng.randint(1, hi), rng.randint(1, hi + 2), rng.randint(0, 3))
    else:
        sub = [[(0 if x == y and rng.random() < 0.8 else rng.randint(lo, hi)) for y in range(A)] for x in range(A)]
        t = tsa_config.Table(name, sub, [rng.randint(1, hi + 2) for _ in range(A)], [rng.randint(0, 3) for _ in range(A)])
    if allow_inf and rng.random() < 0.2:
        t.sub[rng.randrange(A)][rng.randrange(A)] = INF
    if allow_inf and rng.random() < 0.1:
        t.open[rng.randrange(A)] = INF
    if allow_inf and rng.random() < 0.1:
        t.ext[rng.randrange(A)] = INF
    return t


def random_config(rng, alphabet="dna-n", flanks=False, span=8, allow_inf=True):
    cfg = tsa_config.Config(alphabet=alphabet, chars=alphabets.chars(alphabet))
    if flanks:
        cfg.left_flank_length = rng.randint(0, 2)
        cfg.right_flank_length = rng.randint(0, 2)
    cfg.base = [(INF if rng.random() < 0.25 else rng.randint(1, 4)) for _ in range(8)]
    cfg.fns = [
        v_shaped(rng, -rng.randint(0, span), rng.randint(0, span)),
        v_shaped(rng, -rng.randint(0, span), rng.randint(0, span)),
        length_fn(rng, span + 4),
        v_shaped(rng, -rng.randint(0, span), rng.randint(0, span)),
        any_shape(rng, -rng.randint(0, span), rng.randint(0, span)),
        any_shape(rng, -rng.randint(0, span), rng.randint(0, span)),
    ]
    cfg.tables = [table(rng, alphabet, n, allow_inf=allow_inf) for n in tsa_config.TABLE_NAMES]
    return cfg


def random_pair(rng, alphabet="dna-n", max_len=12, n_weight=0.05):
    chars = alphabets.chars(alphabet)
    core = chars[:4]

    def seq(L):
        return "".join((rng.choice(chars) if rng.random() < n_weight else rng.choice(core[: rng.choice([2, 4])])) for _ in range(L))

    r = seq(rng.randint(0, max_len))
    if rng.random() < 0.6 and len(r) > 3:
        # derive the query from the reference: substitutions, indels and a reverse-complement patch
        q = list(r)
        for _ in range(rng.randint(0, 2)):
            if q:
                q[rng.randrange(len(q))] = rng.choice(core)
        if rng.random() < 0.5 and len(q) > 4:
            a = rng.randrange(len(q) - 3)
            b = min(len(q), a + rng.randint(2, 6))
            q[a:b] = list(alphabets.reverse_complement(alphabet, "".join(q[a:b])))
        if rng.random() < 0.4 and q:
            a = rng.randrange(len(q))
            del q[a:a + rng.randint(1, 2)]
        q = "".join(q)
    else:
        q = seq(rng.randint(0, max_len))
    return r, q


def random_range(rng, r, q):
    if rng.random() < 0.5:
        return (0, len(r), 0, len(q))
    ro = rng.randint(0, len(r)); rl = rng.randint(ro, len(r))
    qo = rng.randint(0, len(q)); ql = rng.randint(qo, len(q))
    return (ro, rl, qo, ql)
