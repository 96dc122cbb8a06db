"""Seeded synthetic workloads of BASELINE.json / SURVEY.md section 8(d) (generator = splitmix64,
seed = 0x7453414C49474E xor pair index).  Shared by bench.py and the parity tests so that the GPU path and the
CPU baseline see the same pairs."""
from __future__ import annotations

import os

_MASK = (1 << 64) - 1
_SEED = 0x7453414C49474E
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & _MASK

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        return z ^ (z >> 31)

    def below(self, n: int) -> int:
        return self.next() % n

    def uniform(self, lo: int, hi: int) -> int:  # inclusive
        return lo + self.below(hi - lo + 1)

    def chance(self, p: float) -> bool:
        return self.next() < int(p * (1 << 64))


def revcomp(s: str) -> str:
    return "".join(_COMP[c] for c in reversed(s))


def _mutate(rng: SplitMix64, ref: str, sub_rate: float, indel_rate: float, max_indel: int = 3) -> str:
    out = []
    i = 0
    while i < len(ref):
        if rng.chance(indel_rate):
            k = rng.uniform(1, max_indel)
            if rng.chance(0.5):
                i += k  # deletion
                continue
            out.extend("ACGT"[rng.below(4)] for _ in range(k))  # insertion
        c = ref[i]
        if rng.chance(sub_rate):
            c = "ACGT"[(("ACGT".index(c)) + 1 + rng.below(3)) % 4]
        out.append(c)
        i += 1
    return "".join(out)


def read_pair(index: int, length: int = 150, sub_rate: float = 0.01, indel_rate: float = 0.002, n_tsm: int = 1):
    """C2 pair: reference = `length` uniform ACGT; query = copy with substitutions, short indels and `n_tsm`
    planted reverse template-switch mutations: query[p, p+l) := revcomp(reference[p+o-l, p+o)),
    l ~ U[8, 30], o ~ U[-40, 40], p ~ U[20, length-50]."""
    rng = SplitMix64(_SEED ^ index)
    ref = "".join("ACGT"[rng.below(4)] for _ in range(length))
    qry = list(_mutate(rng, ref, sub_rate, indel_rate))
    for _ in range(n_tsm):
        for _try in range(16):
            l = rng.uniform(8, 30)
            o = rng.uniform(-40, 40)
            p = rng.uniform(20, max(21, length - 50))
            a, b = p + o - l, p + o
            if a >= 0 and b <= len(ref) and p + l <= len(qry):
                qry[p:p + l] = list(revcomp(ref[a:b]))
                break
    return ref, "".join(qry)


def read_pairs(count: int, start: int = 0, length: int = 150, **kw):
    return [read_pair(start + i, length, **kw) for i in range(count)]


def long_pair(index: int, length: int, sub_rate: float = 0.01, indel_rate: float = 0.005, n_tsm: int = 0, spacing: int = 120):
    """C3 / C4 shaped pair (1 kb with 5 planted TSMs, or 10 kb without)."""
    rng = SplitMix64(_SEED ^ (index + (length << 20)))
    ref = "".join("ACGT"[rng.below(4)] for _ in range(length))
    qry = list(_mutate(rng, ref, sub_rate, indel_rate))
    pos = 60
    for _ in range(n_tsm):
        l = rng.uniform(8, 30)
        o = rng.uniform(-40, 40)
        p = pos + rng.below(40)
        a, b = p + o - l, p + o
        if a >= 0 and b <= len(ref) and p + l <= len(qry):
            qry[p:p + l] = list(revcomp(ref[a:b]))
        pos += spacing + 40
    return ref, "".join(qry)


def _table(name: str, chars: str, match: int, sub: int, n_cost: int, gap_open: int, gap_ext: int) -> str:
    rows = []
    for a in chars:
        vals = []
        for b in chars:
            vals.append(n_cost if "N" in (a, b) else (match if a == b else sub))
        rows.append(f"{a} | " + " ".join(f"{v:2d}" for v in vals))
    head = "  | " + " ".join(f"{c:>2}" for c in chars)
    return "\n".join([f"# {name}", "", "SubstitutionCostTable", head, "--+" + "-" * (3 * len(chars)), *rows, "",
                      "GapOpenCostVector", " " + " ".join(chars), " " + " ".join(str(gap_open) for _ in chars), "",
                      "GapExtendCostVector", " " + " ".join(chars), " " + " ".join(str(gap_ext) for _ in chars), ""])


def sample_config_text() -> str:
    """The cost model `tsalign align` uses by default (the values of the reference's sample_tsa_config/config.tsa,
    SURVEY.md appendix B; tests/test_config.py checks them against the golden copy), alphabet dna-n."""
    chars = "ACGTN"
    head = "\n".join([
        "# Limits", "", "left_flank_length = 0", "right_flank_length = 0", "",
        "# Base Cost", "", "rrf_cost = 3", "rqf_cost = 2", "qrf_cost = 2", "qqf_cost = 3",
        "rrr_cost = 3", "rqr_cost = 2", "qrr_cost = 2", "qqr_cost = 3", "",
        "# Jump Costs", "",
        "RQQROffset", " -inf -100 101", "  inf    0 inf", "",
        "RRQQOffset", " -inf -100 101", "  inf    0 inf", "",
        "Length", "   0 5 6 7 8 100", " inf 5 3 1 0 inf", "",
        "LengthDifference", " -inf -100 101", "  inf    0 inf", "",
        "ForwardAntiPrimaryGap", " -inf   1", "    0 inf", "",
        "ReverseAntiPrimaryGap", " -inf", "    0", "",
    ])
    tables = [
        _table("Primary Edit Costs", chars, 0, 2, 0, 3, 1),
        _table("Secondary Forward Edit Costs", chars, 0, 8, 4, 9, 2),
        _table("Secondary Reverse Edit Costs", chars, 0, 8, 4, 9, 2),
        _table("Left Flank Edit Costs", chars, 0, 3, 0, 4, 1),
        _table("Right Flank Edit Costs", chars, 0, 3, 0, 4, 1),
    ]
    return head + "\n" + "\n".join(tables)


def algorithmic_work(n: int, m: int, ts_count: int, n_kinds: int = 8, l_star: int = 95, n_off: int = 1, n_ld: int = 1, n_apg: int = 1, flank_planes: int = 1) -> float:
    """add-min lane operations of one pair,
    W = K (F 7 |R||Q| + N_kind L* n_ld n_apg |R||Q|) + N_kind L* (7 + n_off) |R||Q|, K = template switches + 1."""
    cells = float(n) * float(m)
    k = ts_count + 1
    return k * (flank_planes * 7 * cells + n_kinds * l_star * n_ld * n_apg * cells) + n_kinds * l_star * (7 + n_off) * cells
