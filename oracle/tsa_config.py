"""config.tsa parser -- TEST INFRASTRUCTURE ONLY (independent of the product's C++ parser).

Restates lib_tsalign/src/config/io.rs:33-111 (section order), :181-221 (inf literals),
costs/cost_function/io.rs:81-120 (two-row step functions, first index must be the
type minimum, strictly increasing), costs/gap_affine/io.rs:156-359 (tables keyed by
letter in any column/row order) and config.rs:72-85 + cost_function.rs:170-176
(V-shape validation).
"""
import re
from dataclasses import dataclass, field

INF = (1 << 64) - 1
I64_MIN = -(1 << 63)
I64_MAX = (1 << 63) - 1

TABLE_NAMES = [
    "Primary Edit Costs",
    "Secondary Forward Edit Costs",
    "Secondary Reverse Edit Costs",
    "Left Flank Edit Costs",
    "Right Flank Edit Costs",
]
FN_NAMES = ["RQQROffset", "RRQQOffset", "Length", "LengthDifference", "ForwardAntiPrimaryGap", "ReverseAntiPrimaryGap"]
BASE_NAMES = ["rrf", "rqf", "qrf", "qqf", "rrr", "rqr", "qrr", "qqr"]


class ConfigError(ValueError):
    pass


@dataclass
class Table:
    name: str
    sub: list  # [A][A] by alphabet index
    open: list
    ext: list


@dataclass
class Config:
    alphabet: str
    chars: str
    left_flank_length: int = 0
    right_flank_length: int = 0
    base: list = field(default_factory=list)      # 8, BASE_NAMES order
    fns: list = field(default_factory=list)       # 6 x [(x, cost)]
    tables: list = field(default_factory=list)    # 5 x Table

    @property
    def min_length(self):
        for x, c in self.fns[2]:
            if c != INF:
                return x
        return None

    def evaluate(self, k, x):
        pts = self.fns[k]
        last = None
        for px, pc in pts:
            if px <= x:
                last = pc
            else:
                break
        if last is None:
            raise ConfigError("input before the first point")
        return last


def _parse_value(tok, signed):
    m = re.fullmatch(r"([+-]?)(inf|\d+)", tok)
    if not m:
        raise ConfigError(f"bad value {tok!r}")
    neg = m.group(1) == "-"
    if m.group(2) == "inf":
        if signed:
            return I64_MIN if neg else I64_MAX
        return 0 if neg else INF
    v = int(m.group(2))
    if neg:
        if not signed and v != 0:
            raise ConfigError(f"negative value {tok!r} for an unsigned field")
        v = -v
    return v


def is_v_shaped(pts):
    for (x0, c0), (x1, c1) in zip(pts, pts[1:]):
        ok = (x0 < 0 and x1 > 0) or (x0 < 0 and c0 >= c1) or (x0 >= 0 and c0 <= c1)
        if not ok:
            return False
    return True


def parse(text: str, alphabet: str = "dna-n") -> Config:
    from . import alphabets
    chars = alphabets.chars(alphabet)
    A = len(chars)
    lines = [ln.strip() for ln in text.splitlines()]
    lines = [ln for ln in lines if ln]
    pos = 0

    def take():
        nonlocal pos
        if pos >= len(lines):
            raise ConfigError("unexpected end of config")
        pos += 1
        return lines[pos - 1]

    def title(name):
        ln = take()
        m = re.fullmatch(r"#\s*(.*?)\s*", ln)
        if not m or m.group(1) != name:
            raise ConfigError(f"expected section '# {name}', got {ln!r}")

    def kv(name, signed):
        ln = take()
        m = re.fullmatch(r"([A-Za-z0-9_]+)\s*=\s*(\S+)", ln)
        if not m or m.group(1) != name:
            raise ConfigError(f"expected '{name} = <value>', got {ln!r}")
        return _parse_value(m.group(2), signed)

    cfg = Config(alphabet=alphabet, chars=chars)
    title("Limits")
    cfg.left_flank_length = kv("left_flank_length", True)
    cfg.right_flank_length = kv("right_flank_length", True)
    title("Base Cost")
    cfg.base = [kv(f"{n}_cost", False) for n in BASE_NAMES]
    title("Jump Costs")
    for k, name in enumerate(FN_NAMES):
        if take() != name:
            raise ConfigError(f"expected cost function {name}")
        signed = name != "Length"
        xs = [_parse_value(t, signed) for t in take().split()]
        cs = [_parse_value(t, False) for t in take().split()]
        first = I64_MIN if signed else 0
        if len(xs) != len(cs) or not xs or xs[0] != first or any(a >= b for a, b in zip(xs, xs[1:])):
            raise ConfigError(f"malformed cost function {name}")
        cfg.fns.append(list(zip(xs, cs)))
    for name in TABLE_NAMES:
        title(name)
        if take() != "SubstitutionCostTable":
            raise ConfigError("expected SubstitutionCostTable")
        hdr = take()
        if not hdr.startswith("|"):
            raise ConfigError("expected '| <characters>' header row")
        cols = hdr[1:].split()
        if sorted(cols) != sorted(chars):
            raise ConfigError(f"table columns {cols} do not match alphabet {chars}")
        if not re.fullmatch(r"-+\+-+", take()):
            raise ConfigError("expected separator line")
        sub = [[None] * A for _ in range(A)]
        seen = set()
        for _ in range(A):
            m = re.fullmatch(r"(\S)\s*\|\s*(.*)", take())
            if not m or m.group(1) not in chars:
                raise ConfigError("bad substitution row")
            r = chars.index(m.group(1))
            seen.add(r)
            vals = [_parse_value(t, False) for t in m.group(2).split()]
            if len(vals) != A:
                raise ConfigError("bad substitution row width")
            for cname, v in zip(cols, vals):
                sub[r][chars.index(cname)] = v
        if len(seen) != A:
            raise ConfigError("duplicate substitution rows")
        vecs = []
        for vname in ("GapOpenCostVector", "GapExtendCostVector"):
            if take() != vname:
                raise ConfigError(f"expected {vname}")
            idx = take().split()
            if sorted(idx) != sorted(chars):
                raise ConfigError("bad cost vector index row")
            vals = [_parse_value(t, False) for t in take().split()]
            if len(vals) != A:
                raise ConfigError("bad cost vector width")
            vec = [None] * A
            for cname, v in zip(idx, vals):
                vec[chars.index(cname)] = v
            vecs.append(vec)
        cfg.tables.append(Table(name, sub, vecs[0], vecs[1]))
    # config.rs:72-85
    if not is_v_shaped(cfg.fns[0]):
        raise ConfigError("RQQROffsetCostsNotVShaped")
    if not is_v_shaped(cfg.fns[1]):
        raise ConfigError("RRQQOffsetCostsNotVShaped")
    if not is_v_shaped(cfg.fns[3]):
        raise ConfigError("LengthDifferenceCostsNotVShaped")
    return cfg


def base_agnostic(alphabet, name, match, sub, gap_open, gap_ext):
    from . import alphabets
    A = len(alphabets.chars(alphabet))
    return Table(name, [[match if a == b else sub for b in range(A)] for a in range(A)], [gap_open] * A, [gap_ext] * A)


def rust_default(alphabet: str = "dna-n") -> Config:
    """TemplateSwitchConfig::default(), lib_tsalign/src/config.rs:219-303."""
    from . import alphabets
    chars = alphabets.chars(alphabet)
    cfg = Config(alphabet=alphabet, chars=chars)
    cfg.base = [4, 4, 4, 4, 3, 2, 2, 3]
    cfg.tables = [base_agnostic(alphabet, n, 0, 2, 3, 1) for n in TABLE_NAMES]
    cfg.fns = [
        [(I64_MIN, INF), (-100, 0), (101, INF)],
        [(I64_MIN, INF), (-100, 0), (1, INF)],
        [(0, INF), (5, 0)],
        [(I64_MIN, INF), (-100, 0), (101, INF)],
        [(I64_MIN, INF), (-100, 0), (101, INF)],
        [(I64_MIN, INF), (-100, 0), (101, INF)],
    ]
    return cfg
