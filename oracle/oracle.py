"""ctypes front-end of the CPU oracle (libtsa_oracle.so) -- TEST INFRASTRUCTURE ONLY.

See tsa_oracle.h for the scope rule: tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke() only.
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass

from . import alphabets, tsa_config

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OP_NAMES = [
    "PrimaryInsertion", "PrimaryDeletion", "PrimarySubstitution", "PrimaryMatch",
    "PrimaryFlankInsertion", "PrimaryFlankDeletion", "PrimaryFlankSubstitution", "PrimaryFlankMatch",
    "SecondaryInsertion", "SecondaryDeletion", "SecondarySubstitution", "SecondaryMatch",
    "TemplateSwitchEntrance", "TemplateSwitchExit",
]
OP_INDEX = {n: i for i, n in enumerate(OP_NAMES)}
OP_TS_ENTRANCE, OP_TS_EXIT = 12, 13
RESULT_NAMES = ["FoundTarget", "ExceededCostLimit", "ExceededMemoryLimit", "NoTarget"]
U64_MAX = (1 << 64) - 1


class _CConfig(C.Structure):
    _fields_ = [
        ("alphabet_size", C.c_int32),
        ("sub", C.POINTER(C.c_uint64)),
        ("open", C.POINTER(C.c_uint64)),
        ("ext", C.POINTER(C.c_uint64)),
        ("base", C.c_uint64 * 8),
        ("fn_len", C.c_int32 * 6),
        ("fn_x", C.POINTER(C.c_int64) * 6),
        ("fn_c", C.POINTER(C.c_uint64) * 6),
        ("left_flank_length", C.c_int64),
        ("right_flank_length", C.c_int64),
        ("complement", C.POINTER(C.c_uint8)),
    ]


class _COp(C.Structure):
    _fields_ = [("count", C.c_int64), ("type", C.c_int32), ("primary", C.c_int32), ("secondary", C.c_int32),
                ("direction", C.c_int32), ("value", C.c_int64)]


class _CResult(C.Structure):
    _fields_ = [("result_type", C.c_int32), ("cost", C.c_uint64), ("opened_nodes", C.c_uint64),
                ("closed_nodes", C.c_uint64), ("suboptimal_opened_nodes", C.c_uint64),
                ("ts_total_length", C.c_uint64), ("ops", C.POINTER(_COp)), ("n_ops", C.c_int64)]


class _COptions(C.Structure):
    _fields_ = [("no_ts", C.c_int32), ("total_length_maximise", C.c_int32), ("min_length_lookahead", C.c_int32),
                ("force_label_correcting", C.c_int32), ("cost_limit", C.c_uint64), ("memory_limit", C.c_uint64)]


def build(force=False):
    """Compile libtsa_oracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "libtsa_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("astar_oracle.cpp", "dp_oracle.cpp", "tsa_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        for name in ("tsao_astar_align", "tsao_dp_align"):
            f = getattr(_LIB, name)
            f.restype = C.c_int
            f.argtypes = [C.POINTER(_CConfig), C.c_char_p, C.c_int64, C.c_char_p, C.c_int64,
                          C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(_COptions), C.POINTER(_CResult)]
        _LIB.tsao_rescore.restype = C.c_uint64
        _LIB.tsao_rescore.argtypes = [C.POINTER(_CConfig), C.c_char_p, C.c_int64, C.c_char_p, C.c_int64,
                                      C.c_int64, C.c_int64, C.POINTER(_COp), C.c_int64,
                                      C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        _LIB.tsao_rescore_mode.restype = C.c_uint64
        _LIB.tsao_rescore_mode.argtypes = _LIB.tsao_rescore.argtypes + [C.c_int32]
        _LIB.tsao_result_free.argtypes = [C.POINTER(_CResult)]
    return _LIB


@dataclass
class Op:
    count: int
    type: int
    primary: int = 0
    secondary: int = 0
    direction: int = 0
    value: int = 0

    @property
    def name(self):
        return OP_NAMES[self.type]


@dataclass
class Result:
    result_type: str
    cost: int
    ops: list
    opened_nodes: int = 0
    closed_nodes: int = 0
    suboptimal_opened_nodes: int = 0
    ts_total_length: int = 0

    @property
    def found(self):
        return self.result_type == "FoundTarget"

    def cigar(self):
        return cigar(self.ops)


def cigar(ops):
    """Alignment::write_cigar (alignment.rs:95-110) with display.rs:8-41; equal-cost ranges shown invalid."""
    out = []
    for op in ops:
        t = op.type
        if t == OP_TS_ENTRANCE:
            out.append("[TS%s%s%s:[-]:[-]:%d:" % ("RQ"[op.primary], "RQ"[op.secondary], "FR"[op.direction], op.value))
        elif t == OP_TS_EXIT:
            out.append(":%d]" % op.value)
        else:
            out.append("%d%s" % (op.count, "IDX="[t & 3]))
    return "".join(out)


class FlatConfig:
    """tsa_config.Config -> the C struct tsao_config (keeps the arrays alive)."""

    def __init__(self, cfg: tsa_config.Config):
        self.cfg = cfg
        A = len(cfg.chars)
        sub, opn, ext = [], [], []
        for t in cfg.tables:
            for r in range(A):
                sub.extend(t.sub[r])
            opn.extend(t.open)
            ext.extend(t.ext)
        self._sub = (C.c_uint64 * len(sub))(*sub)
        self._open = (C.c_uint64 * len(opn))(*opn)
        self._ext = (C.c_uint64 * len(ext))(*ext)
        self._comp = (C.c_uint8 * A)(*alphabets.complement_table(cfg.alphabet))
        c = _CConfig()
        c.alphabet_size = A
        c.sub = C.cast(self._sub, C.POINTER(C.c_uint64))
        c.open = C.cast(self._open, C.POINTER(C.c_uint64))
        c.ext = C.cast(self._ext, C.POINTER(C.c_uint64))
        for i, b in enumerate(cfg.base):
            c.base[i] = b
        self._fx, self._fc = [], []
        for k, pts in enumerate(cfg.fns):
            xs = (C.c_int64 * len(pts))(*[p[0] for p in pts])
            cs = (C.c_uint64 * len(pts))(*[p[1] for p in pts])
            self._fx.append(xs)
            self._fc.append(cs)
            c.fn_len[k] = len(pts)
            c.fn_x[k] = C.cast(xs, C.POINTER(C.c_int64))
            c.fn_c[k] = C.cast(cs, C.POINTER(C.c_uint64))
        c.left_flank_length = cfg.left_flank_length
        c.right_flank_length = cfg.right_flank_length
        c.complement = C.cast(self._comp, C.POINTER(C.c_uint8))
        self.c = c


def _options(no_ts=False, total_length_maximise=True, min_length_lookahead=True, force_label_correcting=False,
             cost_limit=None, memory_limit=None):
    o = _COptions()
    o.no_ts = int(no_ts)
    o.total_length_maximise = int(total_length_maximise)
    o.min_length_lookahead = int(min_length_lookahead)
    o.force_label_correcting = int(force_label_correcting)
    o.cost_limit = U64_MAX if cost_limit is None else cost_limit
    o.memory_limit = U64_MAX if memory_limit is None else memory_limit
    return o


def _run(fn, flat: FlatConfig, reference: str, query: str, rng, **kw):
    a = flat.cfg.alphabet
    r = alphabets.encode(a, reference)
    q = alphabets.encode(a, query)
    if rng is None:
        rng = (0, len(r), 0, len(q))
    ro, rl, qo, ql = rng
    res = _CResult()
    opt = _options(**kw)
    rc = fn(C.byref(flat.c), r, len(r), q, len(q), ro, rl, qo, ql, C.byref(opt), C.byref(res))
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    ops = [Op(res.ops[i].count, res.ops[i].type, res.ops[i].primary, res.ops[i].secondary, res.ops[i].direction,
              res.ops[i].value) for i in range(res.n_ops)]
    out = Result(RESULT_NAMES[res.result_type], res.cost, ops, res.opened_nodes, res.closed_nodes,
                 res.suboptimal_opened_nodes, res.ts_total_length)
    lib().tsao_result_free(C.byref(res))
    return out


def astar_align(flat, reference, query, rng=None, **kw):
    """Reference-faithful A* (defaults = CLI defaults: lookahead, maximise total length; align.rs:104-122)."""
    return _run(lib().tsao_astar_align, flat, reference, query, rng, **kw)


def dp_align(flat, reference, query, rng=None, **kw):
    """Scalar layered DP (same optimum; traceback under the documented tie-break)."""
    return _run(lib().tsao_dp_align, flat, reference, query, rng, **kw)


def rescore(flat, reference, query, ops, ref_offset=0, qry_offset=0, as_searched=False):
    """compute_cost restatement. Returns (cost, end_ref, end_qry, ok).  as_searched: see tsa_oracle.h."""
    a = flat.cfg.alphabet
    r = alphabets.encode(a, reference)
    q = alphabets.encode(a, query)
    arr = (_COp * max(1, len(ops)))()
    for i, op in enumerate(ops):
        arr[i] = _COp(op.count, op.type, op.primary, op.secondary, op.direction, op.value)
    er, eq, ok = C.c_int64(), C.c_int64(), C.c_int32()
    cost = lib().tsao_rescore_mode(C.byref(flat.c), r, len(r), q, len(q), ref_offset, qry_offset, arr, len(ops),
                                   C.byref(er), C.byref(eq), C.byref(ok), int(as_searched))
    return cost, er.value, eq.value, bool(ok.value)
