"""Host-side mirror of the reference's Python surface (python_bindings/python/tsalign/__init__.py:33-249,
_types.py:7-73) on top of the C ABI: `Aligner`, `align`, `Alignment`, `AlignmentRange`, the op dataclasses --
plus the batch entry points the GPU path is built for (`Aligner.align_batch`, `StagedBatch`)."""
from __future__ import annotations

import ctypes as C
import pathlib
import struct
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple, Union

from . import _lib
from ._lib import TsaError, TsaOp, TsaOptions, TsaPair, TsaResult, U64_MAX

_ALIGNER_KWARG_NAMES = frozenset({"no_ts", "min_length_strategy", "chaining_strategy", "total_length_strategy", "costs", "costs_file"})
_MIN_LENGTH = {"none", "lookahead", "preprocess_price", "preprocess_filter", "preprocess_lookahead"}
_CHAINING = {"none", "lower_bound"}
_TOTAL_LENGTH = {"none", "maximise"}


@dataclass
class AlignmentRange:
    """Coordinate bounds for a pairwise alignment (_types.py:7-14)."""
    reference_start: int = 0
    reference_end: "int | None" = None
    query_start: int = 0
    query_end: "int | None" = None


@dataclass
class SimpleAlignmentOp:
    kind: str


@dataclass
class TemplateSwitchEntranceOp:
    kind: str
    first_offset: int
    primary: str
    secondary: str
    direction: str
    equal_cost_range: dict


@dataclass
class TemplateSwitchExitOp:
    kind: str
    anti_primary_gap: int


AlignmentOp = Union[SimpleAlignmentOp, TemplateSwitchEntranceOp, TemplateSwitchExitOp]
_INVALID_RANGE = {"min_start": 1, "max_start": -1, "min_end": 1, "max_end": -1}  # EqualCostRange::new_invalid()


class Config:
    """A parsed cost model (lib_tsalign::config::TemplateSwitchConfig)."""

    def __init__(self, text: Optional[str] = None, alphabet: str = "dna-n", lib=None):
        self._lib = lib or _lib.default()
        if alphabet not in _lib.ALPHABETS:
            raise ValueError(f"unknown alphabet {alphabet!r}")
        self.alphabet = alphabet
        if text is None:
            self._h = self._lib.tsa_config_default(_lib.ALPHABETS[alphabet])
            if not self._h:
                raise TsaError(6, "unknown alphabet")
        else:
            status = C.c_int(0)
            err = C.create_string_buffer(512)
            raw = text.encode()
            self._h = self._lib.tsa_config_parse(raw, len(raw), _lib.ALPHABETS[alphabet], C.byref(status), err, len(err))
            if not self._h:
                raise TsaError(status.value, err.value.decode(errors="replace"))

    def text(self) -> str:
        n = self._lib.tsa_config_write(self._h, None, 0)
        buf = C.create_string_buffer(n)
        self._lib.tsa_config_write(self._h, buf, n)
        return buf.value.decode()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.tsa_config_free(h)


def _clean(seq) -> bytes:
    # python_bindings/src/lib.rs:53-56: str(obj).encode()
    return seq if isinstance(seq, bytes) else str(seq).encode()


def _make_pairs(pairs: Sequence[tuple]):
    """[(reference, query) | (reference, query, (ro, rl, qo, ql))] -> (TsaPair array, keep-alive list)."""
    arr = (TsaPair * max(1, len(pairs)))()
    keep = []
    for i, p in enumerate(pairs):
        r, q = _clean(p[0]), _clean(p[1])
        rng = p[2] if len(p) > 2 and p[2] is not None else (0, None, 0, None)
        keep.append((r, q))
        arr[i].reference, arr[i].reference_len = r, len(r)
        arr[i].query, arr[i].query_len = q, len(q)
        arr[i].reference_offset = rng[0] or 0
        arr[i].reference_limit = -1 if rng[1] is None else rng[1]
        arr[i].query_offset = rng[2] or 0
        arr[i].query_limit = -1 if rng[3] is None else rng[3]
    return arr, keep


def _options(no_ts=False, device=0, cost_limit=None, memory_limit=None, max_template_switches=0, first_threshold=0, traceback=True, scout=False, dev_flags=0, postprocess=0,
             flags=0, total_length_strategy="maximise", descendant_strategy="any", force_label_correcting=False) -> TsaOptions:
    o = TsaOptions()
    o.no_ts = int(bool(no_ts))
    o.device = device
    o.cost_limit = U64_MAX if cost_limit is None else int(cost_limit)
    o.memory_limit = U64_MAX if memory_limit is None else int(memory_limit)
    o.max_template_switches = max_template_switches
    o.first_threshold = first_threshold
    o.no_traceback = 0 if traceback else 1
    o.reserved = (1 if scout else 0) | (dev_flags & ~1)   # developer knobs of the engine (tsa_capi.cpp: engine_options)
    o.postprocess = int(postprocess)
    o.flags = int(flags)
    o.total_length_strategy = {"maximise": 0, "none": 1}[total_length_strategy]
    o.descendant_strategy = {"any": 0, "allow-only-all-equal": 1}[descendant_strategy]
    o.force_label_correcting = int(bool(force_label_correcting))
    return o


_OP_STRUCT = struct.Struct("<qiiiiqbbbbi")  # tsa_op: count, type, primary, secondary, direction, value, equal-cost range, reserved
POST_EXTEND_BEYOND_RANGE, POST_EQUAL_COST_RANGES = 1, 2   # TSA_POST_* of include/tsalign_b200.h
FLAG_KEEP_FLANK_RUNS = 1                                   # TSA_FLAG_* of include/tsalign_b200.h


class BatchResult:
    """One entry of a batch: the C struct tsa_result, copied out.  `ops` (the run-length encoded alignment as
    (count, type, primary, secondary, direction, value) tuples) is decoded on first use."""
    __slots__ = ("status", "result_type", "cost", "template_switches", "message", "duration_seconds", "range", "_raw", "_ops", "_ecr")

    def __init__(self, status, result_type, cost, template_switches, raw_ops, message="", duration_seconds=0.0, rng=None):
        self.status, self.result_type, self.cost, self.template_switches = status, result_type, cost, template_switches
        self.message, self.duration_seconds = message, duration_seconds
        self.range = rng                      # (reference_offset, reference_limit, query_offset, query_limit) the ops span
        self._raw, self._ops, self._ecr = raw_ops, None, None

    def _decode(self):
        if self._ops is None and self._raw is not None:
            rows = list(_OP_STRUCT.iter_unpack(self._raw))
            self._ops = [r[:6] for r in rows]
            self._ecr = [r[6:10] if r[1] == 12 else None for r in rows]

    @property
    def ops(self) -> Optional[List[Tuple[int, int, int, int, int, int]]]:
        self._decode()
        return self._ops

    @property
    def equal_cost_ranges(self) -> Optional[List[Optional[Tuple[int, int, int, int]]]]:
        """Per op: (min_start, max_start, min_end, max_end) of a template switch entrance, else None."""
        self._decode()
        return self._ecr

    @property
    def found(self) -> bool:
        return self.status == 0 and self.result_type == "FoundTarget"

    def __repr__(self):
        return (f"BatchResult(status={self.status}, result_type={self.result_type!r}, cost={self.cost}, "
                f"template_switches={self.template_switches}, ops={self.ops}, message={self.message!r})")


def _copy_results(lib, res, n) -> List[BatchResult]:
    out = []
    names = _lib.RESULT_NAMES
    size = C.sizeof(TsaOp)
    for i in range(n):
        r = res[i]
        raw = C.string_at(r.ops, r.n_ops * size) if r.ops else None
        out.append(BatchResult(r.status, names[r.result_type], r.cost, r.template_switches, raw,
                               r.message.decode(errors="replace") if r.status else "", r.duration_seconds,
                               (r.reference_offset, r.reference_limit, r.query_offset, r.query_limit)))
    lib.tsa_results_free(res, n)
    return out


def cigar_of(ops, ranges=None) -> str:
    """Alignment::cigar (alignment.rs:95-110, template_switch_distance/display.rs:8-41,80-94)."""
    out = []
    for idx, (count, t, p, s, d, v) in enumerate(ops):
        if t == 12:
            e = ranges[idx] if ranges else None
            valid = e is not None and e[0] <= e[1] and e[2] <= e[3]      # EqualCostRange::is_valid
            rng = "[%d,%d]:[%d,%d]" % tuple(e) if valid else "[-]:[-]"
            out.append("[TS%s%s%s:%s:%d:" % ("RQ"[p], "RQ"[s], "FR"[d], rng, v))
        elif t == 13:
            out.append(":%d]" % v)
        else:
            out.append("%d%s" % (count, "IDX="[t & 3]))
    return "".join(out)


def _ops_array(ops, ranges=None, extra=16):
    """[(count, type, primary, secondary, direction, value)] (+ equal-cost ranges) -> (TsaOp array with spare capacity, n)."""
    arr = (TsaOp * (len(ops) + extra))()
    for i, o in enumerate(ops):
        a = arr[i]
        a.count, a.type, a.primary, a.secondary, a.direction, a.value = o[:6]
        e = (ranges[i] if ranges and ranges[i] is not None else (1, -1, 1, -1)) if o[1] == 12 else (0, 0, 0, 0)
        a.min_start, a.max_start, a.min_end, a.max_end = e
    return arr


def _ops_list(arr, n):
    ops = [(arr[i].count, arr[i].type, arr[i].primary, arr[i].secondary, arr[i].direction, arr[i].value) for i in range(n)]
    ranges = [(arr[i].min_start, arr[i].max_start, arr[i].min_end, arr[i].max_end) if arr[i].type == 12 else None for i in range(n)]
    return ops, ranges


def postprocess(config: "Config", reference, query, ops, rng, flags: int, ranges=None):
    """Host-only: what the reference does to a found alignment after the search (a_star_aligner.rs:238-253).
    rng = (reference_offset, reference_limit, query_offset, query_limit).  Returns (ops, equal-cost ranges, range, cost)."""
    lib = config._lib
    arr_p, keep = _make_pairs([(reference, query)])
    arr = _ops_array(ops, ranges, extra=len(ops) + 64)
    n = C.c_size_t(len(ops))
    r = [C.c_int64(x) for x in rng]
    cost = C.c_uint64(0)
    rc = lib.tsa_postprocess(config._h, arr_p, flags, arr, C.byref(n), len(arr), C.byref(r[0]), C.byref(r[1]), C.byref(r[2]), C.byref(r[3]), C.byref(cost))
    if rc != 0:
        raise TsaError(rc, "tsa_postprocess")
    out_ops, out_ranges = _ops_list(arr, n.value)
    return out_ops, out_ranges, tuple(x.value for x in r), cost.value


def post_move(config: "Config", reference, query, which: int, ops, ref_offset: int, qry_offset: int, compact_index: int):
    """Host-only single move of a template switch boundary (tsa_post_move).  Returns (moved, ops, compact_index, cost)."""
    lib = config._lib
    arr_p, keep = _make_pairs([(reference, query)])
    arr = _ops_array(ops, None, extra=8)
    n = C.c_size_t(len(ops))
    ci = C.c_size_t(compact_index)
    cost = C.c_uint64(0)
    rc = lib.tsa_post_move(config._h, arr_p, which, arr, C.byref(n), len(arr), ref_offset, qry_offset, C.byref(ci), C.byref(cost))
    if rc < 0:
        raise TsaError(-rc, "tsa_post_move")
    return bool(rc), _ops_list(arr, n.value)[0], ci.value, cost.value


class Alignment:
    """Result of one pairwise alignment (mirror of tsalign.Alignment)."""

    def __init__(self, result: BatchResult, reference: bytes, query: bytes, names: Tuple[str, str], rng, alphabet: str):
        self._r = result
        self._reference, self._query = reference, query
        self._names = names
        self._range = rng
        self._alphabet = alphabet

    @property
    def cost(self) -> int:
        return self._r.cost

    def cigar(self) -> Optional[str]:
        if not self._r.found or self._r.ops is None:
            return None
        return cigar_of(self._r.ops, self._r.equal_cost_ranges)

    def alignments(self) -> Optional[List[Tuple[int, AlignmentOp]]]:
        if not self._r.found or self._r.ops is None:
            return None
        out = []
        for (count, t, p, s, d, v), e in zip(self._r.ops, self._r.equal_cost_ranges):
            if t == 12:
                ecr = dict(zip(("min_start", "max_start", "min_end", "max_end"), e)) if e is not None else dict(_INVALID_RANGE)
                out.append((count, TemplateSwitchEntranceOp("TemplateSwitchEntrance", v, ["Reference", "Query"][p], ["Reference", "Query"][s],
                                                            ["Forward", "Reverse"][d], ecr)))
            elif t == 13:
                out.append((count, TemplateSwitchExitOp("TemplateSwitchExit", v)))
            else:
                out.append((count, SimpleAlignmentOp(_lib.OP_NAMES[t])))
        return out

    def stats(self) -> dict:
        """AlignmentStatistics (alignment_result.rs:53-81,194-227); the A* node counters have no meaning for a dense fill."""
        r = self._r
        n, m = len(self._reference), len(self._query)
        ts = sum(1 for op in (r.ops or []) if op[1] == 13)
        result = {"astar_result_type": r.result_type}
        if r.result_type == "FoundTarget":
            result["cost"] = r.cost
        elif r.result_type == "ExceededCostLimit":
            result["cost_limit"] = r.cost
        return {
            "result": result,
            "sequences": {"reference_name": self._names[0], "reference": self._reference.decode(), "query_name": self._names[1], "query": self._query.decode()},
            "reference_offset": self._range[0], "query_offset": self._range[2],
            "cost": float(r.cost), "cost_per_base": (2.0 * r.cost / (n + m)) if n + m else 0.0,
            "duration_seconds": r.duration_seconds, "opened_nodes": 0.0, "closed_nodes": 0.0, "suboptimal_opened_nodes": 0.0,
            "suboptimal_opened_nodes_ratio": 0.0, "template_switch_amount": float(ts), "runtime": 0.0, "memory": 0.0,
        }


class Aligner:
    """Pairwise sequence aligner with template-switch detection on a B200 (mirror of tsalign.Aligner).

    The search-heuristic arguments of the reference (`min_length_strategy`, `chaining_strategy`) are validated and
    ignored: they never change the optimal cost.  `total_length_strategy` only affects which of several
    equal-cost alignments the reference returns; this implementation uses its own documented tie-break.
    Defaults match the reference: alphabet dna-n, Rust `TemplateSwitchConfig::default()` costs.
    """

    def __init__(self, *, no_ts: bool = False, min_length_strategy: str = "lookahead", chaining_strategy: str = "none",
                 total_length_strategy: str = "maximise", costs: Optional[str] = None,
                 costs_file: Optional[Union[str, pathlib.Path]] = None, alphabet: str = "dna-n", device: int = 0,
                 first_threshold: int = 0, traceback: bool = True, scout: bool = False, dev_flags: int = 0,
                 postprocess: Optional[int] = None, flags: int = 0, descendant_strategy: str = "any",
                 force_label_correcting: bool = False, max_template_switches: int = 0, lib=None) -> None:
        if costs is not None and costs_file is not None:
            raise ValueError("Provide at most one of 'costs' or 'costs_file'.")
        if min_length_strategy not in _MIN_LENGTH:
            raise ValueError(f"unknown min_length_strategy {min_length_strategy!r}")
        if chaining_strategy not in _CHAINING:
            raise ValueError(f"unknown chaining_strategy {chaining_strategy!r}")
        if total_length_strategy not in _TOTAL_LENGTH:
            raise ValueError(f"unknown total_length_strategy {total_length_strategy!r}")
        if costs_file is not None:
            costs = pathlib.Path(costs_file).read_text()
        self._lib = lib or _lib.default()
        self.no_ts = bool(no_ts)
        self.device = device
        self.traceback = bool(traceback)        # False: optimal costs only
        self.scout = bool(scout)                # tuning of the exact pruning only
        # what happens after the search (a_star_aligner.rs:238-253): None = like the reference's callers -- align() extends
        # beyond the range and computes equal-cost ranges (python_bindings/src/lib.rs:124-133), align_batch() returns the
        # searched alignments as they are; an int (POST_* bits) applies to both
        self.postprocess = postprocess
        self.flags = int(flags)                 # FLAG_* bits
        self.total_length_strategy = total_length_strategy
        if descendant_strategy not in ("any", "allow-only-all-equal"):
            raise ValueError(f"unknown descendant_strategy {descendant_strategy!r}")
        self.descendant_strategy = descendant_strategy
        self.force_label_correcting = bool(force_label_correcting)
        self.max_template_switches = int(max_template_switches)
        self.dev_flags = int(dev_flags)         # developer knobs (2: no column windows for medium pairs; 4: small windows, emulator only; 8 / 16: --no-ts alignments always through checkpoints / the code matrix; 32: first window stage fused; 64: pairs wider than 31 through the tiled window stage only; 128: one warp per pair in the primary fill)
        self.first_threshold = first_threshold  # tuning of the exact pruning only; results do not depend on it
        self.config = Config(costs, alphabet, lib=self._lib)

    def _strategy_kwargs(self):
        return dict(flags=self.flags, total_length_strategy=self.total_length_strategy, descendant_strategy=self.descendant_strategy,
                    force_label_correcting=self.force_label_correcting, max_template_switches=self.max_template_switches)

    # -- batch entry point: the call the GPU path is built for ----------------------------------------------
    def align_batch(self, pairs: Sequence[tuple], *, cost_limit: Optional[int] = None, memory_limit: Optional[int] = None,
                    postprocess: Optional[int] = None) -> List[BatchResult]:
        """Align independent pairs: [(reference, query) or (reference, query, (ref_offset, ref_limit, qry_offset, qry_limit))]."""
        arr, keep = _make_pairs(pairs)
        res = (TsaResult * max(1, len(pairs)))()
        err = C.create_string_buffer(512)
        post = postprocess if postprocess is not None else (self.postprocess or 0)
        opt = _options(self.no_ts, self.device, cost_limit, memory_limit, first_threshold=self.first_threshold, traceback=self.traceback, scout=self.scout,
                       dev_flags=self.dev_flags, postprocess=post, **self._strategy_kwargs())
        rc = self._lib.tsa_align_batch(self.config._h, C.byref(opt), arr, len(pairs), res, err, len(err))
        if rc != 0:
            raise TsaError(rc, err.value.decode(errors="replace"))
        del keep
        return _copy_results(self._lib, res, len(pairs))

    def align(self, reference: object, query: object, *, reference_name: str = "reference", query_name: str = "query",
              range: Optional[AlignmentRange] = None, reference_start: Optional[int] = None, reference_limit: Optional[int] = None,
              query_start: Optional[int] = None, query_limit: Optional[int] = None, cost_limit: Optional[int] = None,
              memory_limit: Optional[int] = None) -> Optional[Alignment]:
        """Align two sequences.  Returns None when no target was found within the limits (lib.rs:135-141)."""
        if range is not None:
            reference_start, reference_limit = range.reference_start, range.reference_end
            query_start, query_limit = range.query_start, range.query_end
        r, q = _clean(reference), _clean(query)
        rng = (reference_start or 0, reference_limit, query_start or 0, query_limit)
        post = self.postprocess if self.postprocess is not None else POST_EXTEND_BEYOND_RANGE | POST_EQUAL_COST_RANGES
        res = self.align_batch([(r, q, rng)], cost_limit=cost_limit, memory_limit=memory_limit, postprocess=post)[0]
        if res.status != 0:
            raise TsaError(res.status, res.message)
        if not res.found:
            return None
        return Alignment(res, r, q, (reference_name, query_name), res.range, self.config.alphabet)


def align(reference: object, query: object, **kwargs: object) -> Optional[Alignment]:
    """One-call convenience wrapper (mirror of tsalign.align)."""
    aligner_kwargs = {k: v for k, v in kwargs.items() if k in _ALIGNER_KWARG_NAMES or k in ("alphabet", "device", "lib", "first_threshold", "traceback", "scout", "dev_flags", "postprocess", "flags", "descendant_strategy", "force_label_correcting", "max_template_switches")}
    align_kwargs = {k: v for k, v in kwargs.items() if k not in aligner_kwargs}
    return Aligner(**aligner_kwargs).align(reference, query, **align_kwargs)


class StagedBatch:
    """Inputs resident in HBM; `run()` executes the kernels only (what bench.py times as `value`)."""

    def __init__(self, aligner: Aligner, pairs: Sequence[tuple], *, cost_limit=None, memory_limit=None):
        self._lib = aligner._lib
        self._aligner = aligner
        self.n = len(pairs)
        arr, keep = _make_pairs(pairs)
        status = C.c_int(0)
        err = C.create_string_buffer(512)
        opt = _options(aligner.no_ts, aligner.device, cost_limit, memory_limit, first_threshold=aligner.first_threshold, traceback=aligner.traceback, scout=aligner.scout,
                       dev_flags=aligner.dev_flags, postprocess=aligner.postprocess or 0, **aligner._strategy_kwargs())
        self._h = self._lib.tsa_batch_create(aligner.config._h, C.byref(opt), arr, self.n, C.byref(status), err, len(err))
        if not self._h:
            raise TsaError(status.value, err.value.decode(errors="replace"))

    def run(self) -> None:
        rc = self._lib.tsa_batch_run(self._h)
        if rc != 0:
            raise TsaError(rc, "tsa_batch_run")

    def fetch(self) -> List[BatchResult]:
        res = (TsaResult * max(1, self.n))()
        rc = self._lib.tsa_batch_fetch(self._h, res)
        if rc != 0:
            raise TsaError(rc, "tsa_batch_fetch")
        return _copy_results(self._lib, res, self.n)

    def stats(self) -> dict:
        a, b, c, e, f = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        d = C.c_int32()
        self._lib.tsa_batch_stats(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e), C.byref(f))
        return {"launches": a.value, "jump_launches": b.value, "fill_launches": c.value, "layers": d.value, "h2d_bytes": e.value, "d2h_bytes": f.value}

    def timing(self) -> dict:
        """Device time of the last run() per kernel family (CUDA events on the engine's stream)."""
        j, f = C.c_double(), C.c_double()
        self._lib.tsa_batch_timing(self._h, C.byref(j), C.byref(f))
        w = [C.c_int64() for _ in range(4)]
        self._lib.tsa_batch_work(self._h, *[C.byref(x) for x in w])
        return {"jump_ms": j.value, "fill_ms": f.value, "chains_started": w[0].value, "chains_run": w[1].value,
                "rows_filled": w[2].value, "rows_jumped": w[3].value}

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.tsa_batch_free(h)

    def __del__(self):
        self.close()


# ---- one very long pair without template switches, column-banded over several GPUs (tsa_align_long / tsa_long_*) ------------------
def _long_stats(s) -> dict:
    return {"forward_ms": s.forward_ms, "trace_ms": s.trace_ms, "tiles": s.tiles, "tile_cells": s.tile_cells,
            "boundary_bytes_out": s.boundary_bytes_out, "resident_bytes": s.resident_bytes, "interval": s.interval, "group": s.group,
            "speculated_tiles": s.speculated_tiles, "speculated_used": s.speculated_used, "speculate_ms": s.speculate_ms}


def align_long(aligner: Aligner, reference, query, *, devices: Optional[Sequence[int]] = None, interval: int = 0, group: int = 0,
               rng=None, cost_limit: Optional[int] = None, memory_limit: Optional[int] = None, postprocess: int = 0):
    """One pair without template switches (aligner.no_ts), its query columns cut into one band per device of THIS process; the
    boundary columns stream between the devices over NVLink peer stores.  `memory_limit` (bytes per device) bounds what is resident:
    checkpoint rows / boundary columns and the codes of one tile instead of a code matrix.  Returns (BatchResult, [stats per band])."""
    lib = aligner._lib
    devices = list(devices) if devices else [aligner.device]
    arr, keep = _make_pairs([(reference, query, rng)])
    res = (TsaResult * 1)()
    stats = (_lib.TsaLongStats * len(devices))()
    devs = (C.c_int32 * len(devices))(*devices)
    err = C.create_string_buffer(512)
    opt = _options(aligner.no_ts, devices[0], cost_limit, memory_limit, traceback=aligner.traceback, postprocess=postprocess, **aligner._strategy_kwargs())
    rc = lib.tsa_align_long(aligner.config._h, C.byref(opt), arr, devs, len(devices), interval, group, res, stats, err, len(err))
    if rc != 0:
        raise TsaError(rc, err.value.decode(errors="replace"))
    del keep
    return _copy_results(lib, res, 1)[0], [_long_stats(s) for s in stats]


class LongBand:
    """One band of a long pair in this process (rank `rank` of `world`; torchrun: one process per GPU).  See include/tsalign_b200.h."""

    def __init__(self, aligner: Aligner, reference, query, rank: int, world: int, *, interval: int = 0, group: int = 0, rng=None,
                 memory_limit: Optional[int] = None, postprocess: int = 0):
        self._lib = aligner._lib
        self.rank, self.world = rank, world
        arr, keep = _make_pairs([(reference, query, rng)])
        status = C.c_int(0)
        err = C.create_string_buffer(512)
        opt = _options(aligner.no_ts, aligner.device, None, memory_limit, traceback=aligner.traceback, postprocess=postprocess, **aligner._strategy_kwargs())
        self._h = self._lib.tsa_long_create(aligner.config._h, C.byref(opt), arr, rank, world, interval, group, C.byref(status), err, len(err))
        if not self._h:
            raise TsaError(status.value, err.value.decode(errors="replace"))
        rows, cols = C.c_int64(), C.c_int64()
        self._lib.tsa_long_dims(self._h, C.byref(rows), C.byref(cols))
        self.rows, self.columns = rows.value, cols.value

    def export_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        rc = self._lib.tsa_long_ipc_export(self._h, buf)
        if rc != 0:
            raise TsaError(rc, "tsa_long_ipc_export")
        return buf.raw

    def connect(self, handle: bytes) -> None:
        rc = self._lib.tsa_long_ipc_connect(self._h, C.create_string_buffer(handle, 64))
        if rc != 0:
            raise TsaError(rc, "tsa_long_ipc_connect (CUDA IPC between the ranks' devices)")

    def forward(self) -> None:
        rc = self._lib.tsa_long_forward(self._h)
        if rc != 0:
            raise TsaError(rc, "tsa_long_forward")

    def cost(self):
        c, t = C.c_uint64(), C.c_int32()
        rc = self._lib.tsa_long_cost(self._h, C.byref(c), C.byref(t))
        if rc != 0:
            raise TsaError(rc, "tsa_long_cost")
        return c.value, _lib.RESULT_NAMES[t.value]

    def owner(self, column: int) -> int:
        return self._lib.tsa_long_owner(self._h, column)

    def walk(self, state: tuple):
        """state = (i, j, g, need, cost); returns (state, status, unit ops in walk order as bytes)."""
        st = _lib.TsaLongWalkState(state[0], state[1], state[2], state[3], state[4], 0, 0)
        cap = self.rows + self.columns + 64
        buf = C.create_string_buffer(cap)
        n = C.c_size_t(0)
        rc = self._lib.tsa_long_walk(self._h, C.byref(st), buf, cap, C.byref(n))
        if rc != 0:
            raise TsaError(rc, "tsa_long_walk")
        return (st.i, st.j, st.g, st.need, st.cost), st.status, buf.raw[:n.value]

    def result(self, cost: int, ops_walk_order: bytes) -> BatchResult:
        res = (TsaResult * 1)()
        rc = self._lib.tsa_long_result(self._h, cost, ops_walk_order, len(ops_walk_order), res)
        if rc != 0:
            raise TsaError(rc, "tsa_long_result")
        return _copy_results(self._lib, res, 1)[0]

    def stats(self) -> dict:
        s = _lib.TsaLongStats()
        self._lib.tsa_long_get_stats(self._h, C.byref(s))
        return _long_stats(s)

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.tsa_long_free(h)

    def __del__(self):
        self.close()


def run_long_bands(bands: Sequence[LongBand], exchange=None):
    """Drives the protocol of LongBand over bands that all live in this process (tests; `exchange` is unused).  A multi-process
    launch does the same steps with its own exchange of the 64-byte handles and of the walk state (bench.py --config c5)."""
    world = len(bands)
    for r in range(world - 1):
        bands[r].connect(bands[r + 1].export_handle())
    for b in bands:
        b.forward()
    cost, kind = bands[-1].cost()
    if kind != "FoundTarget":
        return None, kind
    state, ops, r = (bands[-1].rows, bands[-1].columns, 0, 1, cost), b"", world - 1
    while True:
        state, status, seg = bands[r].walk(state)
        ops += seg
        if status == 1:
            break
        if status < 0:
            raise TsaError(11, f"walk failed with status {status}")
        r = bands[r].owner(state[1])
    return bands[-1].result(cost, ops), kind
