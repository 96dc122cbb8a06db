"""ctypes binding of include/tsalign_b200.h.

`default()` loads the in-tree CUDA library (libtsalign_b200.so, built by `__graft_entry__.build()` or
`make -C template_switch_aligner_b200/csrc`).  There is no CPU implementation behind this package: if the
library is missing, or no CUDA device is present, the calls fail loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TSA_B200_LIB: developer knob to benchmark another *CUDA build* of the same sources (e.g. different launch bounds)
LIB_PATH = os.environ.get("TSA_B200_LIB") or os.path.join(_HERE, "libtsalign_b200.so")
U64_MAX = (1 << 64) - 1

TSA_OK = 0
STATUS_NAMES = {
    0: "TSA_OK", 1: "TSA_ERR_NO_DEVICE", 2: "TSA_ERR_CONFIG_PARSE", 3: "TSA_ERR_NOT_V_SHAPED_RQQR",
    4: "TSA_ERR_NOT_V_SHAPED_RRQQ", 5: "TSA_ERR_NOT_V_SHAPED_LENDIFF", 6: "TSA_ERR_ALPHABET", 7: "TSA_ERR_INVALID_CHAR",
    8: "TSA_ERR_INVALID_RANGE", 9: "TSA_ERR_UNSUPPORTED", 10: "TSA_ERR_ARGUMENT", 11: "TSA_ERR_INTERNAL",
}
RESULT_NAMES = ["FoundTarget", "ExceededCostLimit", "ExceededMemoryLimit", "NoTarget"]
ALPHABETS = {"dna": 0, "dna-n": 1, "rna": 2, "rna-n": 3, "dna-iupac": 4, "rna-iupac": 5}
OP_NAMES = [
    "PrimaryInsertion", "PrimaryDeletion", "PrimarySubstitution", "PrimaryMatch",
    "PrimaryFlankInsertion", "PrimaryFlankDeletion", "PrimaryFlankSubstitution", "PrimaryFlankMatch",
    "SecondaryInsertion", "SecondaryDeletion", "SecondarySubstitution", "SecondaryMatch",
    "TemplateSwitchEntrance", "TemplateSwitchExit",
]
EXPORTS = [
    "tsa_config_parse", "tsa_config_default", "tsa_config_write", "tsa_config_free", "tsa_config_alphabet",
    "tsa_align_batch", "tsa_results_free", "tsa_batch_create", "tsa_batch_run", "tsa_batch_fetch", "tsa_batch_stats",
    "tsa_batch_free", "tsa_batch_timing", "tsa_batch_work", "tsa_measure_addmin_peak", "tsa_postprocess", "tsa_post_move", "tsa_device_count", "tsa_version",
    "tsa_align_long", "tsa_long_create", "tsa_long_ipc_export", "tsa_long_ipc_connect", "tsa_long_forward", "tsa_long_cost", "tsa_long_owner",
    "tsa_long_walk", "tsa_long_result", "tsa_long_get_stats", "tsa_long_dims", "tsa_long_free",
]


class TsaOptions(C.Structure):
    _fields_ = [("no_ts", C.c_int32), ("device", C.c_int32), ("cost_limit", C.c_uint64), ("memory_limit", C.c_uint64),
                ("max_template_switches", C.c_int32), ("first_threshold", C.c_int32), ("no_traceback", C.c_int32), ("reserved", C.c_int32),
                ("postprocess", C.c_int32), ("flags", C.c_int32), ("total_length_strategy", C.c_int32), ("descendant_strategy", C.c_int32),
                ("force_label_correcting", C.c_int32), ("reserved2", C.c_int32)]


class TsaPair(C.Structure):
    _fields_ = [("reference", C.c_char_p), ("reference_len", C.c_size_t), ("query", C.c_char_p), ("query_len", C.c_size_t),
                ("reference_offset", C.c_int64), ("reference_limit", C.c_int64), ("query_offset", C.c_int64), ("query_limit", C.c_int64)]


class TsaOp(C.Structure):
    _fields_ = [("count", C.c_int64), ("type", C.c_int32), ("primary", C.c_int32), ("secondary", C.c_int32),
                ("direction", C.c_int32), ("value", C.c_int64),
                ("min_start", C.c_int8), ("max_start", C.c_int8), ("min_end", C.c_int8), ("max_end", C.c_int8), ("reserved", C.c_int32)]


class TsaResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("result_type", C.c_int32), ("cost", C.c_uint64), ("template_switches", C.c_int32),
                ("reserved", C.c_int32), ("ops", C.POINTER(TsaOp)), ("n_ops", C.c_size_t), ("duration_seconds", C.c_double),
                ("message", C.c_char * 96),
                ("reference_offset", C.c_int64), ("reference_limit", C.c_int64), ("query_offset", C.c_int64), ("query_limit", C.c_int64)]


class TsaError(RuntimeError):
    def __init__(self, status, message=""):
        self.status = status
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")


class TsaLongStats(C.Structure):
    _fields_ = [("forward_ms", C.c_double), ("trace_ms", C.c_double), ("tiles", C.c_int64), ("tile_cells", C.c_int64),
                ("boundary_bytes_out", C.c_int64), ("resident_bytes", C.c_int64), ("interval", C.c_int32), ("group", C.c_int32),
                ("speculated_tiles", C.c_int64), ("speculated_used", C.c_int64), ("speculate_ms", C.c_double)]


class TsaLongWalkState(C.Structure):
    _fields_ = [("i", C.c_int32), ("j", C.c_int32), ("g", C.c_int32), ("need", C.c_int32), ("cost", C.c_int64), ("status", C.c_int32), ("reserved", C.c_int32)]


def bind(cdll):
    """Declare the prototypes of include/tsalign_b200.h on a loaded library."""
    cdll.tsa_config_parse.restype = C.c_void_p
    cdll.tsa_config_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
    cdll.tsa_config_default.restype = C.c_void_p
    cdll.tsa_config_default.argtypes = [C.c_int]
    cdll.tsa_config_write.restype = C.c_size_t
    cdll.tsa_config_write.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    cdll.tsa_config_free.restype = None
    cdll.tsa_config_free.argtypes = [C.c_void_p]
    cdll.tsa_config_alphabet.restype = C.c_int
    cdll.tsa_config_alphabet.argtypes = [C.c_void_p]
    cdll.tsa_align_batch.restype = C.c_int
    cdll.tsa_align_batch.argtypes = [C.c_void_p, C.POINTER(TsaOptions), C.POINTER(TsaPair), C.c_size_t, C.POINTER(TsaResult), C.c_char_p, C.c_size_t]
    cdll.tsa_results_free.restype = None
    cdll.tsa_results_free.argtypes = [C.POINTER(TsaResult), C.c_size_t]
    cdll.tsa_batch_create.restype = C.c_void_p
    cdll.tsa_batch_create.argtypes = [C.c_void_p, C.POINTER(TsaOptions), C.POINTER(TsaPair), C.c_size_t, C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
    cdll.tsa_batch_run.restype = C.c_int
    cdll.tsa_batch_run.argtypes = [C.c_void_p]
    cdll.tsa_batch_fetch.restype = C.c_int
    cdll.tsa_batch_fetch.argtypes = [C.c_void_p, C.POINTER(TsaResult)]
    cdll.tsa_batch_stats.restype = None
    cdll.tsa_batch_stats.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 3 + [C.POINTER(C.c_int32)] + [C.POINTER(C.c_int64)] * 2
    cdll.tsa_batch_timing.restype = None
    cdll.tsa_batch_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    cdll.tsa_batch_work.restype = None
    cdll.tsa_batch_work.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 4
    cdll.tsa_measure_addmin_peak.restype = C.c_int
    cdll.tsa_measure_addmin_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    cdll.tsa_batch_free.restype = None
    cdll.tsa_batch_free.argtypes = [C.c_void_p]
    cdll.tsa_postprocess.restype = C.c_int
    cdll.tsa_postprocess.argtypes = [C.c_void_p, C.POINTER(TsaPair), C.c_int32, C.POINTER(TsaOp), C.POINTER(C.c_size_t), C.c_size_t] + [C.POINTER(C.c_int64)] * 4 + [C.POINTER(C.c_uint64)]
    cdll.tsa_post_move.restype = C.c_int
    cdll.tsa_post_move.argtypes = [C.c_void_p, C.POINTER(TsaPair), C.c_int, C.POINTER(TsaOp), C.POINTER(C.c_size_t), C.c_size_t, C.c_int64, C.c_int64,
                                   C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
    cdll.tsa_align_long.restype = C.c_int
    cdll.tsa_align_long.argtypes = [C.c_void_p, C.POINTER(TsaOptions), C.POINTER(TsaPair), C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32,
                                    C.POINTER(TsaResult), C.POINTER(TsaLongStats), C.c_char_p, C.c_size_t]
    cdll.tsa_long_create.restype = C.c_void_p
    cdll.tsa_long_create.argtypes = [C.c_void_p, C.POINTER(TsaOptions), C.POINTER(TsaPair), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
    cdll.tsa_long_ipc_export.restype = C.c_int
    cdll.tsa_long_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
    cdll.tsa_long_ipc_connect.restype = C.c_int
    cdll.tsa_long_ipc_connect.argtypes = [C.c_void_p, C.c_void_p]
    cdll.tsa_long_forward.restype = C.c_int
    cdll.tsa_long_forward.argtypes = [C.c_void_p]
    cdll.tsa_long_cost.restype = C.c_int
    cdll.tsa_long_cost.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
    cdll.tsa_long_owner.restype = C.c_int
    cdll.tsa_long_owner.argtypes = [C.c_void_p, C.c_int64]
    cdll.tsa_long_walk.restype = C.c_int
    cdll.tsa_long_walk.argtypes = [C.c_void_p, C.POINTER(TsaLongWalkState), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    cdll.tsa_long_result.restype = C.c_int
    cdll.tsa_long_result.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.POINTER(TsaResult)]
    cdll.tsa_long_get_stats.restype = None
    cdll.tsa_long_get_stats.argtypes = [C.c_void_p, C.POINTER(TsaLongStats)]
    cdll.tsa_long_dims.restype = None
    cdll.tsa_long_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    cdll.tsa_long_free.restype = None
    cdll.tsa_long_free.argtypes = [C.c_void_p]
    cdll.tsa_device_count.restype = C.c_int
    cdll.tsa_device_count.argtypes = []
    cdll.tsa_version.restype = C.c_char_p
    cdll.tsa_version.argtypes = []
    return cdll


_DEFAULT = None


def default():
    """The product library.  Raises if it has not been built -- there is nothing to fall back to."""
    global _DEFAULT
    if _DEFAULT is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
                "or make -C template_switch_aligner_b200/csrc).  tsalign_b200 has no CPU implementation.")
        _DEFAULT = bind(C.CDLL(LIB_PATH))
    return _DEFAULT
