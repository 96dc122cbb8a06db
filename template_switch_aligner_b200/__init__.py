"""tsalign-b200: B200-native template-switch aligner (drop-in for the alignment path of
sebschmi/template-switch-aligner).  The arithmetic lives in csrc/ (hand-written sm_100a CUDA behind the C ABI of
include/tsalign_b200.h); this package is the Python mirror of the reference's `tsalign` module."""
from .api import (Aligner, Alignment, AlignmentOp, AlignmentRange, BatchResult, Config, SimpleAlignmentOp, StagedBatch,
                  TemplateSwitchEntranceOp, TemplateSwitchExitOp, TsaError, align, cigar_of)

__all__ = ["Aligner", "Alignment", "align", "AlignmentRange", "AlignmentOp", "SimpleAlignmentOp", "TemplateSwitchEntranceOp",
           "TemplateSwitchExitOp", "BatchResult", "Config", "StagedBatch", "TsaError", "cigar_of"]
__version__ = "0.1.0"
