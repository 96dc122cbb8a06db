"""Loads the SIMT-emulator build of the product sources (tests/emul) -- CPU tests of the kernel logic only."""
import ctypes
import os
import subprocess

from template_switch_aligner_b200 import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_EMUL = None


def emul():
    global _EMUL
    if _EMUL is None:
        subprocess.run(["make", "-C", os.path.join(_HERE, "emul"), "-s"], check=True)
        _EMUL = _lib.bind(ctypes.CDLL(os.path.join(_HERE, "emul", "_build", "libtsalign_b200_emul.so")))
    return _EMUL
