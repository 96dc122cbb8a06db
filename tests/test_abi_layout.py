"""The ctypes mirror of include/tsalign_b200.h (template_switch_aligner_b200/_lib.py) must have the layout the C compiler
gives the header's structs: sizes and field offsets, checked against a tiny C program built with gcc."""
import ctypes as C
import os
import subprocess

from template_switch_aligner_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = r"""
#include <stdio.h>
#include <stddef.h>
#include "tsalign_b200.h"
#define F(s, f) printf(#s "." #f " %zu\n", offsetof(s, f))
int main(void) {
    printf("tsa_options %zu\ntsa_pair %zu\ntsa_op %zu\ntsa_result %zu\n", sizeof(tsa_options), sizeof(tsa_pair), sizeof(tsa_op), sizeof(tsa_result));
    F(tsa_options, cost_limit); F(tsa_options, memory_limit); F(tsa_options, no_traceback); F(tsa_options, postprocess);
    F(tsa_pair, query); F(tsa_pair, reference_offset); F(tsa_pair, query_limit);
    F(tsa_op, type); F(tsa_op, value); F(tsa_op, min_start); F(tsa_op, max_end);
    F(tsa_result, cost); F(tsa_result, ops); F(tsa_result, n_ops); F(tsa_result, duration_seconds); F(tsa_result, message);
    F(tsa_result, reference_offset); F(tsa_result, query_limit);
    return 0;
}
"""


def test_struct_layout(tmp_path):
    src = tmp_path / "probe.c"
    src.write_text(PROBE)
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    lines = dict(ln.rsplit(" ", 1) for ln in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    mirror = {"tsa_options": _lib.TsaOptions, "tsa_pair": _lib.TsaPair, "tsa_op": _lib.TsaOp, "tsa_result": _lib.TsaResult}
    for key, value in lines.items():
        if "." in key:
            struct, field = key.split(".")
            assert getattr(mirror[struct], field).offset == int(value), key
        else:
            assert C.sizeof(mirror[key]) == int(value), key
