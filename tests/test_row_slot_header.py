"""The Python mirrors of the twiddle-slot functions in tests/test_row_pass_model.py must be the functions the
kernels and the host actually use: compile a host-only program against aloha_b200/csrc/kernels.cuh and compare
every value.  (nvcc compiles it; nothing runs on a GPU.)"""
import os
import shutil
import subprocess

import pytest

from test_row_pass_model import row_slot, row_slot8

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include <cstdio>
#include "kernels.cuh"
int main() {
    for (unsigned u = 0; u < 8; ++u)
        for (unsigned j = 0; j < (1u << u); ++j) std::printf("%u %u %u %u\n", u, j, alb::row_slot(u, j), alb::row_slot8(u, j));
    std::printf("sizes %zu %zu %zu\n", sizeof(alb::NttJob), sizeof(alb::NttRowGroup), sizeof(alb::ModulusConsts));
    return 0;
}
'''


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc")
def test_python_mirrors_match_the_header(tmp_path):
    src, exe = tmp_path / "slots.cu", tmp_path / "slots"
    src.write_text(SRC)
    subprocess.run(["nvcc", "-std=c++17", "-O0", "-I", os.path.join(ROOT, "aloha_b200", "csrc"), "-o", str(exe), str(src)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    rows = [tuple(map(int, l.split())) for l in out if l and not l.startswith("sizes")]
    assert len(rows) == 255
    for u, j, a, b in rows:
        assert a == row_slot(u, j) and b == row_slot8(u, j), (u, j)
    sizes = [l for l in out if l.startswith("sizes")][0].split()[1:]
    job, group, consts = map(int, sizes)
    assert job % 16 == 0 and group % 16 == 0          # one record = one bulk copy (16-byte granularity)
    assert consts == 64
