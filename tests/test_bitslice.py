"""dwt_b200/csrc/bitslice.cuh on the CPU: the in-register 16 x 16 bit transpose that turns 32 Hilbert-ordered coefficients
into bit-plane words (linearize_tma_kernel) and back (reconstruct_tma_kernel).  The header is host-callable; a small C++
program checks every plane bit against the definition of the bit-sliced store (encode.c:112-131: sign-magnitude)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bitslice_transpose_matches_definition(tmp_path):
    exe = str(tmp_path / "bitslice_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "dwt_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host", "bitslice_check.cpp"), "-o", exe])
    r = subprocess.run([exe, "3000"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
