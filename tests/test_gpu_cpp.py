"""Runs the C++ port of the reference's library-level GPR tests (tests/cpp/gpr_tests.cpp: predict.rs:54-99 and
tests/gpr_tests.rs:72-225) — compiled code calling libhbegp.so through include/hbegp.hpp, no Python in between."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "gpr_tests")


@pytest.mark.gpu
def test_cpp_port_of_reference_gpr_tests():
    if not os.path.exists(BIN):
        subprocess.run(["make", "-C", os.path.join(ROOT, "hbetune_rs_b200", "csrc")], check=True)
    res = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    print(res.stdout[-4000:])
    print(res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-4000:]
    assert " 0 failures" in res.stdout


def test_cpp_header_compiles_without_a_gpu(tmp_path):
    """include/hbegp.hpp is self-contained C++17 over the C ABI."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "hbegp.hpp"\nint main() { hbegp_cpp::BoundedValue b(1.0, 0.5, 2.0); '
                   'return b.with_clamped_value(3.0).value() == 2.0 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", os.path.join(ROOT, "hbetune_rs_b200"), "-lhbegp",
                    "-Wl,-rpath," + os.path.join(ROOT, "hbetune_rs_b200")], check=True)
    assert subprocess.run([str(exe)]).returncode == 0
