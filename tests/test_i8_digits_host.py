"""CPU: the digit algebra of the INT8-sliced trailing update (agpn_chol.cuh: i8_digits2, agpn_i8.cuh: i8_fold4) compiled
for the host and checked against its definition -- q = rint(x) = sum_s a_s 128^(7-s) with balanced digits -- and, end to
end, that the 36 slice pairs reproduce a rank-1024 update of doubles to a few 2^-53 of the row scales."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_digit_algebra_on_the_host(tmp_path):
    exe = str(tmp_path / "i8_digits_host")
    src = os.path.join(ROOT, "tests", "i8_digits_host.cu")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " bad 0" in out.stdout
