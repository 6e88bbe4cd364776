"""CPU: the reduced-radix Fq / G1 code the MSM kernels run (typlonk_b200/csrc/fq30.cuh is plain
host+device C++) compiled for the host and checked against the 64-bit-limb host reference:
Montgomery product, lazy add/sub bounds, zero test, mixed/general addition, doubling, special cases."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fq30_host_check(tmp_path):
    exe = tmp_path / "fq30_check"
    src = os.path.join(ROOT, "tests", "native", "fq30_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), src], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    sys.stdout.write(res.stdout)
    assert res.returncode == 0, res.stdout
    assert "0 failures" in res.stdout
