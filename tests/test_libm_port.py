"""The device libm ports (csrc/lg_libm.cuh) are compiled for the HOST and compared bit-for-bit with the host libm —
the libm the reference's sin/cos/atan/atan2 calls resolve to.  No GPU needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_check(tmp):
    exe = os.path.join(tmp, "libm_port_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-I", os.path.join(ROOT, "gpscalibration_b200", "csrc"),
                           os.path.join(ROOT, "tests", "helpers", "libm_port_check.cpp"), "-o", exe, "-lm"])
    return exe


def run_check(tmp, millions):
    out = subprocess.check_output([build_check(tmp), str(millions)]).decode().split()
    return [int(v) for v in out]


def test_device_libm_port_is_bit_identical_to_host_libm(tmp_path):
    assert run_check(str(tmp_path), 8) == [0, 0, 0, 0]
