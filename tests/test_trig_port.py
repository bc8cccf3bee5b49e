"""csrc/glibc_trig.cuh (the sinf / cosf / atanf / atan2f the CUDA warp kernels evaluate) against this machine's libm.

The header's operations are individually rounded IEEE operations on the host and on the device, so bit-for-bit
agreement with glibc here is what makes the device maps bit-identical to cv2's buildMaps (tests/test_gpu_parity.py
checks that end on the GPU).  The default run samples every 61st float (70 M arguments per function, a prime stride so that
every exponent and mantissa pattern class is hit) plus 20 M atan2f pairs; `SPANO_TRIG_EXHAUSTIVE=1` walks all 2^32 floats
(27 s on 8 cores; run when the header changes -- last exhaustive run: 0 mismatches, glibc 2.39).
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read().replace("\n", " ")
    except OSError:
        return False


@pytest.mark.skipif(not _has_fma(), reason="glibc picks its non-FMA sinf/cosf on this CPU; the port follows the FMA variant")
def test_trig_port_matches_libm(tmp_path):
    exe = str(tmp_path / "trig_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-pthread", os.path.join(ROOT, "tests", "trig_check.cpp"),
                    "-o", exe, "-lm"], check=True)
    stride = "1" if os.environ.get("SPANO_TRIG_EXHAUSTIVE") else "61"
    p = subprocess.run([exe, stride, "20000000"], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "mismatches 0 0 0 0" in p.stdout, p.stdout + p.stderr
