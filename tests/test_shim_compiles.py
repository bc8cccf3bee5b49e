"""shim/spano_shim.cpp type-checks against the REFERENCE's own declarations.

The shim provides the bodies of the reference's callees with the reference's exact C++ signatures
(proj::*_proj::project, proj::get_proj_parameters, sten_proj::disk_reproj, blnd::*, dcut::*, gain::*, test::*).
OpenCV / Eigen headers are absent in the build container, so the shim is parsed (g++ -fsyntax-only -std=c++20) against the
reference's real headers under /root/reference/src plus minimal stand-ins for <opencv2/...> and <Eigen/...>
(tests/shim_stubs).  A changed signature on either side -- an argument added, a const dropped, a member renamed -- fails here.
Skipped where /root/reference does not exist (the GPU box).
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_shim_compiles_against_reference_headers():
    cmd = ["g++", "-std=c++20", "-fsyntax-only", "-Wall", "-Wno-unused", "-I", os.path.join(ROOT, "tests", "shim_stubs"),
           "-I", os.path.join(REF, "math"), "-I", os.path.join(REF, "system"), "-I", os.path.join(REF, "test"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "shim", "spano_shim.cpp")]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]
