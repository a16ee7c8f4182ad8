"""The C++ host mirror of the reference API (include/gs_host.hpp: Alfa_GS, Cloud2RangeNode, Range2CloudNode, Label)
compiled against the C ABI and checked against the oracle by tests/host_api_test.cpp."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "groundsegmentationrosnode_b200")


def _compile():
    from oracle import build as ob

    ob.build_oracle()
    exe = os.path.join(tempfile.mkdtemp(), "host_api_test")
    subprocess.check_call(["g++", "-O1", "-std=c++14", os.path.join(ROOT, "tests", "host_api_test.cpp"), "-o", exe,
                           "-L" + PKG, "-lgs_b200", "-L" + os.path.join(ROOT, "oracle"), "-lgs_oracle",
                           "-Wl,-rpath," + PKG, "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return exe


def test_host_header_compiles_against_the_c_abi():
    from groundsegmentationrosnode_b200.csrc import build as cb

    cb.build()
    _compile()


@pytest.mark.gpu
def test_reference_api_surface_matches_oracle():
    out = subprocess.run([_compile()], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host api ok" in out.stdout
