"""Host-side checks of the device arithmetic that must reproduce the host's bit for bit (no GPU needed):
gs_math.cuh's atan2f clone against this box's libm (reference src/alfa_gs.cpp:368 calls atan2f and truncates
alfa * 100, so one ulp flips centi-radian bins), and the division-free decode of the u16 range encoding
(src/alfa_gs.cpp:359) against a real double division for all 65535 inputs."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_atan2f_clone_and_decode_division_are_bit_exact():
    exe = os.path.join(tempfile.mkdtemp(), "clone_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-x", "c++", os.path.join(ROOT, "tests", "clone_check.cpp"), "-o", exe, "-lm"])
    out = subprocess.run([exe, "3000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad=0 bad_div=0" in out.stdout
