"""Build recipe for the CPU oracle — TEST INFRASTRUCTURE ONLY.

  oracle/libgs_oracle.so      the restated oracle (oracle/gs_oracle.cpp)
  oracle/_ref/libgs_ref.so    the reference's own Labeler/label.cpp + lib/utils.h, compiled verbatim from
                              /root/reference against oracle/shim (only where /root/reference exists;
                              the GPU box uses the prebuilt file that travels with the snapshot)

Both are git-ignored build products.  Usage: python oracle/build.py [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
ORACLE_SO = os.path.join(HERE, "libgs_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libgs_ref.so")
# no FMA contraction: the reference's x86-64 build has none
CXXFLAGS = ["-O2", "-std=c++14", "-ffp-contract=off", "-fPIC", "-shared", "-pthread"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def build_oracle(force=False):
    src = [os.path.join(HERE, "gs_oracle.cpp"), os.path.join(HERE, "..", "include", "gs_b200.h")]
    if force or _stale(ORACLE_SO, src):
        subprocess.check_call(["g++", *CXXFLAGS, src[0], "-o", ORACLE_SO])
    return ORACLE_SO


def build_ref(force=False):
    """Compile the reference's own sources where they lie; returns None when /root/reference is absent
    and no prebuilt library travelled."""
    if not os.path.isdir(REFERENCE):
        return REF_SO if os.path.exists(REF_SO) else None
    src = [os.path.join(HERE, "ref_wrap.cpp"), os.path.join(REFERENCE, "Labeler", "label.cpp"),
           os.path.join(HERE, "shim", "cv_bridge", "cv_bridge.h")]
    if force or _stale(REF_SO, src):
        os.makedirs(os.path.dirname(REF_SO), exist_ok=True)
        subprocess.check_call(["g++", *CXXFLAGS, "-I", os.path.join(HERE, "shim"), "-I", REFERENCE,
                               "-I", os.path.join(REFERENCE, "lib"), src[0], src[1], "-o", REF_SO])
    return REF_SO


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_oracle(force))
    print(build_ref(force))
