"""Build recipe for the CPU oracle — TEST INFRASTRUCTURE ONLY.

  oracle/libgs_oracle.so      the restated oracle (oracle/gs_oracle.cpp)
  oracle/_ref/libgs_ref.so    the reference's OWN sources — Labeler/label.cpp, lib/utils.h, src/alfa_gs.cpp,
                              src/alfa_node.cpp, src/cloud2range_node.cpp, src/range2cloud_node.cpp — compiled
                              verbatim from /root/reference against the stand-in ROS / PCL / OpenCV headers of
                              oracle/shim (only where /root/reference exists; the GPU box uses the prebuilt
                              file that travels with the snapshot)

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
    # the reference's own translation units, compiled where they lie (src/cloud2range_node.cpp and src/range2cloud_node.cpp
    # carry a `main` each and are pulled in by their wrappers)
    ref_units = [os.path.join(REFERENCE, "Labeler", "label.cpp"), os.path.join(REFERENCE, "src", "alfa_gs.cpp"),
                 os.path.join(REFERENCE, "src", "alfa_node.cpp")]
    wrappers = [os.path.join(HERE, w) for w in ("ref_wrap.cpp", "ref_alfa_wrap.cpp", "ref_c2r_wrap.cpp", "ref_r2c_wrap.cpp")]
    deps = [os.path.join(HERE, "shim", "gsref_cv.h"), os.path.join(HERE, "shim", "gsref_ros.h"),
            os.path.join(REFERENCE, "src", "cloud2range_node.cpp"), os.path.join(REFERENCE, "src", "range2cloud_node.cpp")]
    if force or _stale(REF_SO, ref_units + wrappers + deps + [os.path.abspath(__file__)]):
        os.makedirs(os.path.dirname(REF_SO), exist_ok=True)
        # -w: the reference does not compile warning-free (e.g. AlfaNode::process_config returns nothing)
        subprocess.check_call(["g++", *CXXFLAGS, "-w", "-I", os.path.join(HERE, "shim"), "-I", REFERENCE,
                               "-I", os.path.join(REFERENCE, "lib"), *wrappers, *ref_units, "-o", REF_SO])
    return REF_SO


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_oracle(force))
    print(build_ref(force))
