"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/gs_b200.h declares, its host-side
parameter derivation agrees with the oracle's restatement of the reference constructors (src/alfa_gs.cpp:16-63),
and it refuses to work without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from groundsegmentationrosnode_b200 import capi, params as P
from groundsegmentationrosnode_b200.csrc import build as csrc_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    csrc_build.build()
    return capi.load()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "gs_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|void|const char\*)\s+(gs_[a-z_0-9]+)\s*\(", header, flags=re.M))
    assert declared == set(capi.EXPORTS)
    for name in sorted(declared):
        assert getattr(lib, name) is not None
    assert lib.gs_abi_version() == 1


@pytest.mark.parametrize("preset", range(6))
def test_presets_and_derivation_match_oracle(lib, oracle, preset):
    p = capi.default_params(preset)
    q = P.preset(preset)
    for name, _ in P.GsParams._fields_:
        assert getattr(p, name) == getattr(q, name), name
    d, o = capi.derive(p), oracle.derive(q)
    for name, _ in P.GsDerived._fields_:
        a, b = getattr(d, name), getattr(o, name)
        if name == "dilate_off":
            assert list(a)[: d.n_dilate] == list(b)[: o.n_dilate]
        else:
            assert a == b, name


@pytest.mark.parametrize("ws", [3, 5, 7, 9, 11, 13, 15])
def test_dilation_taps_match_oracle(lib, oracle, ws):
    p = P.preset(P.GS_PRESET_HDL64E_2048).copy(window_size=ws)
    d, o = capi.derive(p), oracle.derive(p)
    assert d.n_dilate == o.n_dilate and list(d.dilate_off)[: d.n_dilate] == list(o.dilate_off)[: o.n_dilate]


def test_thresholds_sweep_match_oracle(lib, oracle):
    import math

    for deg in list(range(1, 31)) + [2.5, 7.25]:
        for field in ("ground_angle_threshold", "start_angle_threshold"):
            p = P.preset(P.GS_PRESET_OS1_128).copy(**{field: math.radians(deg)})
            d, o = capi.derive(p), oracle.derive(p)
            assert (d.k_ground, d.s_start) == (o.k_ground, o.s_start)


def test_invalid_parameters_rejected(lib):
    good = P.preset(P.GS_PRESET_HDL64E_2048)
    for bad in (good.copy(window_size=8),  # ks = 6: std::logic_error in the reference (alfa_gs.cpp:631-633)
                good.copy(n_beams=1), good.copy(rpm=0), good.copy(sample_freq=0), good.copy(max_range=0.1),
                good.copy(max_angle=-1.0), good.copy(sample_freq=400000),  # W > 32767: Label::WrapCols takes int16_t
                good.copy(smoothing=P.GS_SMOOTH_SG, window_size=13), good.copy(variant=7)):
        with pytest.raises(capi.GsError) as e:
            capi.derive(bad)
        assert e.value.code == P.GS_EINVAL
        h = C.c_void_p()
        assert lib.gs_create(C.byref(bad), 0, C.byref(h)) == P.GS_EINVAL


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.GsError) as e:
        capi.Context(P.preset(P.GS_PRESET_VLP16))
    assert e.value.code == P.GS_ECUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "groundsegmentationrosnode_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f in ("synth.py",), os.path.join(dirpath, f)
