"""Structural rules of the repository, checked on the sources (no GPU): the product package must not reach into the test
infrastructure (oracle/, baseline/), must not carry a CPU fall-back for the CUDA library, and the C-ABI sources must not use the
batched-memcpy driver entry points the GPU pool refuses."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "egom2p_b200")


def _sources(exts):
    for d, _, files in os.walk(PKG):
        if "build" in os.path.relpath(d, PKG).split(os.sep):
            continue
        for f in files:
            if f.endswith(exts):
                yield os.path.join(d, f)


def test_product_package_never_imports_oracle_or_baseline():
    bad = re.compile(r"^\s*(from|import)\s+(oracle|baseline|egom2p_oracle|synth|masking_oracle|sampling_oracle|ref_gpu)\b", re.M)
    path_hack = re.compile(r"[\"'](oracle|baseline)[\"']")
    for p in _sources((".py",)):
        src = open(p).read()
        assert not bad.search(src), f"{p} imports test infrastructure"
        assert not path_hack.search(src), f"{p} builds a path into oracle/ or baseline/"


def test_library_loader_has_no_fallback():
    src = open(os.path.join(PKG, "_lib.py")).read()
    assert "raise RuntimeError" in src and "no fallback" in src   # a missing .so is an error, never a silent CPU path
    for p in _sources((".py",)):
        s = open(p).read()
        assert "except ImportError" not in s and "except OSError" not in s, f"{p}: swallowed load error could hide a fall-back"


def test_no_batched_memcpy_entry_points_in_sources():
    names = ["cudaMemcpy" + "BatchAsync", "cudaMemcpy3D" + "BatchAsync", "cuMemcpy" + "BatchAsync", "cuMemcpy3D" + "BatchAsync"]
    for p in list(_sources((".cu", ".cuh", ".py", ".h"))) + [os.path.join(ROOT, "include", "egom2p_b200.h"), os.path.join(ROOT, "bench.py")]:
        s = open(p).read()
        for n in names:
            assert n not in s, f"{p} names {n}"


def test_sm100a_only_build_flags():
    src = open(os.path.join(PKG, "build.py")).read()
    assert "arch=compute_100a,code=sm_100a" in src and "-lineinfo" in src
    assert "sm_90" not in src and "sm_80" not in src   # no multi-architecture dispatch
