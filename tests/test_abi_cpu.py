"""CPU-side checks of the product boundary: the C-ABI library loads, exports every symbol that
include/ie_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import interpolation_engine_b200 as ie
from interpolation_engine_b200 import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(ie.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return ie.load_library()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "ie_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ie_[a-z_]+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(ie.ABI), declared ^ set(ie.ABI)
    for name in declared:
        assert hasattr(lib, name), name


def test_rust_bindings_cover_the_header():
    """rust/ie-b200/src/sys.rs (sources only: no Rust toolchain in this image) declares every entry point of the header."""
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ie_b200.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(ie_[a-z_]+)\s*\(", header))
    bound = set(re.findall(r"pub fn (ie_[a-z_]+)\s*\(", open(os.path.join(ROOT, "rust", "ie-b200", "src", "sys.rs")).read()))
    assert declared == bound, declared ^ bound


def test_no_cpu_fallback(lib):
    if lib.ie_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ie.EngineError, match="no CUDA device"):
        ie.Engine(0)


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "interpolation_engine_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("the cpu oracle", ""), f


def test_workload_generators_are_deterministic_and_shaped():
    st = workloads.c4_state()
    assert st.n == 65536
    keys = ie.Arena(st.keys, st.key_offs)
    assert keys.get(0) == b"slot-0" and keys.get(16384) == b"idx-0" and keys.get(65535) == b"q-32767"
    a, b = workloads.c4_templates(2048), workloads.c4_templates(2048)
    assert np.array_equal(a.bytes, b.bytes) and np.array_equal(a.offs, b.offs)
    t0 = a.get(0)
    assert b"{q-{idx-{slot-" in t0 and t0.count(b"{") >= 4
    shard = workloads.c4_templates(1024, start=1024)
    assert shard.n == 1024 and not np.array_equal(shard.bytes[:64], a.bytes[:64])
    k = workloads.c5_keys(10, 100)
    assert k.n == 1000 and k.get(0) == b"persona-0/field-0" and k.get(999) == b"persona-9/field-99"
    sets = workloads.c5_pattern_sets()
    assert len(sets) == 64 and all(1 <= len(s) <= 9 for s in sets)
