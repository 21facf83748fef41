"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/paimon_gpu.h declares,
validates handles, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from paimon_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = N.load()
    decl = declared_symbols()
    assert len(decl) >= 15
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert set(decl) == set(N.exported_symbols())
    assert lib.pg_abi_version() == 2


def test_no_cuda_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = N.load()
    st = lib.pg_init(0)
    assert st == 3 and b"no CPU fallback" in lib.pg_last_error()
    with pytest.raises(N.PaimonGpuError):
        N.init(0)


def test_handle_validation_without_device():
    lib = N.load()
    assert lib.pg_schema_free(12345) == 1
    assert lib.pg_merge_execute(999) == 1
    assert b"unknown" in lib.pg_last_error()


def test_interval_partition_host_logic_matches_oracle():
    import random
    import numpy as np
    from oracle import pyoracle
    lib = N.load()
    rng = random.Random(5)
    for _ in range(200):
        n = rng.randrange(1, 40)
        mn = np.array([rng.randrange(0, 300) for _ in range(n)], np.int64)
        mx = mn + np.array([rng.randrange(0, 60) for _ in range(n)], np.int64)
        sec = np.zeros(n, np.int32)
        run = np.zeros(n, np.int32)
        ns = C.c_int32(0)
        assert lib.pg_interval_partition(n, mn.ctypes.data, mx.ctypes.data, sec.ctypes.data, run.ctypes.data,
                                         C.byref(ns)) == 0
        osec, orun, ons = pyoracle.interval_partition(mn.tolist(), mx.tolist())
        assert ns.value == ons
        assert sec.tolist() == osec.tolist()
        # run ids are labels: compare the partition into runs, not the labels
        def groups(s, r):
            g = {}
            for i, (a, b) in enumerate(zip(s.tolist(), r.tolist())):
                g.setdefault((a, b), []).append(i)
            return sorted(g.values())
        assert groups(sec, run) == groups(osec, orun)


def test_interval_partition_with_string_and_composite_key_bounds():
    """IntervalPartition over non-integer key bounds (strings compare as UTF-8 bytes, tuples field by field): the same
    sections / runs as over integer bounds with the same order."""
    import random
    from paimon_b200.merge_tree_readers import DataFileMeta, IntervalPartition
    rng = random.Random(9)
    for trial in range(20):
        n = rng.randrange(1, 30)
        bounds = []
        for _ in range(n):
            a, b = sorted((rng.randrange(100), rng.randrange(100)))
            bounds.append((a, b))
        def shape(files):
            return [[[f.file_name for f in run.files] for run in sec] for sec in IntervalPartition(files).partition()]
        ints = [DataFileMeta(f"f{i}", 0, 1, a, b) for i, (a, b) in enumerate(bounds)]
        strs = [DataFileMeta(f"f{i}", 0, 1, "k%03dé" % a, "k%03dé" % b) for i, (a, b) in enumerate(bounds)]
        tups = [DataFileMeta(f"f{i}", 0, 1, (a // 10, "x%d" % (a % 10)), (b // 10, "x%d" % (b % 10)))
                for i, (a, b) in enumerate(bounds)]
        assert shape(ints) == shape(strs) == shape(tups)


def test_jni_shim_binds_every_export_and_compiles():
    """jni/paimon_gpu_jni.cc calls every function include/paimon_gpu.h exports, and passes a syntax-only compile
    against the JNI specification's signatures (jni/stub/jni.h; the image has no JDK)."""
    import subprocess
    src = open(os.path.join(ROOT, "jni", "paimon_gpu_jni.cc")).read()
    code = re.sub(r"//[^\n]*", "", src)
    called = set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", code))
    missing = [n for n in declared_symbols() if n not in called]
    assert not missing, f"exports without a JNI binding: {missing}"
    assert len(re.findall(r"JNIEXPORT", src)) >= 40
    res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "jni", "stub"),
                          "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "jni", "paimon_gpu_jni.cc")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_arrow_c_data_structs_match_pyarrow():
    """The ArrowSchema / ArrowArray definitions in include/paimon_gpu.h have the layout pyarrow's C interface uses."""
    from pyarrow.cffi import ffi
    assert ffi.sizeof("struct ArrowSchema") == 72 and ffi.sizeof("struct ArrowArray") == 80
    hdr = open(os.path.join(ROOT, "include", "paimon_gpu.h")).read()
    body = hdr[hdr.index("struct ArrowArray {"):]
    body = body[:body.index("};")]
    order = re.findall(r"(\w+)\s*(?:\)\s*\([^)]*\))?;", body)
    assert [n for n in order if n in ("length", "null_count", "offset", "n_buffers", "n_children", "buffers", "children",
                                      "dictionary", "release", "private_data")] == \
        ["length", "null_count", "offset", "n_buffers", "n_children", "buffers", "children", "dictionary", "release",
         "private_data"]
