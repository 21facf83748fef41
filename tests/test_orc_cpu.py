"""ORC decode, host build: orc_meta.cc + inflate / zstd + orc_device.cuh are the sources the
device path compiles; here they run serially on the host and are pinned against
  * the reference's golden ORC files (tests/golden/orc: the numbers OrcReaderFactoryTest.java:116-117, 264-265 asserts),
  * pyarrow.orc as the byte-level oracle for files written with every type the KeyValue schema uses, NULLs, RLE v1 and
    v2, DIRECT and DICTIONARY strings, ZLIB / ZSTD / uncompressed, many stripes."""
import decimal
import os
import random

import numpy as np
import pyarrow as pa
import pyarrow.orc as orc
import pytest

import orc_util

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "orc")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    return orc_util.build(str(tmp_path_factory.mktemp("orc")))


def test_golden_flat_file_of_the_reference(lib):
    blob = open(os.path.join(GOLDEN, "test-data-flat.orc"), "rb").read()
    n, cols = orc_util.decode(lib, blob, [4, 0, 0, 0, 4, 0, 4, 4, 4])
    assert n == 1920800
    assert int(cols[0][0].astype(np.int64).sum()) == 1844737280400          # OrcReaderFactoryTest.java:116-117
    t = orc.ORCFile(os.path.join(GOLDEN, "test-data-flat.orc")).read()
    for c in range(9):
        ref = t.column(c)
        if len(cols[c]) == 2:
            assert cols[c][1].all() and np.array_equal(ref.to_numpy(zero_copy_only=False), cols[c][0])
        else:
            data, offs, valid = cols[c]
            refs = ref.to_pylist()
            assert valid.all() and int(offs[n]) == sum(len(x) for x in refs)
            for i in range(0, n, 4999):
                assert data[offs[i]:offs[i + 1]].tobytes().decode() == refs[i]


def test_golden_decimal_file_of_the_reference(lib):
    blob = open(os.path.join(GOLDEN, "test-data-decimal.orc"), "rb").read()
    n, cols = orc_util.decode(lib, blob, [8])
    ref = orc.ORCFile(os.path.join(GOLDEN, "test-data-decimal.orc")).read().column(0).to_pylist()
    assert n == 6000 and sum(1 for x in ref if x is None) == 2000            # OrcReaderFactoryTest.java:264-265
    vals, valid = cols[0]
    for r, x, v in zip(ref, vals, valid):
        assert (r is None) == (not v)
        if r is not None:
            assert int(r.scaleb(5)) == int(x)                                # DECIMAL(10,5) as its unscaled long


def _table(n, seed, null_p):
    rng = random.Random(seed)
    g = np.random.default_rng(seed)

    def opt(v):
        return None if rng.random() < null_p else v
    words = ["alpha", "beta", "gamma", "delta", "paimon", "", "lsm-tree", "x" * 40]
    return pa.table({
        "k": pa.array(np.arange(n, dtype=np.int64) * 3),
        "seq": pa.array(g.integers(0, 1 << 40, n)),
        "kind": pa.array(g.integers(0, 4, n).astype(np.int8)),
        "t": pa.array([opt(rng.randrange(-128, 128)) for _ in range(n)], pa.int8()),
        "s": pa.array([opt(rng.randrange(-32768, 32768)) for _ in range(n)], pa.int16()),
        "i": pa.array([opt(rng.choice([7, 7, 7, rng.randrange(-2 ** 31, 2 ** 31)])) for _ in range(n)], pa.int32()),
        "l": pa.array([opt(rng.choice([i * 1000, rng.randrange(-2 ** 62, 2 ** 62)])) for i in range(n)], pa.int64()),
        "f": pa.array([opt(np.float32(rng.uniform(-1e3, 1e3)).item()) for _ in range(n)], pa.float32()),
        "d": pa.array([opt(rng.uniform(-1e9, 1e9)) for _ in range(n)], pa.float64()),
        "b": pa.array([opt(rng.random() < 0.5) for _ in range(n)], pa.bool_()),
        "low": pa.array([opt(rng.choice(words)) for _ in range(n)], pa.string()),               # dictionary-friendly
        "high": pa.array([opt("u%08d-%s" % (rng.randrange(10 ** 8), "z" * rng.randrange(0, 9))) for _ in range(n)], pa.string()),
        "bin": pa.array([opt(bytes(rng.randrange(256) for _ in range(rng.randrange(0, 12)))) for _ in range(n)], pa.binary()),
        "dt": pa.array([opt(rng.randrange(-10000, 30000)) for _ in range(n)], pa.int32()).cast(pa.date32()),
        "dec": pa.array([opt(decimal.Decimal(rng.randrange(-10 ** 12, 10 ** 12)).scaleb(-4)) for _ in range(n)], pa.decimal128(15, 4)),
    })


WIDTHS = [8, 8, 1, 1, 2, 4, 8, 4, 8, 1, 0, 0, 0, 4, 8]


def _check(lib, table, path, **opts):
    orc.write_table(table, path, **opts)
    ref = orc.ORCFile(path).read()
    n, cols = orc_util.decode(lib, open(path, "rb").read(), WIDTHS)
    assert n == table.num_rows
    for c, name in enumerate(table.column_names):
        col = ref.column(name).combine_chunks()
        want_valid = ~np.asarray(col.is_null())
        if len(cols[c]) == 2:
            vals, valid = cols[c]
            assert np.array_equal(valid, want_valid), name
            if name == "dec":
                want = [None if v is None else int(v.scaleb(4)) for v in col.to_pylist()]
            elif name == "dt":
                want = col.cast(pa.int32()).to_pylist()
            elif name == "b":
                want = [None if v is None else int(v) for v in col.to_pylist()]
            elif name in ("f", "d"):
                dt = np.float32 if name == "f" else np.float64
                got = vals.view(dt)
                ref_np = col.to_numpy(zero_copy_only=False).astype(dt)
                assert np.array_equal(got[want_valid].view(np.uint8), ref_np[want_valid].view(np.uint8)), name
                continue
            else:
                want = col.to_pylist()
            got = [int(x) if v else None for x, v in zip(vals, valid)]
            assert got == want, name
        else:
            data, offs, valid = cols[c]
            assert np.array_equal(valid, want_valid), name
            want = col.to_pylist()
            for i in range(n):
                b = data[offs[i]:offs[i + 1]].tobytes()
                w = want[i]
                assert (b == b"") if w is None else (b == (w.encode() if isinstance(w, str) else w)), (name, i)


@pytest.mark.parametrize("opts", [
    dict(compression="uncompressed"),
    dict(compression="zlib"),
    dict(compression="zstd"),
    dict(compression="zlib", file_version="0.11"),                       # RLE v1, DIRECT / DICTIONARY (v1) encodings
    dict(compression="zlib", stripe_size=64 * 1024, compression_block_size=65536),     # many stripes, many chunks
    dict(compression="zstd", dictionary_key_size_threshold=1.0),         # every string column dictionary-encoded
    dict(compression="uncompressed", dictionary_key_size_threshold=0.0, file_version="0.11"),
])
def test_all_types_against_pyarrow(lib, tmp_path, opts):
    for n, null_p in ((0, 0.0), (1, 0.0), (31, 0.3), (5000, 0.25), (40000, 0.0), (20000, 0.9)):
        _check(lib, _table(n, seed=n + 1, null_p=null_p), str(tmp_path / "t.orc"), **opts)


def test_rle_v2_patched_base_and_delta_runs(lib, tmp_path):
    """Integer sequences that make the ORC writer choose every RLE v2 sub-encoding: constant runs (SHORT_REPEAT / fixed
    DELTA), arithmetic progressions (DELTA), small values with rare huge outliers (PATCHED_BASE), random (DIRECT)."""
    g = np.random.default_rng(5)
    n = 30000
    outl = g.integers(0, 100, n).astype(np.int64)
    outl[g.integers(0, n, 60)] = g.integers(1 << 40, 1 << 50, 60)
    cols = {"const": np.full(n, 42, np.int64), "prog": np.arange(n, dtype=np.int64) * -7 + 3,
            "outliers": outl, "rand": g.integers(-2 ** 62, 2 ** 62, n), "mono": np.cumsum(g.integers(0, 5, n)).astype(np.int64),
            "neg_outliers": -outl}
    t = pa.table({k: pa.array(v) for k, v in cols.items()})
    path = str(tmp_path / "rle.orc")
    orc.write_table(t, path, compression="uncompressed")
    nrows, got = orc_util.decode(lib, open(path, "rb").read(), [8] * len(cols))
    assert nrows == n
    for c, k in enumerate(cols):
        assert np.array_equal(got[c][0], cols[k]), k
