"""ORC stripe decode on the device behind the same format seam as Parquet: the reference's
golden ORC files (OrcReaderFactoryTest.java:116-117, 264-265), pyarrow.orc as the byte-level oracle for every type a
KeyValue schema uses, and the fused path ORC files -> device runs -> merge against the merge oracle."""
import os
import random

import numpy as np
import pyarrow as pa
import pyarrow.orc as orc
import pytest

from oracle import pyoracle
from paimon_b200 import _native as N
from paimon_b200 import datagen
from paimon_b200.columnar import KeyValueBatch
from paimon_b200.format import FileFormat, FormatReaderContext, LocalFileIO, read_section
from paimon_b200.merge_function import DeduplicateMergeFunction, PartialUpdateMergeFunction
from paimon_b200.types import DataField, KeyValueSchema, RowType

from parquet_util import arrow_to_batch, to_arrow

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "orc")


def _close_all(readers):
    out = []
    for r in readers:
        try:
            out.append(r.read_batch())
        finally:
            r.close()
    return out


def test_golden_file_data_through_the_device_decoder():
    """The rows of the reference's golden ORC files (read directly by the host build of the same decoders in
    tests/test_orc_cpu.py; they are plain tables without _SEQUENCE_NUMBER / _VALUE_KIND columns) re-encoded as KeyValue
    ORC files and decoded on the device: every column against pyarrow.orc, plus the reference's own checksums."""
    import tempfile
    t = orc.ORCFile(os.path.join(GOLDEN, "test-data-flat.orc")).read()
    n = 300_000
    fields = [("_col0", "INT"), ("_col1", "STRING"), ("_col2", "STRING"), ("_col3", "STRING"), ("_col4", "INT"),
              ("_col5", "STRING"), ("_col6", "INT"), ("_col7", "INT"), ("_col8", "INT")]
    vt = RowType(tuple(DataField(nm, ty, i > 0) for i, (nm, ty) in enumerate(fields)))
    schema = KeyValueSchema.of(vt, ["_col0"])
    kv = pa.table({"_KEY__col0": t.column(0).slice(0, n), "_SEQUENCE_NUMBER": pa.array(np.arange(n, dtype=np.int64)),
                   "_VALUE_KIND": pa.array(np.zeros(n, np.int8)), **{nm: t.column(i).slice(0, n) for i, (nm, _) in enumerate(fields)}})
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "flat.orc")
        orc.write_table(kv, p, compression="zlib", file_version="0.11", stripe_size=256 * 1024)      # like the golden file
        readers, info = read_section(schema, [(open(p, "rb").read(), 0)], 1, file_format="orc")
        got = _close_all(readers)[0]
        assert info.n_rows == n and info.n_chunks >= schema.n_cols
        want = arrow_to_batch(schema, orc.ORCFile(p).read())
        assert got.equals(want), got.first_difference(want)
        assert int(np.asarray(got.columns[3].data, np.int64).sum()) == n * (n + 1) // 2     # _col0 = 1, 2, 3, ... in the golden file
        # DECIMAL(10,5) with 2,000 NULLs (OrcReaderFactoryTest.java:264-265) as its unscaled long
        d = orc.ORCFile(os.path.join(GOLDEN, "test-data-decimal.orc")).read().column(0)
        m = d.length()
        vt2 = RowType((DataField("pk", "BIGINT", False), DataField("dec", "DECIMAL(10,5)", True)))
        s2 = KeyValueSchema.of(vt2, ["pk"])
        kv2 = pa.table({"_KEY_pk": pa.array(np.arange(m, dtype=np.int64)), "_SEQUENCE_NUMBER": pa.array(np.arange(m, dtype=np.int64)),
                        "_VALUE_KIND": pa.array(np.zeros(m, np.int8)), "pk": pa.array(np.arange(m, dtype=np.int64)), "dec": d})
        p2 = os.path.join(tmp, "dec.orc")
        orc.write_table(kv2, p2, compression="zstd")
        readers, _ = read_section(s2, [(open(p2, "rb").read(), 0)], 1, file_format="orc")
        vals = _close_all(readers)[0].columns[4].to_pylist()
        ref = d.to_pylist()
        assert m == 6000 and sum(1 for v in vals if v is None) == 2000
        assert all((r is None and v is None) or (r is not None and int(r.scaleb(5)) == v) for r, v in zip(ref, vals))


def all_types_schema():
    vt = RowType((DataField("pk", "BIGINT", False), DataField("t", "TINYINT", True), DataField("s", "SMALLINT", True),
                  DataField("i", "INT", True), DataField("l", "BIGINT", True), DataField("f", "FLOAT", True),
                  DataField("d", "DOUBLE", True), DataField("b", "BOOLEAN", True), DataField("low", "STRING", True),
                  DataField("high", "STRING", True), DataField("bin", "BINARY", True), DataField("dt", "DATE", True)))
    return KeyValueSchema.of(vt, ["pk"])


def random_batch(schema, n, seed, null_p, key0=0):
    rng = random.Random(seed)

    def opt(v):
        return None if rng.random() < null_p else v
    words = ["alpha", "beta", "gamma", "delta", "paimon", "", "lsm-tree", "x" * 40]
    rows = []
    for k in range(n):
        key = key0 + 3 * k
        rows.append((key, rng.randrange(1 << 40), rng.choice([0, 0, 0, 3]), key, opt(rng.randrange(-128, 128)),
                     opt(rng.randrange(-32768, 32768)), opt(rng.choice([7, 7, rng.randrange(-2 ** 31, 2 ** 31)])),
                     opt(rng.choice([k * 1000, rng.randrange(-2 ** 62, 2 ** 62)])), opt(np.float32(rng.uniform(-1e3, 1e3)).item()),
                     opt(rng.uniform(-1e9, 1e9)), opt(rng.random() < 0.5), opt(rng.choice(words)),
                     opt("u%08d-%s" % (rng.randrange(10 ** 8), "z" * rng.randrange(0, 9))),
                     opt(bytes(rng.randrange(256) for _ in range(rng.randrange(0, 12)))), opt(rng.randrange(-10000, 30000))))
    return KeyValueBatch.from_rows(schema, rows)


def write_kv_orc(batch, path, **opts):
    t = to_arrow(batch)
    # DATE travels as int32 in the test batches: give the ORC file a real date column
    names = t.column_names
    if "dt" in names:
        t = t.set_column(names.index("dt"), "dt", t.column("dt").cast(pa.date32()))
    orc.write_table(t, path, **opts)


@pytest.mark.parametrize("opts", [
    dict(compression="uncompressed"),
    dict(compression="zlib"),
    dict(compression="zstd"),
    dict(compression="zlib", file_version="0.11"),
    dict(compression="zlib", stripe_size=64 * 1024, compression_block_size=65536),
    dict(compression="zstd", dictionary_key_size_threshold=1.0),
])
def test_all_types_against_pyarrow(tmp_path, opts):
    schema = all_types_schema()
    for n, null_p in ((0, 0.0), (1, 0.0), (33, 0.3), (5000, 0.25), (30000, 0.0), (12000, 0.9)):
        batch = random_batch(schema, n, seed=n + 3, null_p=null_p)
        path = str(tmp_path / f"a{n}.orc")
        write_kv_orc(batch, path, **opts)
        fmt = FileFormat.from_identifier("orc")
        rd = fmt.create_reader_factory(schema).create_reader(FormatReaderContext(LocalFileIO(), path))
        try:
            got = rd.read_batch()
        finally:
            rd.close()
        if n == 0:
            assert got is None
        else:
            assert got.equals(batch), got.first_difference(batch)


@pytest.mark.parametrize("engine", ["dedup", "partial-update"])
def test_merge_on_read_over_orc_files(tmp_path, engine):
    """MergeFileSplitRead over ORC data files: sections, runs made of several files, device decode + merge."""
    from paimon_b200.merge_tree_readers import DataFileMeta, MergeFileSplitRead, concat_batches
    schema = datagen.schema_c3(n_i64=3, n_f64=2, n_str=2)
    rng = np.random.default_rng(21)
    metas, file_runs = [], []
    for f in range(5):
        keys = np.sort(rng.choice(40000, size=9000, replace=False)).astype(np.int64)
        file_runs.append(datagen.make_run(schema, f, keys, seed=5, null_prob=0.4, delete_prob=0.1))
    for j in range(4):                                     # one run of four key-disjoint files
        keys = np.arange(j * 10000, j * 10000 + 9000, 2, dtype=np.int64)
        file_runs.append(datagen.make_run(schema, 9, keys, seed=5, null_prob=0.4))
    for i, run in enumerate(file_runs):
        path = str(tmp_path / f"data-{i}.orc")
        orc.write_table(to_arrow(run), path, compression=["zlib", "zstd", "uncompressed"][i % 3], stripe_size=128 * 1024)
        k = run.columns[0].data
        metas.append(DataFileMeta(path, 0, run.n_rows, int(k[0]), int(k[-1])))
    factory = (DeduplicateMergeFunction.factory() if engine == "dedup"
               else PartialUpdateMergeFunction.factory({"ignore-delete": "true"}, schema.value_type, ["pk"]))
    rd = MergeFileSplitRead(schema, factory).create_merge_reader(metas)
    batches = []
    while True:
        b = rd.read_batch()
        if b is None:
            break
        batches.append(b)
    rd.close()
    got = concat_batches(schema, batches)
    want = pyoracle.merge(schema, factory.create().with_drop_delete(True), file_runs)
    assert got.equals(want), got.first_difference(want)


def test_unsupported_orc_files_are_refused(tmp_path):
    schema = all_types_schema()
    batch = random_batch(schema, 100, seed=1, null_p=0.1)
    p = str(tmp_path / "snappy.orc")
    write_kv_orc(batch, p, compression="snappy")
    with pytest.raises(N.UnsupportedOnDevice, match="compression"):
        read_section(schema, [(open(p, "rb").read(), 0)], 1, file_format="orc")
    with pytest.raises(N.PaimonGpuError):
        read_section(schema, [(b"ORC-not-really" * 10, 0)], 1, file_format="orc")
