"""Decode one synthetic Parquet file (2 M rows x 17 columns) a few times; used under ncu for per-kernel times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from paimon_b200 import datagen
from paimon_b200.format import FileFormat, FormatReaderContext, LocalFileIO
from parquet_util import write_kv_parquet
schema = datagen.schema_c3(n_i64=6, n_f64=4, n_str=4)
run = datagen.make_runs(schema, 1, 4_000_000, seed=3, null_prob=0.5)[0]
mode = sys.argv[1] if len(sys.argv) > 1 else "dict"
path = "/tmp/probe_%s.parquet" % mode
write_kv_parquet(run, path, row_group_size=1_000_000, use_dictionary=(mode in ("dict", "snappy")),
                 compression="snappy" if mode.startswith("snappy") else "none")
rd = FileFormat.from_identifier("parquet").create_reader_factory(schema).create_reader(FormatReaderContext(LocalFileIO(), path))
for _ in range(3):
    t0 = time.perf_counter(); r = rd.as_sorted_run_reader(); dt = time.perf_counter() - t0
    info = rd.info(); r.close()
print(mode, "rows", info.n_rows, "file MB", os.path.getsize(path) >> 20, "pages", info.n_data_pages, "gpu decode ms", round(info.ms_decode, 2), "call ms", round(dt * 1e3, 1))
rd.close()
