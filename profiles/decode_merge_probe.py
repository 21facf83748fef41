"""Parquet files -> device decode -> k-way merge without leaving HBM: 16 files (one sorted run each) of the C3 row
shape, decoded by the device decoder and merged by the partial-update merge.  Prints per-phase device times.
Usage: decode_merge_probe.py [rows_per_file] [plain|dict|snappy]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from paimon_b200 import datagen
from paimon_b200.format import FileFormat, FormatReaderContext, LocalFileIO
from paimon_b200.merge_function import PartialUpdateMergeFunction
from paimon_b200.sort_merge_reader import SortMergeReader
from parquet_util import write_kv_parquet

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
mode = sys.argv[2] if len(sys.argv) > 2 else "plain"
schema = datagen.schema_c3()
n_runs = 16
runs = datagen.make_runs(schema, n_runs, rows * n_runs, seed=5, null_prob=0.5)
paths, file_bytes = [], 0
for i, run in enumerate(runs):
    p = f"/tmp/dm_{mode}_{i}.parquet"
    write_kv_parquet(run, p, row_group_size=1 << 20, use_dictionary=(mode == "dict"),
                     compression="snappy" if mode == "snappy" else "none")
    paths.append(p); file_bytes += os.path.getsize(p)
fmt = FileFormat.from_identifier("parquet")
spec = PartialUpdateMergeFunction.factory({}, schema.value_type, ["pk"]).create()
best = None
results = []
for it in range(4):
    t0 = time.perf_counter()
    frs = [fmt.create_reader_factory(schema).create_reader(FormatReaderContext(LocalFileIO(), p)) for p in paths]
    t1 = time.perf_counter()
    if it % 2 == 0:
        rrs = [fr.as_sorted_run_reader() for fr in frs]                   # decode on the device, one file after the other
    else:
        from concurrent.futures import ThreadPoolExecutor               # ... or 8 files at a time (own streams)
        with ThreadPoolExecutor(8) as ex:
            rrs = list(ex.map(lambda fr: fr.as_sorted_run_reader(), frs))
    t2 = time.perf_counter()
    dec_ms = sum(fr.info().ms_decode for fr in frs)
    mr = SortMergeReader.create_sort_merge_reader(rrs, None, None, spec)
    mr.execute()
    st = mr.stats()
    t3 = time.perf_counter()
    res = dict(concurrent=bool(it % 2), rows_in=st.rows_in, rows_out=st.rows_out, decode_gpu_ms=round(dec_ms, 2), merge_gpu_ms=round(st.ms_total, 2),
               open_parse_wall_ms=round((t1 - t0) * 1e3, 1), decode_wall_ms=round((t2 - t1) * 1e3, 1),
               merge_wall_ms=round((t3 - t2) * 1e3, 1))
    mr.close()
    for fr in frs:
        fr.close()
    results.append(res)
    if best is None or res["decode_gpu_ms"] + res["merge_gpu_ms"] < best["decode_gpu_ms"] + best["merge_gpu_ms"]:
        best = res
n_in = best["rows_in"]
dev_ms = best["decode_gpu_ms"] + best["merge_gpu_ms"]
for r in results[2:]:
    print("  ", r)
print(mode, "files", n_runs, "file MB", file_bytes >> 20, best, "decode+merge rows/s (device time)", round(n_in / (dev_ms * 1e-3)))
