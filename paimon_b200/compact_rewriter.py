"""The rewrite side of a compaction: merged device batch -> Parquet data files + DataFileMeta.

Mirrors (same names, same argument meaning):
  KeyValueDataFileWriter.write / result      paimon-core/.../io/KeyValueDataFileWriter.java:108-184
  RollingFileWriterImpl                      paimon-core/.../io/RollingFileWriterImpl.java:64-105
  MergeTreeCompactRewriter.rewriteCompaction paimon-core/.../mergetree/compact/MergeTreeCompactRewriter.java:78-116
  CompactResult(before, after)               paimon-core/.../compact/CompactResult.java

The merge of a section, the drop-delete filter, the Parquet encode and the file statistics all happen on the
device; the host writes the encoded bytes through the FileIO and assembles the DataFileMeta.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N
from .format import LocalFileIO
from .merge_function import MergeFunctionFactory
from .merge_tree_readers import (DataFileMeta, IntervalPartition, KeyValueFileReaderFactory, MergeTreeReaders,
                                 SortedRun)
from .sort_merge_reader import SortMergeReader
from .types import KeyValueSchema, PhysicalType, is_varlen


@dataclass
class SimpleColStats:
    """min / max / nullCount of one column (SimpleColStats in paimon-common/.../format/SimpleColStats.java)."""
    min: object
    max: object
    null_count: int


@dataclass
class WrittenFile:
    meta: DataFileMeta
    value_stats: List[SimpleColStats]
    ms_encode: float
    n_pages: int


def file_column_names(schema: KeyValueSchema) -> List[str]:
    """[_KEY_*, _SEQUENCE_NUMBER, _VALUE_KIND, value...] (KeyValue.schema, KeyValue.java:130-138)."""
    return [f.name for f in schema.file_fields()]


class KeyValueDataFileWriter:
    """Encodes rows [row0, row0 + n_rows) of a device batch (a merge handle holding a batch, or a run handle) as
    one Parquet data file and returns its DataFileMeta."""

    def __init__(self, schema: KeyValueSchema, path: str, level: int, file_io: Optional[LocalFileIO] = None,
                 row_group_rows: int = 0, page_rows: int = 0):
        self.schema = schema
        self.path = path
        self.level = level
        self.file_io = file_io or LocalFileIO()
        self.opts = N.PgParquetWriteOptions(row_group_rows, page_rows)
        self.lib = N.load()

    def write(self, source_handle: int, row0: int = 0, n_rows: int = -1) -> WrittenFile:
        names = file_column_names(self.schema)
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        fh = C.c_uint64(0)
        N.check(self.lib.pg_parquet_encode(source_handle, arr, row0, n_rows, C.byref(self.opts), C.byref(fh)))
        try:
            meta = N.PgFileMeta()
            N.check(self.lib.pg_parquet_file_meta(fh.value, C.byref(meta)))
            buf = np.empty(max(meta.file_bytes, 1), np.uint8)
            N.check(self.lib.pg_parquet_file_fetch(fh.value, buf.ctypes.data, meta.file_bytes))
            with open(self.path, "wb") as f:
                f.write(buf[: meta.file_bytes].tobytes())
            stats = [self._column_stats(fh.value, c) for c in range(self.schema.n_cols)]
        finally:
            self.lib.pg_parquet_file_free(fh.value)
        # min / max key = the key ROW of the first / last record of the file (the batch is sorted by key):
        # KeyValueDataFileWriter.java:116-118,166-167 keeps the first and the last key it saw.  Every key field,
        # var-len ones included — column statistics would truncate composite keys and have nothing for strings.
        n_file = int(meta.n_rows)
        min_key = max_key = None
        if n_file > 0:
            min_key = self._key_row(source_handle, row0)
            max_key = self._key_row(source_handle, row0 + n_file - 1)
        dfm = DataFileMeta(file_name=self.path, file_size=int(meta.file_bytes), row_count=n_file,
                           min_key=min_key, max_key=max_key,
                           min_sequence_number=int(meta.min_sequence_number),
                           max_sequence_number=int(meta.max_sequence_number), level=self.level,
                           delete_row_count=int(meta.delete_row_count))
        return WrittenFile(dfm, stats[self.schema.n_key + 2:], float(meta.ms_encode), int(meta.n_pages))

    def _key_row(self, source_handle: int, row: int):
        """The primary key of one row of the device batch: a scalar for single-field keys, else a tuple."""
        from .sort_merge_reader import fetch_slice
        one = fetch_slice(self.schema, source_handle, row, row + 1)
        vals = [one.columns[i].to_pylist()[0] for i in range(self.schema.n_key)]
        return vals[0] if len(vals) == 1 else tuple(vals)

    def _column_stats(self, fh: int, c: int) -> SimpleColStats:
        nulls, has = C.c_int64(0), C.c_int32(0)
        mn, mx = np.zeros(1, np.int64), np.zeros(1, np.int64)
        N.check(self.lib.pg_parquet_file_column_stats(fh, c, C.byref(nulls), C.byref(has), mn.ctypes.data,
                                                      mx.ctypes.data))
        if not has.value:
            return SimpleColStats(None, None, int(nulls.value))
        t = self.schema.physical_types()[c]
        if t in (PhysicalType.FLOAT, PhysicalType.DOUBLE):
            return SimpleColStats(float(mn.view(np.float64)[0]), float(mx.view(np.float64)[0]), int(nulls.value))
        if t == PhysicalType.BOOL:
            return SimpleColStats(bool(mn[0]), bool(mx[0]), int(nulls.value))
        return SimpleColStats(int(mn[0]), int(mx[0]), int(nulls.value))


class RollingFileWriter:
    """Cuts one batch into files of at most `target_file_rows` rows (the reference rolls on the byte size of the
    output stream, RollingFileWriterImpl.java:85-100; rows are what the device knows before encoding)."""

    def __init__(self, schema: KeyValueSchema, directory: str, level: int, target_file_rows: int,
                 file_io: Optional[LocalFileIO] = None, prefix: str = "data", **writer_args):
        self.schema, self.directory, self.level = schema, directory, level
        self.target = max(8, (int(target_file_rows) + 7) & ~7)       # files start at multiples of 8 rows
        self.file_io = file_io
        self.prefix = prefix
        self.writer_args = writer_args
        self.results: List[WrittenFile] = []

    def write(self, source_handle: int, n_rows: int) -> List[WrittenFile]:
        for i, r0 in enumerate(range(0, n_rows, self.target)):
            path = os.path.join(self.directory, f"{self.prefix}-{len(self.results)}.parquet")
            w = KeyValueDataFileWriter(self.schema, path, self.level, self.file_io, **self.writer_args)
            self.results.append(w.write(source_handle, r0, min(self.target, n_rows - r0)))
        return self.results


@dataclass
class CompactResult:
    before: List[DataFileMeta] = field(default_factory=list)
    after: List[DataFileMeta] = field(default_factory=list)
    written: List[WrittenFile] = field(default_factory=list)


class MergeTreeCompactRewriter:
    """rewriteCompaction(outputLevel, dropDelete, sections): every section is merged on the device, the merged
    batch never leaves HBM before it is encoded (MergeTreeCompactRewriter.java:78-116)."""

    def __init__(self, schema: KeyValueSchema, mf_factory: MergeFunctionFactory, directory: str,
                 user_defined_seq_comparator=None, file_io: Optional[LocalFileIO] = None, device: int = 0,
                 target_file_rows: int = 4 << 20, **writer_args):
        self.schema = schema
        self.mf_factory = mf_factory
        self.directory = directory
        self.udsc = user_defined_seq_comparator
        self.reader_factory = KeyValueFileReaderFactory(schema, file_io, device)
        self.file_io = file_io
        self.device = device
        self.target_file_rows = target_file_rows
        self.writer_args = writer_args

    def rewrite(self, output_level: int, drop_delete: bool, sections: Sequence[Sequence[SortedRun]]) -> CompactResult:
        return self.rewrite_compaction(output_level, drop_delete, sections)

    def rewrite_compaction(self, output_level: int, drop_delete: bool,
                           sections: Sequence[Sequence[SortedRun]]) -> CompactResult:
        result = CompactResult()
        spec = self.mf_factory.create().with_drop_delete(drop_delete)
        rolling = RollingFileWriter(self.schema, self.directory, output_level, self.target_file_rows, self.file_io,
                                    prefix=f"compact-l{output_level}", **self.writer_args)
        for section in sections:
            for run in section:
                result.before += run.files
            runs = MergeTreeReaders.open_runs(section, self.reader_factory)
            try:
                merge = SortMergeReader.create_sort_merge_reader(runs, None, self.udsc, spec, device=self.device)
            except Exception:
                for r in runs:
                    r.close()
                raise
            try:
                merge.execute()
                n_out = merge.device_batch().n_rows
                if n_out:
                    rolling.write(merge._merge_h, n_out)
            finally:
                merge.close()
        result.written = rolling.results
        result.after = [w.meta for w in rolling.results]
        return result

    @staticmethod
    def sections_of(files: Sequence[DataFileMeta]) -> List[List[SortedRun]]:
        return IntervalPartition(files).partition()
