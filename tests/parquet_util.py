"""Write KeyValue batches as Paimon-shaped Parquet data files with pyarrow (the byte-level decode oracle,
SURVEY.md §8c): file schema [_KEY_*, _SEQUENCE_NUMBER BIGINT NOT NULL, _VALUE_KIND TINYINT NOT NULL, value...]."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

from paimon_b200.columnar import Column, KeyValueBatch, unpack_validity
from paimon_b200.types import PhysicalType, is_varlen

_PA = {PhysicalType.INT8: pa.int8(), PhysicalType.INT16: pa.int16(), PhysicalType.INT32: pa.int32(),
       PhysicalType.INT64: pa.int64(), PhysicalType.FLOAT: pa.float32(), PhysicalType.DOUBLE: pa.float64(),
       PhysicalType.STRING: pa.string(), PhysicalType.BINARY: pa.binary(), PhysicalType.BOOL: pa.bool_()}


def to_arrow(batch: KeyValueBatch) -> pa.Table:
    fields, arrays = [], []
    for f, col in zip(batch.schema.file_fields(), batch.columns):
        t = f.physical
        n = len(col)
        mask = None if col.valid is None else ~unpack_validity(col.valid, n)
        if is_varlen(t):
            vals = col.to_pylist()
            arr = pa.array(vals, type=_PA[t])
        else:
            data = np.asarray(col.data[:n])
            if t == PhysicalType.BOOL:
                data = data.astype(bool)
            arr = pa.array(data, type=_PA[t], mask=mask)
        fields.append(pa.field(f.name, _PA[t], nullable=f.nullable))
        arrays.append(arr)
    return pa.Table.from_arrays(arrays, schema=pa.schema(fields))


def write_kv_parquet(batch: KeyValueBatch, path: str, **kw) -> None:
    opts = dict(compression="none", use_dictionary=True, data_page_version="1.0", write_statistics=False)
    opts.update(kw)
    pq.write_table(to_arrow(batch), path, **opts)


def arrow_to_batch(schema, table: pa.Table) -> KeyValueBatch:
    cols = []
    for f, name in zip(schema.file_fields(), table.column_names):
        arr = table.column(name).combine_chunks()
        cols.append(Column.from_pylist(f.physical, arr.to_pylist()))
    return KeyValueBatch(schema, cols)
