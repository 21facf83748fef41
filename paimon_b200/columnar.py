"""Host-side columnar batches (Arrow buffer layout) handed to / returned by the C ABI.

The reference's in-memory batch is ``VectorizedColumnBatch`` over ``ColumnVector``s
(paimon-common/src/main/java/org/apache/paimon/data/columnar/VectorizedColumnBatch.java:40-62,
heap/HeapLongVector.java:30-85, heap/HeapBytesVector.java:38-140).  We keep the same
struct-of-arrays shape but with Arrow buffers (validity bitmap LSB-first, int32 offsets for
var-len) because that is what crosses the C ABI and what paimon-arrow imports on the Java side.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .types import KeyValueSchema, PhysicalType, is_varlen, numpy_dtype


def pack_validity(mask: np.ndarray) -> np.ndarray:
    """bool[n] -> Arrow bitmap (LSB first), padded to a multiple of 8 bytes."""
    bits = np.packbits(np.asarray(mask, dtype=np.uint8), bitorder="little")
    pad = (-len(bits)) % 8
    if pad:
        bits = np.concatenate([bits, np.zeros(pad, np.uint8)])
    return bits


def unpack_validity(bitmap: Optional[np.ndarray], n: int) -> np.ndarray:
    if bitmap is None:
        return np.ones(n, dtype=bool)
    return np.unpackbits(np.asarray(bitmap, np.uint8), bitorder="little")[:n].astype(bool)


@dataclass
class Column:
    type: PhysicalType
    data: np.ndarray                       # fixed width: values[n]; var-len: uint8 bytes
    offsets: Optional[np.ndarray] = None   # var-len: int32[n+1]
    valid: Optional[np.ndarray] = None     # Arrow bitmap or None (= no nulls)

    def __len__(self) -> int:
        return len(self.offsets) - 1 if self.offsets is not None else len(self.data)

    @staticmethod
    def from_pylist(t: PhysicalType, values: Sequence) -> "Column":
        t = PhysicalType(t)
        n = len(values)
        mask = np.array([v is not None for v in values], dtype=bool)
        valid = None if mask.all() else pack_validity(mask)
        if is_varlen(t):
            chunks, offs = [], np.zeros(n + 1, np.int32)
            for i, v in enumerate(values):
                b = b"" if v is None else (v.encode() if isinstance(v, str) else bytes(v))
                chunks.append(b)
                offs[i + 1] = offs[i] + len(b)
            data = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy()
            return Column(t, data, offs, valid)
        arr = np.array([0 if v is None else v for v in values], dtype=numpy_dtype(t))
        return Column(t, arr, None, valid)

    def to_pylist(self) -> list:
        n = len(self)
        mask = unpack_validity(self.valid, n)
        out = []
        if is_varlen(self.type):
            raw = self.data.tobytes()
            for i in range(n):
                if not mask[i]:
                    out.append(None)
                else:
                    b = raw[self.offsets[i]:self.offsets[i + 1]]
                    out.append(b.decode() if self.type == PhysicalType.STRING else b)
        else:
            for i in range(n):
                out.append(self.data[i].item() if mask[i] else None)
        return out

    def canonical(self) -> "Column":
        """Zero the data under null slots so two columns can be compared with array_equal."""
        n = len(self)
        mask = unpack_validity(self.valid, n)
        if is_varlen(self.type):
            offs = np.asarray(self.offsets[: n + 1], np.int64)
            lens = (offs[1:] - offs[:-1]) * mask                     # payload under a NULL slot does not count
            new_offs = np.zeros(n + 1, np.int64)
            np.cumsum(lens, out=new_offs[1:])
            if n and int(new_offs[-1]) != int(offs[-1]) - int(offs[0]) or (n and offs[0] != 0):
                idx = np.repeat(offs[:-1] - new_offs[:-1], lens) + np.arange(int(new_offs[-1]))
                data = np.asarray(self.data)[idx] if len(idx) else np.zeros(0, np.uint8)
            else:
                data = np.asarray(self.data[: int(offs[-1]) if n else 0])
            return Column(self.type, data, new_offs.astype(np.int32), pack_validity(mask))
        d = np.array(self.data[:n], copy=True)
        d[~mask] = 0
        return Column(self.type, d, None, pack_validity(mask))

    def equals(self, other: "Column") -> bool:
        a, b = self.canonical(), other.canonical()
        if len(a) != len(b) or a.type != b.type:
            return False
        if not np.array_equal(unpack_validity(a.valid, len(a)), unpack_validity(b.valid, len(b))):
            return False
        if is_varlen(a.type):
            return np.array_equal(a.offsets, b.offsets) and np.array_equal(a.data, b.data)
        # bit-exact, also for floats (NaN payloads included)
        return a.data.tobytes() == b.data.tobytes()


@dataclass
class KeyValueBatch:
    """One sorted run (or one merged output batch) in file-column order."""

    schema: KeyValueSchema
    columns: List[Column]

    @property
    def n_rows(self) -> int:
        return len(self.columns[0]) if self.columns else 0

    @property
    def sequence_numbers(self) -> np.ndarray:
        return self.columns[self.schema.n_key].data[: self.n_rows]

    @property
    def value_kinds(self) -> np.ndarray:
        return self.columns[self.schema.n_key + 1].data[: self.n_rows]

    def key_column(self, i: int) -> Column:
        return self.columns[i]

    def value_column(self, i: int) -> Column:
        return self.columns[self.schema.n_key + 2 + i]

    def to_rows(self) -> list:
        cols = [c.to_pylist() if c is not None else [None] * self.n_rows for c in self.columns]
        return [tuple(c[i] for c in cols) for i in range(self.n_rows)]

    def project(self, keep: Sequence[bool]) -> "KeyValueBatch":
        """The batch with the value columns of a read-type projection only (others become None, the way a projected
        merge returns them); `keep` has one entry per VALUE field."""
        nk = self.schema.n_key + 2
        cols = [c if (i < nk or keep[i - nk]) else None for i, c in enumerate(self.columns)]
        return KeyValueBatch(self.schema, cols)

    def equals(self, other: "KeyValueBatch") -> bool:
        if self.n_rows != other.n_rows or len(self.columns) != len(other.columns):
            return False
        for a, b in zip(self.columns, other.columns):
            if (a is None) != (b is None):
                return False
            if a is not None and not a.equals(b):
                return False
        return True

    def first_difference(self, other: "KeyValueBatch") -> str:
        if self.n_rows != other.n_rows:
            return f"row count {self.n_rows} != {other.n_rows}"
        for ci, (a, b) in enumerate(zip(self.columns, other.columns)):
            if (a is None) != (b is None):
                return f"column {ci}: present in one batch only"
            if a is None:
                continue
            if not a.equals(b):
                la, lb = a.to_pylist(), b.to_pylist()
                for i, (x, y) in enumerate(zip(la, lb)):
                    if x != y and not (x != x and y != y):
                        return f"column {ci} row {i}: {x!r} != {y!r}"
                return f"column {ci}: buffers differ (bit-level)"
        return "equal"

    @staticmethod
    def from_rows(schema: KeyValueSchema, rows: Sequence[Sequence]) -> "KeyValueBatch":
        types = schema.physical_types()
        cols = [Column.from_pylist(t, [r[i] for r in rows]) for i, t in enumerate(types)]
        return KeyValueBatch(schema, cols)
