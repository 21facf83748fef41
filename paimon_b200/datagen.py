"""Seeded synthetic sorted runs of the benchmark shapes C1–C3 (BASELINE.json configs).

Key space K = N/2; every run is a sorted sample without replacement of N/R keys from [0, K);
sequence number = run * 2^32 + row ordinal (unique per bucket, newer run wins); kind = INSERT;
values = splitmix64(key xor column-salt).  Host generation is numpy (tests, CPU baseline sample); the
benchmark generates the same shapes directly in HBM with torch ops (bench.py).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .columnar import Column, KeyValueBatch, pack_validity
from .types import DataField, KeyValueSchema, PhysicalType, RowType


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def schema_c1() -> KeyValueSchema:
    vt = RowType((DataField("pk", "BIGINT", False), DataField("val", "BIGINT", True)))
    return KeyValueSchema.of(vt, ["pk"])


def schema_c2() -> KeyValueSchema:
    fields = [DataField("pk", "BIGINT", False)] + [DataField(f"c{i}", "BIGINT", True) for i in range(10)]
    return KeyValueSchema.of(RowType(tuple(fields)), ["pk"])


def schema_c3(n_i64: int = 20, n_f64: int = 15, n_str: int = 14) -> KeyValueSchema:
    """50-column wide row: pk + 20 BIGINT + 15 DOUBLE + 14 VARCHAR(24)."""
    fields = [DataField("pk", "BIGINT", False)]
    fields += [DataField(f"i{i}", "BIGINT", True) for i in range(n_i64)]
    fields += [DataField(f"d{i}", "DOUBLE", True) for i in range(n_f64)]
    fields += [DataField(f"s{i}", "VARCHAR(24)", True) for i in range(n_str)]
    return KeyValueSchema.of(RowType(tuple(fields)), ["pk"])


def run_keys(rng: np.random.Generator, key_space: int, n: int) -> np.ndarray:
    n = min(n, key_space)
    if n * 4 >= key_space:
        keys = rng.permutation(key_space)[:n]
    else:
        keys = np.unique(rng.integers(0, key_space, size=int(n * 1.2) + 16))
        while len(keys) < n:
            keys = np.unique(np.concatenate([keys, rng.integers(0, key_space, size=n)]))
        keys = rng.permutation(keys)[:n]
    return np.sort(keys).astype(np.int64)


def make_run(schema: KeyValueSchema, run_index: int, keys: np.ndarray, seed: int, null_prob: float = 0.0,
             delete_prob: float = 0.0, str_len=(8, 24)) -> KeyValueBatch:
    rng = np.random.default_rng(seed * 1000003 + run_index)
    n = len(keys)
    cols: List[Column] = []
    def key_column(t):
        if t in (PhysicalType.STRING, PhysicalType.BINARY):
            # 16-character lower-case hex of the key (big endian: string order == integer order)
            hexes = np.array([b"%016x" % int(k) for k in keys], dtype="S16") if n else np.zeros(0, "S16")
            data = np.frombuffer(hexes.tobytes(), np.uint8).copy()
            return Column(t, data, (np.arange(n + 1, dtype=np.int64) * 16).astype(np.int32))
        return Column(t, keys.astype(np.int64))
    for f in schema.key_type.fields:
        cols.append(key_column(f.physical))
    cols.append(Column(PhysicalType.INT64, (np.int64(run_index) << np.int64(32)) + np.arange(n, dtype=np.int64)))
    kinds = np.zeros(n, np.int8)
    if delete_prob > 0:
        kinds[rng.random(n) < delete_prob] = 3
    cols.append(Column(PhysicalType.INT8, kinds))
    pk_names = {f.name[len("_KEY_"):] for f in schema.key_type.fields}
    for ci, f in enumerate(schema.value_type.fields):
        t = f.physical
        salt = np.uint64((run_index + 1) * 0x100 + ci)
        h = splitmix64(keys.astype(np.uint64) ^ (salt << np.uint64(40)))
        valid = None
        if f.name not in pk_names and f.nullable and null_prob > 0:
            valid = pack_validity(rng.random(n) >= null_prob)
        if f.name in pk_names:
            cols.append(key_column(t))
        elif t == PhysicalType.INT64:
            cols.append(Column(t, h.view(np.int64), None, valid))
        elif t == PhysicalType.DOUBLE:
            # finite doubles with full mantissas: exercises bit-exact fp folds
            d = (h >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53)) * 2000.0 - 1000.0
            cols.append(Column(t, d, None, valid))
        elif t == PhysicalType.INT32:
            cols.append(Column(t, (h & np.uint64(0x7fffffff)).astype(np.int32), None, valid))
        elif t in (PhysicalType.STRING, PhysicalType.BINARY):
            lens = (h % np.uint64(str_len[1] - str_len[0] + 1)).astype(np.int64) + str_len[0]
            offs = np.zeros(n + 1, np.int64)
            np.cumsum(lens, out=offs[1:])
            total = int(offs[-1])
            # bytes derived from the key so that equal inputs give equal files
            pos = np.arange(total, dtype=np.int64) - np.repeat(offs[:-1], lens)
            src = np.repeat(h, lens)
            data = (((src >> ((pos % 8) * 8).astype(np.uint64)) & np.uint64(0x3f)) + np.uint64(0x30)).astype(np.uint8)
            cols.append(Column(t, data, offs.astype(np.int32), valid))
        else:
            raise ValueError(f"datagen: unsupported type {t}")
    return KeyValueBatch(schema, cols)


def make_runs(schema: KeyValueSchema, n_runs: int, total_rows: int, seed: int = 1, null_prob: float = 0.0,
              delete_prob: float = 0.0, key_space: Optional[int] = None) -> List[KeyValueBatch]:
    key_space = key_space or max(total_rows // 2, 1)
    per_run = total_rows // n_runs
    out = []
    for r in range(n_runs):
        rng = np.random.default_rng(1000 + r + seed * 7919)
        keys = run_keys(rng, key_space, per_run)
        out.append(make_run(schema, r, keys, seed, null_prob, delete_prob))
    return out
