"""Python transcription of the reference's merge test vocabulary.

Reference (paimon-core/src/test/java/org/apache/paimon/):
  utils/ReusingTestData.java:40-144            — (key, sequenceNumber, valueKind, value) records, the
                                                 "k, seq, +/-, v | ..." string DSL and random generators
  mergetree/compact/MergeFunctionTestUtils.java:35-150 — independent expected-result calculators
  utils/TestReusingRecordReader.java           — key row = (int key), value row = (int key, bigint value)
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import List, Optional, Sequence

from paimon_b200.columnar import KeyValueBatch
from paimon_b200.types import DataField, KeyValueSchema, RowKind, RowType

# TestReusingRecordReader: key (INT), value (INT NOT NULL key copy, BIGINT nullable value)
VALUE_TYPE = RowType((DataField("f0", "INT", False), DataField("f1", "BIGINT", True)))
SCHEMA = KeyValueSchema.of(VALUE_TYPE, ["f0"])


@dataclass(frozen=True)
class ReusingTestData:
    key: int
    sequence_number: int
    value_kind: RowKind
    value: Optional[int]

    def sort_key(self):
        return (self.key, self.sequence_number)


def parse(s: str) -> List[ReusingTestData]:
    """ReusingTestData.parse (ReusingTestData.java:90-106): '+' = INSERT, anything else = DELETE."""
    out = []
    for kv in s.split("|"):
        if not kv.strip():
            continue
        parts = kv.split(",")
        assert len(parts) == 4, f"Found invalid data string {kv}"
        out.append(ReusingTestData(int(parts[0]), int(parts[1]),
                                   RowKind.INSERT if parts[2].strip() == "+" else RowKind.DELETE,
                                   int(parts[3])))
    return out


def _get_value(rng: random.Random) -> int:
    while True:                                  # ReusingTestData.java:141-144: [-5,5) \ {0}
        v = rng.randrange(10) - 5
        if v != 0:
            return v


def generate_ordered_no_duplicated_keys(rng: random.Random, n: int, only_add: bool, used_seq: set) -> List[ReusingTestData]:
    """ReusingTestData.generateOrderedNoDuplicatedKeys (:123-139); sequence numbers unique across readers."""
    result = {}
    while len(result) < n:
        key = rng.randrange(n * 3)
        while True:
            seq = rng.randrange(2 ** 63 - 1)
            if seq not in used_seq:
                used_seq.add(seq)
                break
        kind = RowKind.INSERT if (rng.random() < 0.5 or only_add) else RowKind.DELETE
        result[key] = ReusingTestData(key, seq, kind, _get_value(rng))
    return [result[k] for k in sorted(result)]


def generate_random_readers(rng: random.Random, only_add: bool, max_readers: int = 20, max_rows: int = 100):
    """CombiningRecordReaderTestBase.generateRandomData (:70-80)."""
    used = set()
    return [generate_ordered_no_duplicated_keys(rng, rng.randrange(max_rows) + 1, only_add, used)
            for _ in range(rng.randrange(max_readers) + 1)]


def to_batch(data: Sequence[ReusingTestData]) -> KeyValueBatch:
    rows = [(d.key, d.sequence_number, int(d.value_kind), d.key, d.value) for d in data]
    return KeyValueBatch.from_rows(SCHEMA, rows)


def from_batch(batch: KeyValueBatch) -> List[ReusingTestData]:
    return [ReusingTestData(k, seq, RowKind(kind), v) for (k, seq, kind, k2, v) in batch.to_rows()
            if (k == k2 or (_ for _ in ()).throw(AssertionError("value.f0 != key")))]


def _groups(data):
    data = sorted(data, key=ReusingTestData.sort_key)
    seqs = [(d.key, d.sequence_number) for d in data]
    assert len(set(seqs)) == len(seqs), "Found two records with the same sequenceNumber. This is invalid."
    groups = {}
    for d in data:
        groups.setdefault(d.key, []).append(d)
    return data, groups


def expected_for_deduplicate(data):                      # MergeFunctionTestUtils.java:35-47
    data, _ = _groups(data)
    return [d for i, d in enumerate(data) if i + 1 >= len(data) or d.key != data[i + 1].key]


def expected_for_first_row(data):                        # :133-145
    data, _ = _groups(data)
    return [d for i, d in enumerate(data) if i == 0 or d.key != data[i - 1].key]


def expected_for_partial_update(data, add_only: bool):   # :49-85 (ignore-delete = !addOnly)
    _, groups = _groups(data)
    out = []
    for group in groups.values():
        if len(group) == 1:
            out.append(group[0])                          # ReducerMergeFunctionWrapper passthrough
        elif add_only:
            out.append(group[-1])
        elif not any(d.value_kind == RowKind.INSERT for d in group):
            first = group[0]
            out.append(ReusingTestData(first.key, 0, RowKind.DELETE, first.value))
        else:
            out.append([d for d in group if d.value_kind.is_add()][-1])
    return out


def expected_for_agg_sum(data, add_only: bool, remove_record_on_delete: bool):   # :87-131
    _, groups = _groups(data)
    out = []
    for group in groups.values():
        last = group[-1]
        if len(group) == 1:
            out.append(group[0])
        elif add_only or not remove_record_on_delete:
            total = sum(d.value if d.value_kind.is_add() else -d.value for d in group)
            out.append(ReusingTestData(last.key, last.sequence_number, RowKind.INSERT, total))
        elif not any(d.value_kind == RowKind.INSERT for d in group):
            out.append(ReusingTestData(last.key, last.sequence_number, RowKind.DELETE, last.value))
        else:
            kind, total = None, None
            for d in group:
                if d.value_kind == RowKind.INSERT:
                    kind = RowKind.INSERT
                    total = d.value if total is None else total + d.value
                else:
                    kind = RowKind.DELETE
                    total = d.value
            out.append(ReusingTestData(last.key, last.sequence_number, kind, total))
    return out


def spread_over_runs(data: Sequence[ReusingTestData]) -> List[List[ReusingTestData]]:
    """Turn one unsorted stream with duplicate keys (SortBufferWriteBufferTestBase style) into sorted
    runs with unique keys per run: the j-th occurrence of a key goes to run j."""
    runs: List[dict] = []
    for d in data:
        for r in runs:
            if d.key not in r:
                r[d.key] = d
                break
        else:
            runs.append({d.key: d})
    return [[r[k] for k in sorted(r)] for r in runs]
