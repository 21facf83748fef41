"""Pins the CPU oracle (oracle/paimon_oracle.c) to the reference's own fixed vectors and
expected-result calculators.  Everything here runs on CPU.

Transcribed from paimon-core/src/test/java/org/apache/paimon/mergetree/compact/:
  SortMergeReaderTestBase.java:51-89, CombiningRecordReaderTestBase.java:54-112,
  LoserTreeTest.java:51-68, MergeFunctionTestUtils.java:35-150,
  PartialUpdateMergeFunctionTest.java, aggregate/AggregateMergeFunctionTest.java:42-70,
  aggregate/FieldAggregatorTest.java:94-543, IntervalPartitionTest.java
and ../SortBufferWriteBufferTestBase.java:160-235 (partial-update / agg-sum option sets).
"""
import random

import numpy as np
import pytest

from oracle import pyoracle
from paimon_b200.columnar import KeyValueBatch
from paimon_b200.merge_function import (AggregateMergeFunction, DeduplicateMergeFunction,
                                        FirstRowMergeFunction, PartialUpdateMergeFunction)
from paimon_b200.types import DataField, KeyValueSchema, RowKind, RowType

from reusing_test_data import (SCHEMA, VALUE_TYPE, expected_for_agg_sum, expected_for_deduplicate,
                               expected_for_first_row, expected_for_partial_update, from_batch,
                               generate_random_readers, parse, spread_over_runs, to_batch)

ENGINES = [pyoracle.SORT_LOSER_TREE, pyoracle.SORT_MIN_HEAP, pyoracle.SORT_BRUTE_FORCE]

FIXED_VECTORS = {
    "empty1": [""],
    "empty3": ["", "", ""],
    "alternate_keys": [                               # SortMergeReaderTestBase.java:58-69
        "1, 1, +, 100 | 3, 2, +, 300 | 5, 3, +, 200 | 7, 4, +, 600 | 9, 20, +, 400",
        "0, 5, +, 0", "0, 10, +, 0", "",
        "2, 6, +, 200 | 4, 7, +, 400 | 6, 8, +, 600 | 8, 9, +, 800"],
    "duplicate_keys": ["1, 1, +, 100 | 3, 3, +, 300", "1, 4, +, 200 | 3, 5, +, 300"],   # :72-77
    "long_tail": [                                    # :80-89
        "1, 1, +, 100 | 2, 500, +, 200",
        "1, 3, +, 100 | 3, 4, +, 300 | 5, 501, +, 500 | 7, 503, +, 700 | "
        "8, 504, +, 800 | 9, 505, +, 900 | 10, 506, +, 1000 | "
        "11, 507, +, 1100 | 12, 508, +, 1200 | 13, 509, +, 1300"],
}


def run_merge(readers, spec, engine):
    out = pyoracle.merge(SCHEMA, spec, [to_batch(r) for r in readers], engine)
    return from_batch(out)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", sorted(FIXED_VECTORS))
def test_fixed_vectors_deduplicate(name, engine):
    readers = [parse(s) for s in FIXED_VECTORS[name]]
    spec = DeduplicateMergeFunction.factory().create()
    flat = [d for r in readers for d in r]
    assert run_merge(readers, spec, engine) == expected_for_deduplicate(flat)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", sorted(FIXED_VECTORS))
def test_fixed_vectors_first_row(name, engine):
    readers = [parse(s) for s in FIXED_VECTORS[name]]
    spec = FirstRowMergeFunction.factory().create()
    flat = [d for r in readers for d in r]
    assert run_merge(readers, spec, engine) == expected_for_first_row(flat)


@pytest.mark.parametrize("engine", ENGINES)
def test_random_deduplicate(engine):                  # CombiningRecordReaderTestBase.testRandom
    rng = random.Random(1234 + engine)
    spec = DeduplicateMergeFunction.factory().create()
    for _ in range(100):
        readers = generate_random_readers(rng, only_add=False)
        flat = [d for r in readers for d in r]
        assert run_merge(readers, spec, engine) == expected_for_deduplicate(flat)


@pytest.mark.parametrize("engine", ENGINES)
def test_random_first_row(engine):
    rng = random.Random(99 + engine)
    spec = FirstRowMergeFunction.factory().create()
    for _ in range(100):
        readers = generate_random_readers(rng, only_add=True)
        flat = [d for r in readers for d in r]
        assert run_merge(readers, spec, engine) == expected_for_first_row(flat)


@pytest.mark.parametrize("engine", [pyoracle.SORT_LOSER_TREE, pyoracle.SORT_MIN_HEAP])
def test_loser_tree_is_ordered(engine):               # LoserTreeTest.java:51-68
    rng = random.Random(7)
    spec = DeduplicateMergeFunction.factory().create()
    for _ in range(100):
        readers = generate_random_readers(rng, only_add=False)
        runs = [to_batch(r) for r in readers]
        out_run, out_row = pyoracle.merge_order(SCHEMA, spec, runs, engine)
        got = [readers[r][i] for r, i in zip(out_run.tolist(), out_row.tolist())]
        want = sorted((d for r in readers for d in r), key=lambda d: (d.key, d.sequence_number))
        assert got == want


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("add_only", [True, False])
def test_random_partial_update(add_only, engine):     # SortBufferWriteBufferTestBase.java:165-197
    rng = random.Random(4321 + engine)
    opts = {"ignore-delete": str(not add_only).lower()}
    spec = PartialUpdateMergeFunction.factory(opts, VALUE_TYPE, ["f0"]).create()
    for _ in range(60):
        readers = generate_random_readers(rng, only_add=add_only)
        flat = [d for r in readers for d in r]
        assert run_merge(readers, spec, engine) == expected_for_partial_update(flat, add_only)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("add_only,remove_on_delete", [(True, False), (False, False), (False, True)])
def test_random_agg_sum(add_only, remove_on_delete, engine):   # :199-235
    rng = random.Random(555 + engine)
    opts = {"fields.f1.aggregate-function": "sum",
            "aggregation.remove-record-on-delete": str(remove_on_delete).lower()}
    spec = AggregateMergeFunction.factory(opts, VALUE_TYPE, ["f0"]).create()
    for _ in range(60):
        readers = generate_random_readers(rng, only_add=add_only)
        flat = [d for r in readers for d in r]
        assert run_merge(readers, spec, engine) == expected_for_agg_sum(flat, add_only, remove_on_delete)


def test_write_buffer_style_streams():
    """SortBufferWriteBufferTestBase.testRandom: one stream with duplicated keys (many versions per key)."""
    from reusing_test_data import ReusingTestData, _get_value
    rng = random.Random(31337)
    for n in (100, 200):
        data, used = [], set()
        for _ in range(n):
            seq = rng.randrange(2 ** 63 - 1)
            while seq in used:
                seq = rng.randrange(2 ** 63 - 1)
            used.add(seq)
            kind = RowKind.INSERT if rng.random() < 0.5 else RowKind.DELETE
            data.append(ReusingTestData(rng.randrange(n), seq, kind, _get_value(rng)))
        readers = spread_over_runs(data)
        spec = DeduplicateMergeFunction.factory().create()
        assert run_merge(readers, spec, pyoracle.SORT_LOSER_TREE) == expected_for_deduplicate(data)
        spec = PartialUpdateMergeFunction.factory({"ignore-delete": "true"}, VALUE_TYPE, ["f0"]).create()
        assert run_merge(readers, spec, pyoracle.SORT_LOSER_TREE) == expected_for_partial_update(data, False)
        for rod in (False, True):
            spec = AggregateMergeFunction.factory(
                {"fields.f1.aggregate-function": "sum", "aggregation.remove-record-on-delete": str(rod).lower()},
                VALUE_TYPE, ["f0"]).create()
            assert run_merge(readers, spec, pyoracle.SORT_LOSER_TREE) == expected_for_agg_sum(data, False, rod)


def test_drop_delete_reader():                        # DropDeleteReader.java:50-68
    readers = [parse("1, 1, +, 10 | 2, 2, +, 20"), parse("1, 3, -, 11 | 3, 4, -, 30")]
    spec = DeduplicateMergeFunction.factory().create().with_drop_delete()
    got = run_merge(readers, spec, pyoracle.SORT_LOSER_TREE)
    assert [(d.key, d.value) for d in got] == [(2, 20)]


def test_deduplicate_ignore_delete():                 # DeduplicateMergeFunction.java:47-55
    readers = [parse("1, 1, +, 10 | 2, 2, -, 20"), parse("1, 3, -, 11 | 2, 4, -, 21")]
    spec = DeduplicateMergeFunction.factory({"ignore-delete": "true"}).create()
    got = run_merge(readers, spec, pyoracle.SORT_LOSER_TREE)
    # key 1: retract ignored -> +10 kept; key 2: both retract -> null result -> nothing emitted
    assert [(d.key, d.sequence_number, d.value) for d in got] == [(1, 1, 10)]


def test_partial_update_delete_without_option_throws():   # PartialUpdateMergeFunction.java:155-164
    readers = [parse("1, 1, +, 10"), parse("1, 3, -, 11")]
    spec = PartialUpdateMergeFunction.factory({}, VALUE_TYPE, ["f0"]).create()
    with pytest.raises(pyoracle.OracleError, match="Partial update can not accept delete records"):
        run_merge(readers, spec, pyoracle.SORT_LOSER_TREE)
    # a lone DELETE passes through the wrapper untouched (ReducerMergeFunctionWrapper.java:70-72)
    got = run_merge([parse("1, 3, -, 11")], spec, pyoracle.SORT_LOSER_TREE)
    assert [(d.key, d.value_kind, d.value) for d in got] == [(1, RowKind.DELETE, 11)]


# ---------------------------------------------------------------- MergeFunction known-answer tests

def int_row_type(n, names=None):
    names = names or [f"f{i}" for i in range(n)]
    return RowType(tuple(DataField(nm, "INT", True) for nm in names))


class FuncDriver:
    """Drives a merge function the way the reference's unit tests do: add(...) then validate(...)
    on the running result.  State after n adds == merge of n single-row runs (same key, seq 0..n-1)
    with the wrapper bypassed."""

    def __init__(self, factory, row_type, pk=("f0",)):
        self.row_type = row_type
        self.schema = KeyValueSchema(RowType((DataField("_KEY_k", "INT", False),)), row_type)
        self.spec = factory.create()
        self.rows = []

    def reset(self):
        self.rows = []

    def add(self, *f, kind=RowKind.INSERT):
        self.rows.append((1, len(self.rows), int(kind)) + tuple(f))

    def result(self):
        runs = [KeyValueBatch.from_rows(self.schema, [r]) for r in self.rows]
        out = pyoracle.merge(self.schema, self.spec, runs, pyoracle.SORT_LOSER_TREE, bypass_wrapper=True)
        assert out.n_rows == 1
        return out.to_rows()[0]

    def validate(self, *f):
        assert self.result()[3:] == tuple(f)


def test_pu_update_non_null():                        # PartialUpdateMergeFunctionTest.java:42-62
    rt = int_row_type(7)
    d = FuncDriver(PartialUpdateMergeFunction.factory({}, rt, ["f0"]), rt)
    d.add(1, 1, 1, 1, 1, 1, 1)
    d.add(1, 2, 2, 2, 2, 2, None)
    d.validate(1, 2, 2, 2, 2, 2, 1)


SEQ_GROUP_OPTS = {"fields.f3.sequence-group": "f1,f2", "fields.f6.sequence-group": "f4,f5"}
D = RowKind.DELETE


def test_pu_sequence_group():                         # :64-97
    rt = int_row_type(7)
    d = FuncDriver(PartialUpdateMergeFunction.factory(SEQ_GROUP_OPTS, rt, ["f0"]), rt)
    d.add(1, 1, 1, 1, 1, 1, 1)
    d.add(1, 2, 2, 2, 2, 2, None)
    d.validate(1, 2, 2, 2, 1, 1, 1)
    d.add(1, 3, 3, 1, 3, 3, 3)
    d.validate(1, 2, 2, 2, 3, 3, 3)
    d.add(1, 1, 1, 3, 1, 1, None, kind=D)
    d.validate(1, None, None, 3, 3, 3, 3)
    d.add(1, 1, 1, 3, 1, 1, 4, kind=D)
    d.validate(1, None, None, 3, None, None, 4)
    d.add(1, 4, 4, 4, 5, 5, 5)
    d.validate(1, 4, 4, 4, 5, 5, 5)
    d.add(1, 1, 1, 6, 1, 1, 6, kind=D)
    d.validate(1, None, None, 6, None, None, 6)


def test_pu_sequence_group_partial_delete():          # :99-134
    rt = int_row_type(7)
    opts = dict(SEQ_GROUP_OPTS, **{"partial-update.remove-record-on-sequence-group": "f6"})
    d = FuncDriver(PartialUpdateMergeFunction.factory(opts, rt, ["f0"]), rt)
    d.add(1, 1, 1, 1, 1, 1, 1)
    d.add(1, 2, 2, 2, 2, 2, None)
    d.validate(1, 2, 2, 2, 1, 1, 1)
    d.add(1, 3, 3, 1, 3, 3, 3)
    d.validate(1, 2, 2, 2, 3, 3, 3)
    d.add(1, 1, 1, 3, 1, 1, None, kind=D)
    d.validate(1, None, None, 3, 3, 3, 3)
    d.add(1, 1, 1, 3, 1, 1, 4, kind=D)
    d.validate(1, 1, 1, 3, 1, 1, 4)
    d.add(1, 4, 4, 4, 5, 5, 5)
    d.validate(1, 4, 4, 4, 5, 5, 5)
    d.add(1, 1, 1, 6, 1, 1, 6, kind=D)
    d.validate(1, 1, 1, 6, 1, 1, 6)


MULTI_SEQ_OPTS = {"fields.f3,f4.sequence-group": "f1,f2", "fields.f7,f8.sequence-group": "f5,f6"}


def test_pu_multi_sequence_fields():                  # :175-217
    rt = int_row_type(9)
    d = FuncDriver(PartialUpdateMergeFunction.factory(MULTI_SEQ_OPTS, rt, ["f0"]), rt)
    d.add(1, None, None, None, None, 1, 1, 1, 3)
    d.add(1, 2, 2, None, None, 2, 2, 1, 3)
    d.validate(1, None, None, None, None, 2, 2, 1, 3)
    d.reset()
    d.add(1, 1, 1, 1, 1, 1, 1, 1, 3)
    d.add(1, 2, 2, 2, 2, 2, 1, 1, None)
    d.validate(1, 2, 2, 2, 2, 1, 1, 1, 3)
    d.add(1, 1, 3, 1, 3, 3, 3, 3, 2)
    d.validate(1, 2, 2, 2, 2, 3, 3, 3, 2)
    d.add(1, 1, 1, 3, 3, 1, 1, None, None, kind=D)
    d.validate(1, None, None, 3, 3, 3, 3, 3, 2)
    d.add(1, 1, 1, 3, 1, 1, 1, 4, 4, kind=D)
    d.validate(1, None, None, 3, 3, None, None, 4, 4)
    d.add(1, 4, 4, 4, 4, 5, 5, 5, 5)
    d.validate(1, 4, 4, 4, 4, 5, 5, 5, 5)
    d.add(1, 1, 1, 6, 1, 1, 1, 6, 1, kind=D)
    d.validate(1, None, None, 6, 1, None, None, 6, 1)


def test_pu_sequence_group_default_agg_func():        # :219-246
    rt = int_row_type(7)
    opts = dict(SEQ_GROUP_OPTS, **{"fields.default-aggregate-function": "last_non_null_value"})
    d = FuncDriver(PartialUpdateMergeFunction.factory(opts, rt, ["f0"]), rt)
    d.add(1, 1, 1, 1, 1, 1, 1)
    d.add(1, 2, 2, 2, 2, 2, None)
    d.validate(1, 2, 2, 2, 1, 1, 1)
    d.add(1, 3, 3, 1, 3, 3, 3)
    d.validate(1, 2, 2, 2, 3, 3, 3)
    d.add(1, 4, None, 4, 5, None, 5)
    d.validate(1, 4, 2, 4, 5, 3, 5)


def test_pu_multi_sequence_fields_default_agg_func():  # :248-275
    rt = int_row_type(9)
    opts = dict(MULTI_SEQ_OPTS, **{"fields.default-aggregate-function": "last_non_null_value"})
    d = FuncDriver(PartialUpdateMergeFunction.factory(opts, rt, ["f0"]), rt)
    d.add(1, 1, 1, 1, 1, 1, 1, 1, 1)
    d.add(1, 2, 2, 2, 2, 2, 2, None, None)
    d.validate(1, 2, 2, 2, 2, 1, 1, 1, 1)
    d.add(1, 3, 3, 1, 1, 3, 3, 3, 3)
    d.validate(1, 2, 2, 2, 2, 3, 3, 3, 3)
    d.add(1, 4, None, 4, 4, 5, None, 5, 5)
    d.validate(1, 4, 2, 4, 4, 5, 3, 5, 5)


def test_pu_first_value():                            # :569-589
    rt = int_row_type(4)
    opts = {"fields.f1.sequence-group": "f2,f3", "fields.f2.aggregate-function": "first_value",
            "fields.f3.aggregate-function": "last_value"}
    d = FuncDriver(PartialUpdateMergeFunction.factory(opts, rt, ["f0"]), rt)
    d.add(1, 1, 1, 1)
    d.add(1, 2, 2, 2)
    d.validate(1, 2, 1, 2)
    d.add(1, 0, 3, 3)
    d.validate(1, 2, 3, 2)


def test_pu_multi_sequence_fields_first_value():      # :591-616
    rt = int_row_type(5)
    opts = {"fields.f1,f2.sequence-group": "f3,f4", "fields.f3.aggregate-function": "first_value",
            "fields.f4.aggregate-function": "last_value"}
    d = FuncDriver(PartialUpdateMergeFunction.factory(opts, rt, ["f0"]), rt)
    d.add(1, 1, 1, 1, 1)
    d.add(1, 2, 2, 2, 2)
    d.validate(1, 2, 2, 1, 2)
    d.add(1, 0, 1, 3, 3)
    d.validate(1, 2, 2, 3, 2)


def test_pu_with_aggregation():                       # :618-677 (f0 is the pk: listagg option is never read)
    rt = int_row_type(8)
    opts = {"fields.f1.sequence-group": "f2,f3,f4", "fields.f7.sequence-group": "f6",
            "fields.f0.aggregate-function": "listagg", "fields.f2.aggregate-function": "sum",
            "fields.f4.aggregate-function": "last_value", "fields.f6.aggregate-function": "last_non_null_value",
            "fields.f4.ignore-retract": "true", "fields.f6.ignore-retract": "true"}
    d = FuncDriver(PartialUpdateMergeFunction.factory(opts, rt, ["f0"]), rt)
    d.add(1, 1, 1, 1, 1, 1, 1, 1)
    d.add(1, 2, 1, 2, 2, 2, 2, 0)
    d.validate(1, 2, 2, 2, 2, 2, 1, 1)
    d.add(1, 1, 1, 1, 1, 1, 2, 0)
    d.validate(1, 2, 3, 2, 2, 1, 1, 1)
    d.add(1, 1, -1, 1, 1, 2, 2, 0)
    d.add(1, 3, None, None, None, None, None, 2)
    d.validate(1, 3, 2, None, None, 2, 1, 2)
    d.add(1, 3, 1, 1, 1, 1, 1, 3)
    d.validate(1, 3, 3, 1, 1, 1, 1, 3)
    d.add(1, 3, 2, 1, 1, 1, 1, 3, kind=RowKind.UPDATE_BEFORE)
    d.validate(1, 3, 1, None, 1, 1, 1, 3)
    d.add(1, 3, 2, 1, 1, 1, 1, 3, kind=D)
    d.validate(1, 3, -1, None, 1, 1, 1, 3)
    d.add(1, 2, 2, 1, 1, 1, 1, 3, kind=D)
    d.validate(1, 3, -3, None, 1, 1, 1, 3)


def test_pu_multi_sequence_fields_with_aggregation():  # :679-742
    rt = int_row_type(9)
    opts = {"fields.f1,f2.sequence-group": "f3,f4,f5", "fields.f7,f8.sequence-group": "f6",
            "fields.f0.aggregate-function": "listagg", "fields.f3.aggregate-function": "sum",
            "fields.f4.aggregate-function": "first_value", "fields.f5.aggregate-function": "last_value",
            "fields.f6.aggregate-function": "last_non_null_value",
            "fields.f4.ignore-retract": "true", "fields.f6.ignore-retract": "true"}
    d = FuncDriver(PartialUpdateMergeFunction.factory(opts, rt, ["f0"]), rt)
    d.add(1, None, None, 1, 1, 1, 1, 1, 1)
    d.validate(1, None, None, None, None, None, 1, 1, 1)
    d.add(1, None, None, 1, 1, 1, 0, 1, 1, kind=D)
    d.validate(1, None, None, None, None, None, 1, 1, 1)
    d.add(1, 1, 1, 1, 1, 1, 1, 1, 1)
    d.add(1, 1, 2, 1, 2, 2, None, 2, 0)
    d.validate(1, 1, 2, 2, 1, 2, 1, 2, 0)
    d.add(1, 1, 1, 1, 3, 1, 1, 2, 0)
    d.validate(1, 1, 2, 3, 3, 2, 1, 2, 0)
    d.add(1, 1, 3, None, None, None, None, 4, 2)
    d.validate(1, 1, 3, 3, 3, None, 1, 4, 2)
    d.add(1, 2, 3, 1, 1, 1, 1, 4, 3)
    d.validate(1, 2, 3, 4, 3, 1, 1, 4, 3)
    d.add(1, 2, 3, 2, 1, 2, 1, 4, 3, kind=RowKind.UPDATE_BEFORE)
    d.validate(1, 2, 3, 2, 3, None, 1, 4, 3)
    # the reference passes 8 values here (GenericRow.of(1,3,2,3,1,1,4,3)); the 9th field is absent ->
    # reading it would fail in Java; the test only works because that field is never touched.  We pad null.
    d.add(1, 3, 2, 3, 1, 1, 4, 3, None, kind=D)
    d.validate(1, 3, 2, -1, 3, None, 1, 4, 3)
    d.add(1, 2, 2, 2, 1, 1, 1, 1, 3, kind=D)
    d.validate(1, 3, 2, -3, 3, None, 1, 4, 3)


def test_agg_default_agg_func():                      # AggregateMergeFunctionTest.java:42-70
    rt = int_row_type(5, ["k", "a", "b", "c", "d"])
    opts = {"fields.default-aggregate-function": "first_non_null_value", "fields.b.aggregate-function": "sum"}
    d = FuncDriver(AggregateMergeFunction.factory(opts, rt, ["k"]), rt)
    d.add(1, None, 1, 1, 1)
    d.add(1, 2, None, 2, 2)
    d.add(1, 3, 3, None, 3)
    d.add(1, 4, 4, 4, None)
    d.add(1, 5, 5, 5, 5)
    d.validate(1, 2, 13, 1, 1)


def fold(agg_name, sql_type, steps, ignore_retract=False):
    """Fold [(kind, value), ...] through one aggregator column; returns the final accumulator."""
    rt = RowType((DataField("k", "INT", False), DataField("v", sql_type, True)))
    opts = {"fields.v.aggregate-function": agg_name}
    if ignore_retract:
        opts["fields.v.ignore-retract"] = "true"
    d = FuncDriver(AggregateMergeFunction.factory(opts, rt, ["k"]), rt)
    for kind, v in steps:
        d.add(1, v, kind=kind)
    return d.result()[4]


I, R = RowKind.INSERT, RowKind.DELETE


@pytest.mark.parametrize("t", ["TINYINT", "SMALLINT", "INT", "BIGINT", "FLOAT", "DOUBLE"])
def test_field_sum_and_product_agg(t):                # FieldAggregatorTest.java:431-543
    assert fold("sum", t, [(I, 10)]) == 10            # agg(null, 10)
    assert fold("sum", t, [(I, 1), (I, 10)]) == 11
    assert fold("sum", t, [(I, 10), (R, 5)]) == 5     # retract(10, 5)
    assert fold("sum", t, [(R, 5)]) == -5             # retract(null, 5)
    assert fold("product", t, [(I, 10)]) == 10
    assert fold("product", t, [(I, 1), (I, 10)]) == 10
    assert fold("product", t, [(I, 10), (R, 5)]) == 2
    assert fold("product", t, [(R, 5)]) is None


def test_field_sum_wraps_like_java():                 # FieldSumAgg.java:57-62 casts back to byte/short
    assert fold("sum", "TINYINT", [(I, 127), (I, 1)]) == -128
    assert fold("sum", "SMALLINT", [(I, 32767), (I, 1)]) == -32768
    assert fold("sum", "INT", [(I, 2 ** 31 - 1), (I, 1)]) == -2 ** 31
    assert fold("sum", "BIGINT", [(I, 2 ** 63 - 1), (I, 1)]) == -2 ** 63


def test_field_sum_double_is_left_fold():             # FieldSumAgg.java:69-74
    vals = [1e16, 1.0, -1e16, 1.0]
    want = 0.0
    first = True
    for v in vals:
        want = v if first else want + v
        first = False
    assert fold("sum", "DOUBLE", [(I, v) for v in vals]) == want == 1.0


def test_field_max_min_agg():                         # :415-428
    assert fold("max", "INT", [(I, 1), (I, 10)]) == 10
    assert fold("min", "INT", [(I, 1), (I, 10)]) == 1
    assert fold("max", "STRING", [(I, "abc"), (I, "abd")]) == "abd"
    assert fold("min", "STRING", [(I, "abc"), (I, "ab")]) == "ab"
    # Double.compare total order (InternalRowUtils.java:412-414): -0.0 < 0.0, NaN is the largest
    assert str(fold("max", "DOUBLE", [(I, -0.0), (I, 0.0)])) == "0.0"
    assert str(fold("min", "DOUBLE", [(I, 0.0), (I, -0.0)])) == "-0.0"
    assert np.isnan(fold("max", "DOUBLE", [(I, float("nan")), (I, 1.0)]))
    with pytest.raises(pyoracle.OracleError, match="does not support retraction"):
        fold("max", "INT", [(I, 1), (R, 10)])
    assert fold("max", "INT", [(I, 1), (R, 10)], ignore_retract=True) == 1   # FieldIgnoreRetractAgg


def test_field_bool_aggs():                           # :94-107
    assert fold("bool_and", "BOOLEAN", [(I, 0), (I, 1)]) == 0
    assert fold("bool_and", "BOOLEAN", [(I, 1), (I, 1)]) == 1
    assert fold("bool_or", "BOOLEAN", [(I, 0), (I, 1)]) == 1
    assert fold("bool_or", "BOOLEAN", [(I, 0), (I, 0)]) == 0


def test_field_last_first_aggs():                     # :110-157
    assert fold("last_non_null_value", "INT", [(I, 1)]) == 1
    assert fold("last_non_null_value", "INT", [(I, 1), (I, None)]) == 1
    assert fold("last_value", "INT", [(I, 1)]) == 1
    assert fold("last_value", "INT", [(I, 1), (I, None)]) is None
    assert fold("first_value", "INT", [(I, 1), (I, 2)]) == 1
    assert fold("first_value", "INT", [(I, None), (I, 2)]) is None
    assert fold("first_non_null_value", "INT", [(I, None), (I, None)]) is None
    assert fold("first_non_null_value", "INT", [(I, None), (I, 1), (I, 2)]) == 1
    assert fold("last_value", "INT", [(I, 1), (R, 5)]) is None               # FieldLastValueAgg.retract
    assert fold("last_non_null_value", "INT", [(I, 1), (R, 5)]) is None
    assert fold("last_non_null_value", "INT", [(I, 1), (R, None)]) == 1


# ---------------------------------------------------------------- IntervalPartition

def _sections(files):
    mn = [f[0] for f in files]
    mx = [f[1] for f in files]
    sec, run, n = pyoracle.interval_partition(mn, mx)
    out = [dict() for _ in range(n)]
    for i, (s, r) in enumerate(zip(sec.tolist(), run.tolist())):
        out[s].setdefault(r, []).append(files[i])
    return [sorted(sorted(fs) for fs in s.values()) for s in out]


def test_interval_partition_fixed():                  # IntervalPartitionTest.java:59-69 ("[lo,hi]" lists)
    # runTest("[1, 2] [3, 4] [5, 180] [5, 190] [200, 600] [210, 700]", ...)
    files = [(1, 2), (3, 4), (5, 180), (5, 190), (200, 600), (210, 700)]
    got = _sections(files)
    assert got == [[[(1, 2)]], [[(3, 4)]], [[(5, 180)], [(5, 190)]], [[(200, 600)], [(210, 700)]]]
    # a chain packs into one run; overlapping files open new runs
    files = [(1, 10), (11, 20), (21, 30), (5, 25)]
    got = _sections(files)
    assert len(got) == 1 and sorted(got[0]) == [[(1, 10), (11, 20), (21, 30)], [(5, 25)]]


def test_interval_partition_random_properties():      # IntervalPartitionTest.testRandom invariants
    rng = random.Random(2024)
    for _ in range(200):
        n = rng.randrange(1, 30)
        files = []
        for _ in range(n):
            lo = rng.randrange(0, 200)
            files.append((lo, lo + rng.randrange(0, 40)))
        secs = _sections(files)
        flat = sorted(f for s in secs for r in s for f in r)
        assert flat == sorted(files)
        prev_max = None
        for s in secs:
            lo = min(f[0] for r in s for f in r)
            hi = max(f[1] for r in s for f in r)
            if prev_max is not None:
                assert lo > prev_max                  # sections are key-disjoint and ordered
            prev_max = hi
            for r in s:                               # files inside a run do not overlap
                for a, b in zip(r, r[1:]):
                    assert a[1] < b[0]


# ---------------------------------------------------------------- committed golden file

def test_golden_file_is_current():
    """tests/golden/sort_merge_reader_vectors.json is what tests/golden/make_golden.py writes."""
    import json
    import subprocess
    import sys
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    before = open(os.path.join(here, "golden", "sort_merge_reader_vectors.json")).read()
    subprocess.check_call([sys.executable, os.path.join(here, "golden", "make_golden.py")], stdout=subprocess.DEVNULL)
    assert open(os.path.join(here, "golden", "sort_merge_reader_vectors.json")).read() == before
    assert len(json.loads(before)["cases"]) == 25


@pytest.mark.parametrize("engine", ENGINES)
def test_oracle_matches_golden_file(engine):
    from golden_util import load_cases, records, spec_for
    n = 0
    for case in load_cases():
        readers = [records(r) for r in case["readers"]]
        for name, want in case["expected"].items():
            got = run_merge(readers, spec_for(name), engine)
            assert got == records(want), (case["name"], name)
            n += 1
    assert n >= 100


def test_string_key_compaction_shape_engines_agree():
    """The C4 shape of bench.py (hex string key, deletes, drop-delete): loser tree, min-heap and brute force agree."""
    import bench
    from paimon_b200 import datagen
    schema = bench.make_schema("c4")
    spec = bench.make_spec("c4", schema)
    runs = datagen.make_runs(schema, 7, 7000, seed=5, null_prob=0.5, delete_prob=0.05)
    outs = [pyoracle.merge(schema, spec, runs, e) for e in ENGINES]
    assert outs[0].equals(outs[1]) and outs[0].equals(outs[2])
    kinds = np.asarray(outs[0].columns[schema.n_key + 1].data[: outs[0].n_rows])
    assert not np.isin(kinds, [1, 3]).any()             # drop-delete
    keys = outs[0].columns[0].to_pylist()
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
