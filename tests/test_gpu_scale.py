"""Parity at the BASELINE.json shapes, sized so that the CPU oracle still finishes in seconds:
the full 50-column C3 schema with 16 runs, C2, the C4 string-key shape with 32 runs and deletes, and C3-agg — each at
millions of rows, i.e. three sampled partition levels above level 0, thousands of plan tiles, and emit tiles that take
their var-len byte bases from a look-back over hundreds of predecessors.  Every column is compared bit for bit."""
import numpy as np
import pytest

from oracle import pyoracle
from paimon_b200 import datagen
from paimon_b200.merge_function import (AggregateMergeFunction, DeduplicateMergeFunction,
                                        PartialUpdateMergeFunction)
from paimon_b200.sort_merge_reader import merge_runs
from paimon_b200.types import DataField, KeyValueSchema, RowType

pytestmark = pytest.mark.gpu


def assert_same(schema, spec, runs):
    want = pyoracle.merge(schema, spec, runs, pyoracle.SORT_LOSER_TREE)
    got = merge_runs(schema, spec, runs)
    assert got.n_rows == want.n_rows
    assert got.equals(want), got.first_difference(want)
    return got


def test_c3_full_schema_16_runs_5m_rows():
    schema = datagen.schema_c3()                                    # pk + 20 BIGINT + 15 DOUBLE + 14 VARCHAR
    assert schema.n_cols == 53
    runs = datagen.make_runs(schema, 16, 5_000_000, seed=31, null_prob=0.5)
    spec = PartialUpdateMergeFunction.factory({}, schema.value_type, ["pk"]).create()
    got = assert_same(schema, spec, runs)
    keys = got.columns[0].data
    assert np.all(keys[1:] > keys[:-1])


def test_c2_8_runs_6m_rows():
    schema = datagen.schema_c2()
    runs = datagen.make_runs(schema, 8, 6_000_000, seed=32)
    assert_same(schema, DeduplicateMergeFunction.factory().create(), runs)


def test_c3agg_16_runs_2m_rows():
    schema = datagen.schema_c3()
    runs = datagen.make_runs(schema, 16, 2_000_000, seed=33, null_prob=0.5)
    opts = {f"fields.{f.name}.aggregate-function": "sum" for f in schema.value_type.fields
            if f.name != "pk" and f.physical.name in ("INT64", "DOUBLE")}
    spec = AggregateMergeFunction.factory(opts, schema.value_type, ["pk"]).create()
    assert_same(schema, spec, runs)


def schema_c4():
    fields = [DataField("pk", "VARCHAR(16)", False)]
    fields += [DataField(f"i{i}", "BIGINT", True) for i in range(4)]
    fields += [DataField(f"d{i}", "DOUBLE", True) for i in range(2)]
    fields += [DataField(f"n{i}", "INT", True) for i in range(2)]
    fields += [DataField(f"s{i}", "VARCHAR(64)", True) for i in range(3)]
    return KeyValueSchema.of(RowType(tuple(fields)), ["pk"])


@pytest.mark.parametrize("drop_delete", [True, False])
def test_c4_string_key_32_runs_with_deletes(drop_delete):
    schema = schema_c4()
    runs = datagen.make_runs(schema, 32, 2_000_000, seed=34, null_prob=0.5, delete_prob=0.05)
    spec = DeduplicateMergeFunction.factory().create()
    if drop_delete:
        spec = spec.with_drop_delete()
    assert_same(schema, spec, runs)
