"""Loader for tests/golden/sort_merge_reader_vectors.json (see tests/golden/make_golden.py)."""
import json
import os

from paimon_b200.merge_function import (AggregateMergeFunction, DeduplicateMergeFunction, FirstRowMergeFunction,
                                        PartialUpdateMergeFunction)
from paimon_b200.types import RowKind

from reusing_test_data import VALUE_TYPE, ReusingTestData

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sort_merge_reader_vectors.json")


def load_cases():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


def records(rows):
    return [ReusingTestData(k, s, RowKind.INSERT if kind == "+" else RowKind.DELETE, v) for k, s, kind, v in rows]


def spec_for(expected_name):
    """The merge function whose result the golden entry holds (option sets of SortBufferWriteBufferTestBase.java:
    160-235 / MergeFunctionTestUtils.java)."""
    if expected_name == "deduplicate":
        return DeduplicateMergeFunction.factory().create()
    if expected_name == "first_row":
        return FirstRowMergeFunction.factory().create()
    if expected_name == "partial_update":
        return PartialUpdateMergeFunction.factory({}, VALUE_TYPE, ["f0"]).create()
    if expected_name == "partial_update_ignore_delete":
        return PartialUpdateMergeFunction.factory({"ignore-delete": "true"}, VALUE_TYPE, ["f0"]).create()
    if expected_name == "agg_sum":
        return AggregateMergeFunction.factory({"fields.f1.aggregate-function": "sum"}, VALUE_TYPE, ["f0"]).create()
    if expected_name == "agg_sum_remove_record_on_delete":
        return AggregateMergeFunction.factory({"fields.f1.aggregate-function": "sum",
                                               "aggregation.remove-record-on-delete": "true"}, VALUE_TYPE, ["f0"]).create()
    raise KeyError(expected_name)
