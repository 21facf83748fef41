#!/usr/bin/env python
"""Writes tests/golden/sort_merge_reader_vectors.json: the reference's fixed SortMergeReader vectors
(paimon-core/src/test/java/org/apache/paimon/mergetree/compact/SortMergeReaderTestBase.java:51-89, the string DSL of
ReusingTestData.java:82-106) plus 20 seeded random rounds in the shape of CombiningRecordReaderTestBase.java:54-80,
each with the result the reference's own expected-result calculators give (MergeFunctionTestUtils.java:35-150,
transcribed in tests/reusing_test_data.py).  Language neutral: the Java side can replay the same file.

Record = [key, sequence, kind ('+' insert / '-' delete), value].  Run with:  python tests/golden/make_golden.py
(no GPU, no reference checkout needed: the vectors are transcribed, the calculators are plain Python)."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from reusing_test_data import (expected_for_agg_sum, expected_for_deduplicate, expected_for_first_row,  # noqa: E402
                               expected_for_partial_update, generate_random_readers, parse)
from test_oracle_golden import FIXED_VECTORS                                                            # noqa: E402


def rec(d):
    return [d.key, d.sequence_number, "+" if int(d.value_kind) in (0, 2) else "-", d.value]


def case(name, readers, add_only):
    flat = [d for r in readers for d in r]
    out = {"name": name, "add_only": add_only, "readers": [[rec(d) for d in r] for r in readers],
           "expected": {"deduplicate": [rec(d) for d in expected_for_deduplicate(flat)]}}
    if add_only:
        out["expected"]["first_row"] = [rec(d) for d in expected_for_first_row(flat)]
    out["expected"]["partial_update_ignore_delete" if not add_only else "partial_update"] = \
        [rec(d) for d in expected_for_partial_update(flat, add_only)]
    out["expected"]["agg_sum"] = [rec(d) for d in expected_for_agg_sum(flat, add_only, False)]
    out["expected"]["agg_sum_remove_record_on_delete"] = [rec(d) for d in expected_for_agg_sum(flat, add_only, True)]
    return out


def main():
    cases = [case(name, [parse(s) for s in readers], True) for name, readers in sorted(FIXED_VECTORS.items())]
    rng = random.Random(20240922)
    for i in range(20):
        add_only = i % 2 == 0
        cases.append(case(f"random_{i}", generate_random_readers(rng, add_only, max_readers=8, max_rows=30), add_only))
    with open(os.path.join(HERE, "sort_merge_reader_vectors.json"), "w") as f:
        json.dump({"source": "SortMergeReaderTestBase.java:51-89, CombiningRecordReaderTestBase.java:54-80, "
                             "MergeFunctionTestUtils.java:35-150", "cases": cases}, f, indent=None, separators=(",", ":"))
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
