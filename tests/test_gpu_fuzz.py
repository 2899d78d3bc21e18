"""GPU suite (-m gpu): randomly generated schemas and query trees, libcolq.so vs the CPU oracle, bit for bit.

The reference's tests pin five hand-written queries (QueryTest.java); this widens them mechanically: random tables with
int and string columns, random to-one (with Nones) and to-many associations between them (self associations included),
and random query trees that walk forward and reverse association columns with criteria on any level -- i.e. random
instances of exactly the semantics of E/DataSystemSerialIndices.java:53-102.  Every case runs in several physical layouts
(device / host-resident / dictionary-encoded) and planner strategies (lazy / eager / chains not deferred).
"""
import pytest

from fuzz_cases import make_case
from test_gpu_parity import ALL_VARIANTS, both, engines  # noqa: F401  (fixture + comparison helper)

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("seed", range(24))
def test_random_schema_and_query_trees(engines, seed):  # noqa: F811
    build, queries = make_case(seed)
    variants = ALL_VARIANTS if seed % 3 == 0 else ALL_VARIANTS[seed % len(ALL_VARIANTS):][:2] + ALL_VARIANTS[:1]
    both(engines, build, queries, variants=variants)
