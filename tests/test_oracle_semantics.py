"""CPU suite: the C oracle against a second, independent restatement of the reference engine (tests/java_model.py).

The reference's own tests pin five hand-written queries (tests/test_oracle_goldens.py).  Here the oracle is compared on
randomly generated schemas and query trees -- the generator of the GPU fuzz suite, with small tables -- against

  * a row-object model that follows the Java line by line (per-row ``Association`` objects taken from the cross-linked
    reverse column, LIFO leaf walks, sets as BitSets): matched rows AND every node's final cardinality, and
  * the declarative reading of the result (every criterion holds, every child subtree has a matching associated row),
    which is the property the GPU planner's single post-order pass relies on.

So the chain of evidence is: reference goldens -> oracle; oracle == row-object model == declarative semantics on random
cases (here, CPU); libcolq.so == oracle on the same random cases at larger sizes (tests/test_gpu_fuzz.py, GPU).
"""
import numpy as np
import pytest

import tck
from colq import Association, Criteria, InMemoryTable, Query, int_range, of_columns, of_ints, of_strings, str_contains
from colq.in_memory import IntegerColumn
from fuzz_cases import make_case
from java_model import JavaModelDataSystem, declarative_matches
from oracle_system import OracleDataSystem

SMALL = [1, 2, 37, 64, 65, 300]


@pytest.mark.parametrize("seed", range(40))
def test_oracle_equals_row_object_model_and_declarative_semantics(seed):
    build, queries = make_case(seed, size_choices=SMALL)
    oracle, model = OracleDataSystem(), JavaModelDataSystem()
    build(oracle)
    build(model)
    for make_query in queries:
        ro, rm = oracle.execute(make_query()), model.execute(make_query())
        assert type(ro).__name__ == type(rm).__name__ == "Success", (ro, rm)
        assert oracle.last_indices.tolist() == model.last_indices
        assert oracle.node_cardinalities() == model.node_cardinalities()
        assert declarative_matches(model.tables, make_query()) == model.last_indices
    oracle.close()


@pytest.mark.parametrize("case", tck.REFERENCE_TESTS + tck.FAILURE_TESTS, ids=lambda f: f.__name__)
def test_row_object_model_passes_the_reference_cases(case):
    """The model itself is pinned to the reference's QueryTest vectors and failure messages, like the oracle."""
    case(JavaModelDataSystem)


def test_leaf_walk_order_does_not_matter():
    """A node with a leaf child and a deeper subtree: the leaf walks pass it several times (LIFO over BFS order,
    E/DataSystemSerialIndices.java:92-97); the last walk through a node carries its final bits, so the root ends at the
    declarative fixed point whatever the order."""
    rng = np.random.default_rng(7)
    n = 120
    a = InMemoryTable.of_columns(IntegerColumn(rng.integers(0, 10, n, dtype=np.int32)))
    b = InMemoryTable.of_columns(IntegerColumn(rng.integers(0, 10, n, dtype=np.int32)))
    c = InMemoryTable.of_columns(IntegerColumn(rng.integers(0, 10, n, dtype=np.int32)))
    d = InMemoryTable.of_columns(IntegerColumn(rng.integers(0, 10, n, dtype=np.int32)))
    a.associate_to(b, fk=rng.integers(-1, n, n, dtype=np.int32))   # a.1 -> b
    a.associate_to(c, fk=rng.integers(-1, n, n, dtype=np.int32))   # a.2 -> c
    c.associate_to(d, fk=rng.integers(-1, n, n, dtype=np.int32))   # c.2 -> d (c.1 is the reverse of a.2)

    def make():
        q = Query("a")
        q.root_node.add_criteria(Criteria.IntCriteria(0, int_range(0, 8)))
        q.root_node.create_child(1).add_criteria(Criteria.IntCriteria(0, int_range(2, 9)))
        mid = q.root_node.create_child(2)
        mid.create_child(2).add_criteria(Criteria.IntCriteria(0, int_range(0, 4)))
        return q

    oracle, model = OracleDataSystem(), JavaModelDataSystem()
    for ds in (oracle, model):
        for name, t in (("a", a), ("b", b), ("c", c), ("d", d)):
            ds.register(name, t)
        ds.execute(make())
    assert oracle.last_indices.tolist() == model.last_indices == declarative_matches(model.tables, make())
    assert 0 < len(model.last_indices) < n
    assert oracle.node_cardinalities() == model.node_cardinalities()


def test_many_self_association_two_hops():
    """The shape of QueryTest.java:231 (states adjacent to states adjacent to ...) with None / One / Many rows."""
    names = of_strings("North A", "South B", "North C", "East D", "South E")
    t = of_columns(names, of_ints(0, 1, 2, 3, 4))
    t.associate_to(t, Association.to_many(1, 3), Association.to_one(2), Association.to_none(), Association.to_many(0, 4),
                   Association.to_one(0))
    q = Query("t")
    q.root_node.add_criteria(Criteria.StringCriteria(0, str_contains("North")))
    q.root_node.create_child(2).add_criteria(Criteria.StringCriteria(0, str_contains("South"))) \
        .create_child(2).add_criteria(Criteria.StringCriteria(0, str_contains("North")))
    oracle, model = OracleDataSystem(), JavaModelDataSystem()
    for ds in (oracle, model):
        ds.register("t", t)
        ds.execute(q)
    assert oracle.last_indices.tolist() == model.last_indices == declarative_matches(model.tables, q) == [0]
