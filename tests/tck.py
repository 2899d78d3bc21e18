"""Test Compatibility Kit: the reference's functional query tests restated once, run against any DataSystem.

The reference wishes for exactly this (README.md:149-153).  Each function mirrors one test of
data-system-serial-indices-arrays/src/test/java/dgroomes/queryengine/QueryTest.java (cited per function) and takes a
factory returning a fresh engine with ``register(name, table)`` / ``execute(query)``: the CPU oracle
(oracle/oracle_system.py) in the CPU suite, ``colq.DataSystemColq`` (libcolq.so on a B200) in the GPU suite.
"""
from colq import (Association, BooleanColumn, Criteria, InMemoryTable, Query, QueryResult, int_between_exclusive,
                  int_greater_than, int_range, of_columns, of_ints, of_strings, str_compare_gt, str_compare_lt,
                  str_contains, str_equals)
from colq import geography as G
from colq.in_memory import IntegerColumn, StringColumn


def failed(msg):
    return AssertionError(msg)  # TestUtil.failed (TestUtil.java:10-40)


def success_columns(result):
    if isinstance(result, QueryResult.Failure):
        raise failed(result.message)
    assert isinstance(result, QueryResult.Success)
    return result.result_set.columns()


def int_query_one_column_table(new_system):
    """QueryTest.intQuery_oneColumnTable (QueryTest.java:37-73)."""
    ds = new_system()
    table = of_columns(of_ints(-1, 0, 1, 2, 3))
    ds.register("ints", table)
    query = Query("ints")
    query.root_node.add_criteria(Criteria.IntCriteria(0, int_greater_than(0)))
    columns = success_columns(ds.execute(query))
    assert len(columns) == 1
    assert isinstance(columns[0], IntegerColumn)
    assert columns[0].ints().tolist() == [1, 2, 3]


def int_query_two_column_table(new_system):
    """QueryTest.intQuery_twoColumnTable (QueryTest.java:78-108)."""
    ds = new_system()
    table = of_columns(of_strings("Minneapolis", "Rochester", "Duluth"), of_ints(425_336, 121_395, 86_697))
    ds.register("cities", table)
    query = Query("cities")
    query.root_node.add_criteria(Criteria.IntCriteria(1, int_between_exclusive(100_000, 150_000)))
    columns = success_columns(ds.execute(query))
    assert len(columns) == 2
    assert isinstance(columns[0], StringColumn)
    assert columns[0].strings() == ["Rochester"]


def multi_criteria_root_entity(new_system):
    """QueryTest.multiCriteria_rootEntity (QueryTest.java:113-144)."""
    ds = new_system()
    table = of_columns(of_strings("a", "a", "b", "c", "c", "d"))
    ds.register("strings", table)
    query = Query("strings")
    (query.root_node.add_criteria(Criteria.StringCriteria(0, str_compare_gt("a")))
        .add_criteria(Criteria.StringCriteria(0, str_compare_lt("d"))))
    columns = success_columns(ds.execute(query))
    assert len(columns) == 1
    assert columns[0].strings() == ["b", "c", "c"]


def query_on_association_property(new_system):
    """QueryTest.queryOnAssociationProperty (QueryTest.java:150-229)."""
    ds = new_system()
    cities = of_columns(of_strings("Minneapolis", "Pierre", "Duluth"))
    ds.register("cities", cities)
    states = of_columns(of_strings("Minnesota", "South Dakota"))
    ds.register("states", states)
    cities.associate_to(states, Association.to_one(0), Association.to_one(1), Association.to_one(0))
    for state, want in (("South Dakota", ["Pierre"]), ("Minnesota", ["Minneapolis", "Duluth"])):
        query = Query("cities")
        query.root_node.create_child(1).add_criteria(Criteria.StringCriteria(0, str_equals(state)))
        columns = success_columns(ds.execute(query))
        assert len(columns) == 2
        assert columns[0].strings() == want


def multi_criteria_including_intermediate_entity(new_system):
    """QueryTest.multiCriteria_includingIntermediateEntity (QueryTest.java:231-343)."""
    ds = new_system()
    sections = of_columns(
        of_strings("maple trees", "lilacs", "", "", "", "", "Boston ferns", "rose bush", "cedar trees"),
        of_strings("trees", "shrubs", "", "", "", "", "ferns", "shrubs", "trees"))
    ds.register("sections", sections)
    sections.associate_to(sections,
                          Association.to_many(1, 3), Association.to_many(0, 2, 4), Association.to_many(1, 5),
                          Association.to_many(0, 4, 6), Association.to_many(1, 3, 5, 7), Association.to_many(2, 4, 8),
                          Association.to_many(3, 7), Association.to_many(4, 6, 8), Association.to_many(5, 7))
    query = Query("sections")
    (query.root_node.add_criteria(Criteria.StringCriteria(1, str_equals("trees")))
        .create_child(2).add_criteria(Criteria.StringCriteria(1, str_equals("shrubs")))
        .create_child(2).add_criteria(Criteria.StringCriteria(1, str_equals("ferns"))))
    columns = success_columns(ds.execute(query))
    assert len(columns) == 4
    assert columns[0].strings() == ["cedar trees"]


REFERENCE_TESTS = [int_query_one_column_table, int_query_two_column_table, multi_criteria_root_entity,
                   query_on_association_property, multi_criteria_including_intermediate_entity]


# ---- failure paths: never asserted by the reference's tests; messages follow Verifier.java / DataSystemSerialIndices.java
def failure_unregistered_table(new_system):
    ds = new_system()
    r = ds.execute(Query("nope"))  # DataSystemSerialIndices.java:54-57
    assert isinstance(r, QueryResult.Failure)
    assert r.message == "The query targets the table 'nope' but that table is not registered"


def failure_type_mismatch(new_system):
    ds = new_system()
    ds.register("t", of_columns(of_strings("a"), of_ints(1)))
    q = Query("t")
    q.root_node.add_criteria(Criteria.IntCriteria(0, int_range(0, 1)))
    r = ds.execute(q)  # Verifier.java:73-74
    assert isinstance(r, QueryResult.Failure)
    assert r.message == "The column is a string column but the criterion is not a string predicate."
    q = Query("t")
    q.root_node.add_criteria(Criteria.StringCriteria(1, str_equals("a")))
    r = ds.execute(q)  # Verifier.java:78-79
    assert isinstance(r, QueryResult.Failure)
    assert r.message == "The column is an integer column but the criterion is not an integer predicate."


def failure_ordinal_out_of_bounds(new_system):
    ds = new_system()
    ds.register("t", of_columns(of_ints(1, 2)))
    q = Query("t")
    q.root_node.add_criteria(Criteria.IntCriteria(5, int_range(0, 1)))
    r = ds.execute(q)  # Verifier.java:62-65
    assert isinstance(r, QueryResult.Failure)
    assert r.message == "The query ordinal '5' is out of bounds for the table with 1 columns"
    # the reference's off-by-one (`size() < ordinal`): ordinal == width reaches columns().get() -> IndexOutOfBounds
    q = Query("t")
    q.root_node.add_criteria(Criteria.IntCriteria(1, int_range(0, 1)))
    try:
        ds.execute(q)
    except IndexError:
        pass
    else:
        raise failed("expected IndexOutOfBoundsException for ordinal == width (Verifier.java:62,67)")


def failure_boolean_and_association_criteria(new_system):
    ds = new_system()
    t = of_columns(of_ints(1, 2), BooleanColumn([1, 0]))
    u = of_columns(of_ints(7))
    t.associate_to(u, Association.to_one(0), Association.to_none())
    ds.register("t", t)
    q = Query("t")
    q.root_node.add_criteria(Criteria.IntCriteria(1, int_range(0, 1)))
    r = ds.execute(q)  # Verifier.java:82-84
    assert isinstance(r, QueryResult.Failure) and r.message == "Boolean columns are not supported yet."
    q = Query("t")
    q.root_node.add_criteria(Criteria.IntCriteria(2, int_range(0, 1)))
    r = ds.execute(q)  # Verifier.java:85-87
    assert isinstance(r, QueryResult.Failure) and r.message == "Association columns can't be matched on with a scalar criteria."


def failure_child_not_association(new_system):
    ds = new_system()
    ds.register("t", of_columns(of_ints(1, 2)))
    q = Query("t")
    q.root_node.create_child(0)
    r = ds.execute(q)  # Verifier.java:102-104
    assert isinstance(r, QueryResult.Failure)
    assert r.message == ("The column at ordinal 0 is not an association column. It is a "
                         "dgroomes.in_memory.InMemoryColumn$IntegerColumn")


FAILURE_TESTS = [failure_unregistered_table, failure_type_mismatch, failure_ordinal_out_of_bounds,
                 failure_boolean_and_association_criteria, failure_child_not_association]


# ---- the app's two workload queries (Runner.java:230-236, 254-259) against the oracle-derived goldens
def plymouth(new_system, expected, base):
    ds = new_system()
    geo = G.build_tables(1, base=base)
    G.register_geography(ds, geo)
    columns = success_columns(ds.execute(G.plymouth_query()))
    assert len(columns) == 3
    assert columns[0].ints().tolist() == expected["oracle_derived"]["plymouth_zip_codes"]


def north_south_north(new_system, expected, base):
    ds = new_system()
    geo = G.build_tables(1, base=base)
    G.register_geography(ds, geo)
    columns = success_columns(ds.execute(G.north_south_north_query()))
    assert len(columns) == 5
    assert columns[1].strings() == expected["oracle_derived"]["north_south_north_state_names"]


# ---- cases the reference's tests do not reach: None, mixed One/Many reverse columns, several children
def none_and_mixed_reverse(new_system):
    ds = new_system()
    owners = of_columns(of_strings("ann", "bob", "cy", "dee"))
    pets = of_columns(of_strings("rex", "tom", "kit", "jaws", "polly"), of_ints(3, 9, 1, 40, 2))
    # pets -> owner: ann has rex+kit (Many on the reverse side), bob has tom (One), cy none, dee none; jaws/polly unowned
    pets.associate_to(owners, Association.to_one(0), Association.to_one(1), Association.to_one(0),
                      Association.to_none(), Association.to_none())
    ds.register("owners", owners)
    ds.register("pets", pets)
    # owners having a pet older than 2 (through the REVERSE column of pets.2, ordinal 1 on owners)
    q = Query("owners")
    q.root_node.create_child(1).add_criteria(Criteria.IntCriteria(1, int_greater_than(2)))
    assert success_columns(ds.execute(q))[0].strings() == ["ann", "bob"]
    # pets whose owner's name contains "n" (forward to-one with None rows)
    q = Query("pets")
    q.root_node.create_child(2).add_criteria(Criteria.StringCriteria(0, str_contains("n")))
    assert success_columns(ds.execute(q))[0].strings() == ["rex", "kit"]
    # no criteria at all on the child: every owner with at least one pet
    q = Query("owners")
    q.root_node.create_child(1)
    assert success_columns(ds.execute(q))[0].strings() == ["ann", "bob"]


def two_children_on_one_node(new_system):
    ds = new_system()
    people = of_columns(of_strings("p0", "p1", "p2", "p3"))
    towns = of_columns(of_strings("north", "south"))
    jobs = of_columns(of_strings("cook", "smith", "clerk"), of_ints(10, 20, 30))
    people.associate_to(towns, Association.to_one(0), Association.to_one(1), Association.to_one(0), Association.to_one(1))
    people.associate_to(jobs, Association.to_many(0, 1), Association.to_one(2), Association.to_one(2), Association.to_none())
    ds.register("people", people)
    q = Query("people")
    q.root_node.create_child(1).add_criteria(Criteria.StringCriteria(0, str_equals("north")))
    q.root_node.create_child(2).add_criteria(Criteria.IntCriteria(1, int_range(25, 35)))
    assert success_columns(ds.execute(q))[0].strings() == ["p2"]


def boolean_criteria(new_system):
    """predicates over a BooleanColumn (M/InMemoryColumn.java:28-44) -- the
    BooleanColumnFilterable.where(Predicate<Boolean>) the reference declares (DS/ColumnFilterable.java:20-22) and
    never reaches (E/Verifier.java:82-84).  All four truth tables, the AND with other criteria, and a boolean
    criterion on a child pruning its parent."""
    flags = [True, False, False, True, True, False, True]
    for pred, want in [(lambda b: b, [0, 3, 4, 6]), (lambda b: not b, [1, 2, 5]), (lambda b: True, list(range(7))),
                       (lambda b: False, [])]:
        ds = new_system()
        ds.register("t", of_columns(of_ints(*range(7)), BooleanColumn(flags)))
        q = Query("t")
        q.root_node.add_criteria(Criteria.BooleanCriteria(1, pred))
        assert success_columns(ds.execute(q))[0].ints().tolist() == want
    ds = new_system()
    ds.register("t", of_columns(of_ints(*range(7)), BooleanColumn(flags), of_strings("a", "b", "c", "a", "b", "a", "a")))
    q = Query("t")
    (q.root_node.add_criteria(Criteria.BooleanCriteria(1, lambda b: b)).add_criteria(Criteria.IntCriteria(0, int_range(1, 6)))
        .add_criteria(Criteria.StringCriteria(2, str_equals("a"))))
    assert success_columns(ds.execute(q))[0].ints().tolist() == [3, 6]
    # a child's boolean criterion prunes the parent, through a to-one and through a to-many column
    ds = new_system()
    pets = of_columns(of_strings("p0", "p1", "p2", "p3"), BooleanColumn([False, True, False, False]))
    owners = of_columns(of_strings("ann", "bob", "cy"))
    owners.associate_to(pets, Association.to_many(0, 2), Association.to_one(1), Association.to_none())
    pets.associate_to(owners, Association.to_one(0), Association.to_none(), Association.to_one(2), Association.to_one(0))
    ds.register("owners", owners)
    ds.register("pets", pets)
    q = Query("owners")
    q.root_node.create_child(1).add_criteria(Criteria.BooleanCriteria(1, lambda b: b))
    assert success_columns(ds.execute(q))[0].strings() == ["bob"]
    q = Query("pets")
    q.root_node.add_criteria(Criteria.BooleanCriteria(1, lambda b: not b))
    q.root_node.create_child(3).add_criteria(Criteria.StringCriteria(0, str_equals("ann")))
    assert success_columns(ds.execute(q))[0].strings() == ["p0", "p3"]
    # a boolean criterion on a column of another kind is a type mismatch, like the reference's other mismatches
    q = Query("pets")
    q.root_node.add_criteria(Criteria.BooleanCriteria(0, lambda b: b))
    r = ds.execute(q)
    assert isinstance(r, QueryResult.Failure) and r.message == "The column is a string column but the criterion is not a string predicate."


EXTRA_TESTS = [none_and_mixed_reverse, two_children_on_one_node, boolean_criteria]
