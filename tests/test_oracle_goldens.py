"""CPU suite: pins the oracle (oracle/colq_oracle.c) to every known answer the reference's tests hold for the hot path
 and freezes the oracle-derived answers of the two headline queries."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

import tck
from colq import geography as G
from oracle_system import OracleDataSystem

ROOT = Path(__file__).resolve().parent.parent


def new_oracle():
    return OracleDataSystem()


@pytest.mark.parametrize("case", tck.REFERENCE_TESTS, ids=lambda f: f.__name__)
def test_reference_query_tests(case):
    case(new_oracle)


@pytest.mark.parametrize("case", tck.FAILURE_TESTS + tck.EXTRA_TESTS, ids=lambda f: f.__name__)
def test_failure_and_extra_cases(case):
    case(new_oracle)


def test_loader_cardinalities(expected, base_geography):
    """TheTest.loadObjectGraph (app/src/test/java/dgroomes/TheTest.java:22-26)."""
    want = expected["reference_pinned"]["loader_cardinalities"]
    assert base_geography["zip_code"].shape[0] == want["zips"]
    assert base_geography["city_state"].shape[0] == want["cities"]
    assert base_geography["state_code_offsets"].shape[0] - 1 == want["states"]
    assert base_geography["adj_targets"].shape[0] == 219
    meta = json.loads((ROOT / "tests/golden/geography_meta.json").read_text())
    for k, v in base_geography.items():
        assert hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest() == meta["sha256"][k], k


def test_plymouth_and_nsn_goldens(expected, base_geography):
    tck.plymouth(new_oracle, expected, base_geography)
    tck.north_south_north(new_oracle, expected, base_geography)


def test_plymouth_intermediates(expected, base_geography):
    ds = new_oracle()
    geo = G.build_tables(1, base=base_geography)
    G.register_geography(ds, geo)
    ds.execute(G.plymouth_query())
    want = expected["oracle_derived"]["plymouth_node_cardinalities"]
    assert ds.node_cardinalities() == [want["zips"], want["cities"], want["states"], want["states_adjacent"],
                                       want["cities_plymouth"]]
    pop = base_geography["zip_pop"]
    assert int(((pop >= 10_000) & (pop < 10_100)).sum()) == expected["oracle_derived"]["zips_with_population_10000_10099"]
    i = int(np.argmax(pop))  # Runner.java:199-221
    mp = expected["oracle_derived"]["max_population_zip"]
    assert (int(base_geography["zip_code"][i]), int(pop[i])) == (mp["code"], mp["population"])


def test_universe_replication_and_threads(base_geography):
    """U exact copies => matches are {u * 29353 + r}; the OpenMP variant equals the serial engine."""
    U = 7
    geo = G.build_tables(U, base=base_geography)
    one = new_oracle()
    G.register_geography(one, G.build_tables(1, base=base_geography))
    one.execute(G.plymouth_query())
    base_rows = one.last_indices.astype(np.int64)
    want = (np.arange(U)[:, None] * G.N_ZIPS + base_rows[None, :]).reshape(-1)
    for threads in (1, 4):
        ds = OracleDataSystem(n_threads=threads)
        G.register_geography(ds, geo)
        r = ds.execute(G.plymouth_query())
        assert np.array_equal(ds.last_indices, want)
        assert r.result_set.size() == 31 * U


def test_compare_to_is_utf16_order():
    """String.compareTo orders by UTF-16 code units: a supplementary code point sorts below U+E000..U+FFFF."""
    from colq import Criteria, Query, of_columns, of_strings, str_compare_lt
    ds = new_oracle()
    strings = ["\U0001F600", "ﬁ", "z", "", ""]
    ds.register("s", of_columns(of_strings(*strings)))
    q = Query("s")
    q.root_node.add_criteria(Criteria.StringCriteria(0, str_compare_lt("")))
    got = ds.execute(q).result_set.columns()[0].strings()
    pred = str_compare_lt("")
    assert got == [s for s in strings if pred(s)] == ["\U0001F600", "z", ""]


def test_compare_to_supplementary_planes_and_mixed_lengths():
    """two different supplementary-plane lead bytes (0xF0 vs 0xF4) must not tie.  The oracle decodes to
    UTF-16 code units; the expectation is Python's own UTF-16-BE byte comparison (colq.data_system._java_compare_to)."""
    from colq import Criteria, Query, of_columns, of_strings
    from colq.data_system import StringPredicate
    strings = ["\U00010000", "\U0010FFFF", "\U0001F600", "\U000F0000a", "\uE000", "\uFFFD", "\uD7FF", "a\U00010400", "a\U0010F400",
               "a", "", "\U0001F600\U0001F600", "é", "z\uFB01"]
    for needle in ("\U0001F600", "\U0010FFFF", "\U00010000", "\uE000", "a\U00010400", "", "\U000F0000"):
        for op in (2, 3, 4, 5):
            ds = new_oracle()
            ds.register("s", of_columns(of_strings(*strings)))
            q = Query("s")
            pred = StringPredicate(op, needle)
            q.root_node.add_criteria(Criteria.StringCriteria(0, pred))
            got = ds.execute(q).result_set.columns()[0].strings()
            assert got == [s for s in strings if pred(s)], (op, needle)
