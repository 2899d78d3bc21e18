"""GPU suite (>= 2 GPUs): association hops that LEAVE the shard (SURVEY.md 8e "not supported" in round 1, 8f4).

``DataSystemColqGroup`` takes whole application tables, splits every one of them into row ranges over the GPUs and keeps
the association keys GLOBAL, so any hop may cross shards in either direction -- the general case of
ExecutionContext.Node.filterParent (E/ExecutionContext.java:100-122) and InMemoryTable.associateTo
(M/InMemoryTable.java:44-90).  Everything is compared with the unsharded CPU oracle: the reference's own QueryTest cases,
the failure cases, the headline queries with all three tables sharded (the 51-row states table lives on rank 0 only), and
the randomly generated schemas / query trees of test_gpu_fuzz.
"""
import numpy as np
import pytest

import tck
from colq import QueryResult, geography as G
from oracle_system import OracleDataSystem
from fuzz_cases import make_case

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return min(torch.cuda.device_count(), 8)


@pytest.fixture()
def group_factory():
    if n_gpus() < 2:
        pytest.skip("needs at least 2 GPUs")
    from colq.local_group import DataSystemColqGroup
    made = []

    def new(**kw):
        for ds in made:      # one group at a time owns the GPUs' peer mappings
            ds.close()
        made.clear()
        ds = DataSystemColqGroup(list(range(n_gpus())), **kw)
        made.append(ds)
        return ds

    yield new
    for ds in made:
        ds.close()


@pytest.mark.parametrize("case", tck.REFERENCE_TESTS + tck.FAILURE_TESTS + tck.EXTRA_TESTS, ids=lambda f: f.__name__)
@pytest.mark.parametrize("sharded", [True, False], ids=["all-sharded", "all-replicated"])
def test_tck_on_all_gpus(group_factory, case, sharded):
    case(lambda: group_factory(default_sharded=sharded))


def test_headline_queries_every_table_sharded(group_factory, expected, base_geography):
    tck.plymouth(lambda: group_factory(), expected, base_geography)
    tck.north_south_north(lambda: group_factory(), expected, base_geography)


@pytest.mark.parametrize("U", [7, 40])
@pytest.mark.parametrize("lazy", [True, False])
def test_plymouth_universes_cross_shard(group_factory, base_geography, U, lazy):
    """The zip and city tables are split by plain row ranges (NOT by universe), so zip -> city keys leave the shard at every
    range boundary, and cities -> states / states <-> states always do (all states live on rank 0)."""
    geo = G.build_tables(U, base=base_geography)
    oracle = OracleDataSystem()
    G.register_geography(oracle, geo)
    ds = group_factory(lazy_fk=lazy)
    G.register_geography(ds, geo)
    for make in (G.plymouth_query, G.north_south_north_query):
        want = oracle.execute(make())
        got = ds.execute(make())
        assert isinstance(got, QueryResult.Success), getattr(got, "message", got)
        assert np.array_equal(ds.last_indices, oracle.last_indices)
        assert got.result_set.size() == want.result_set.size()
    names = [n for n, *_ in ds.last_queries[0].profile()]
    assert names, names
    oracle.close()


def test_mixed_placement(group_factory, base_geography):
    """zips and cities sharded, states replicated: cross-shard zip -> city pull plus the small-mask exchange for the
    sharded -> replicated push, in one plan."""
    geo = G.build_tables(9, base=base_geography)
    oracle = OracleDataSystem()
    G.register_geography(oracle, geo)
    oracle.execute(G.plymouth_query())
    ds = group_factory()
    ds.register("states", geo.states, sharded=False)
    ds.register("cities", geo.cities, sharded=True)
    ds.register("zips", geo.zips, sharded=True)
    got = ds.execute(G.plymouth_query())
    assert isinstance(got, QueryResult.Success), getattr(got, "message", got)
    assert np.array_equal(ds.last_indices, oracle.last_indices)
    names = [n for n, *_ in ds.last_queries[0].profile()]
    assert "peer_bits_allgather" in names and any("publish" in n for n in names), names
    oracle.close()


@pytest.mark.parametrize("seed", range(16))
def test_random_schemas_every_table_sharded(group_factory, seed):
    build, queries = make_case(seed)
    oracle = OracleDataSystem()
    build(oracle)
    ds = group_factory(lazy_fk=bool(seed % 2 == 0))
    build(ds)
    for q in queries:
        want = oracle.execute(q())
        got = ds.execute(q())
        assert isinstance(want, QueryResult.Success) and isinstance(got, QueryResult.Success), (getattr(want, "message", None), getattr(got, "message", None))
        assert np.array_equal(ds.last_indices, oracle.last_indices), "row index set differs from the oracle"
    oracle.close()
