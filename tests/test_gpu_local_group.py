"""GPU suite (>= 2 GPUs): ONE process drives all GPUs through colq_comm_init_local -- the multi-GPU mode a single-JVM host
(E/DataSystemSerialIndices.java:14-22) can use.  The sharded result must equal the unsharded oracle, for the exact and the
perturbed workload (the state mask exists on the last rank only before the exchange), including the re-run of the whole
group when the result blocks start out too small."""
import numpy as np
import pytest

from colq import _ffi, geography as G
from colq.engine import DataSystemColq
from oracle_system import OracleDataSystem


def _sharded_systems(group, U, perturbed, options=None):
    out = []
    for rank, ctx in enumerate(group.ctxs):
        geo = G.build_tables(U, n_ranks=group.n_ranks, rank=rank, rename_plymouth_except_last_rank=perturbed)
        ds = DataSystemColq(context=ctx, options=options or {})
        G.register_geography(ds, geo, sharded=True)
        ds._sync_tables()
        out.append(ds)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("perturbed", [False, True])
def test_one_process_drives_all_gpus(perturbed):
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    from colq.local_group import ColqLocalGroup
    U = 4 * n + 1
    oracle = OracleDataSystem()
    G.register_geography(oracle, G.build_tables(U))
    oracle.execute(G.plymouth_query())
    want = oracle.last_indices.copy()
    oracle.execute(G.north_south_north_query())
    want_nsn = oracle.last_indices.copy()
    oracle.close()

    group = ColqLocalGroup(list(range(n)))
    try:
        for options in ({}, {_ffi.OPT_ROOT_FUSED: 0}, {_ffi.OPT_FUSED_GATHER: 0}, {_ffi.OPT_TAIL_PUBLISH: 0, _ffi.OPT_ROOT_FUSED: 0}):
            systems = _sharded_systems(group, U, perturbed, options)
            queries = []
            for ds in systems:
                cq, why = ds._translate(G.plymouth_query())
                assert cq is not None, why
                queries.append(cq)
            # index_capacity 64 is far too small for 31 * U rows: the whole group re-runs with larger blocks
            results = group.execute(queries, index_capacity=64)
            for r in results:
                assert r.count == want.shape[0] and np.array_equal(r.indices, want), (options, perturbed)
            # steady state: several executions back to back, then one fetch
            for _ in range(3):
                group.execute_async(queries)
            for r in group.fetch(queries):
                assert np.array_equal(r.indices, want)
            for cq in queries:
                cq.close()
            # the replicated table alone: no exchange, same answer on every rank
            queries = [ds._translate(G.north_south_north_query())[0] for ds in systems]
            for r in group.execute(queries):
                assert np.array_equal(r.indices, want_nsn)
            for cq in queries:
                cq.close()
            for ds, ctx in zip(systems, group.ctxs):
                for h in set(ds._handles.values()):
                    ctx.table_destroy(h)
    finally:
        group.close()
