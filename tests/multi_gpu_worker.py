"""Worker for tests/test_gpu_multi.py: run under torchrun with one rank per GPU.

Checks, through the reference-facing Python mirror over the C ABI, that the sharded execution (universe ranges per
rank, replicated states, mask exchange, final gather) returns exactly the unsharded oracle result -- for the exact-copy
workload, for the perturbed workload where only the last rank holds PLYMOUTH rows, with peer-memory and NCCL exchanges.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "java-columnar-query-engine_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))

from colq import _ffi, geography as G  # noqa: E402
from colq.engine import ColqContext, DataSystemColq  # noqa: E402
from oracle_system import OracleDataSystem  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ColqContext(local)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), world, rank)

    U = 4 * world + 1   # uneven split on purpose
    checked = 0
    for perturbed in (False, True):
        # the unsharded expectation from the oracle (exact copies: the perturbation moves no result row)
        oracle = OracleDataSystem()
        G.register_geography(oracle, G.build_tables(U))
        oracle.execute(G.plymouth_query())
        want = oracle.last_indices.copy()
        oracle.execute(G.north_south_north_query())
        want_nsn = oracle.last_indices.copy()
        oracle.close()

        geo = G.build_tables(U, n_ranks=world, rank=rank, rename_plymouth_except_last_rank=perturbed)
        # (peer exchange, lazy FK, deferred chains, fused compaction, gather fused into the index writer, tail publish, fused root)
        for variant in ((1, True, 1, 1, 1, 1, 1), (1, True, 1, 1, 1, 0, 1), (1, True, 1, 1, 0, 1, 1),
                                                              (1, True, 1, 1, 0, 1, 0), (1, True, 1, 1, 0, 0, 0), (1, True, 1, 2, 0, 1, 1),
                                                              (1, True, 1, 1, 1, 1, 0), (1, True, 0, 1, 1, 0, 1), (1, False, 1, 1, 1, 1, 1),
                                                              (1, False, 1, 1, 0, 1, 0), (1, True, 1, 0, 0, 1, 1), (0, True, 1, 1, 0, 1, 1),
                                                              (0, True, 1, 1, 0, 1, 0), (0, False, 1, 2, 0, 1, 1), (1, True, 1, 1, 1, 1, 1, 0),
                        (1, True, 1, 1, 1, 1, 2), (1, True, 1, 1, 0, 1, 2), (0, True, 1, 1, 0, 1, 2), (1, True, 1, 1, 1, 0, 2, 0)):
            peer, lazy, defer, fused, fgather, tail, rootf = variant[:7]
            lazy_wait = variant[7] if len(variant) > 7 else 1   # COLQ_OPT_LAZY_GATHER_WAIT
            if True:
                ds = DataSystemColq(context=ctx, lazy_fk=lazy, options={_ffi.OPT_PEER_EXCHANGE: peer, _ffi.OPT_DEFER_CHAINS: defer,
                                                                        _ffi.OPT_FUSED_COMPACT: fused, _ffi.OPT_FUSED_GATHER: fgather,
                                                                        _ffi.OPT_TAIL_PUBLISH: tail, _ffi.OPT_ROOT_FUSED: rootf,
                                                                        _ffi.OPT_LAZY_GATHER_WAIT: lazy_wait})
                ds._tables.clear()
                G.register_geography(ds, geo, sharded=True)
                ds._sync_tables()
                cq, why = ds._translate(G.plymouth_query())
                assert cq is not None, why
                res = cq.execute(want_indices=True, index_capacity=64)   # small on purpose: exercises the regrow path
                assert res.count == want.shape[0], (rank, perturbed, peer, lazy, res.count, want.shape[0])
                assert np.array_equal(res.indices, want), (rank, perturbed, peer, lazy)
                # a second execution of the same query (other slot parity), count only, then the indices through fetch
                cq.execute_async()
                res = cq.fetch(want_indices=True, index_capacity=want.shape[0] + 8)
                assert np.array_equal(res.indices, want), (rank, perturbed, "second execution")
                # a burst of back-to-back executions (COLQ_OPT_PIPELINE: no memset, the string scan starts while the previous
                # execution's root kernel is still waiting for the peers)
                for _ in range(6):
                    cq.execute_async()
                res = cq.fetch(want_indices=True, index_capacity=want.shape[0] + 8)
                assert res.count == want.shape[0] and np.array_equal(res.indices, want), (rank, perturbed, "pipelined burst")
                # the public call: this rank's rows of the result table (the gathered list holds ALL ranks' rows)
                got = ds.execute(G.plymouth_query())
                mine = want[(want >= geo.zip_row_base) & (want < geo.zip_row_base + geo.zips.size())]
                assert got.result_set.size() == mine.shape[0], (rank, got.result_set.size(), mine.shape[0])
                assert np.array_equal(got.result_set.columns()[0].ints(), geo.zips.columns()[0].ints()[mine - geo.zip_row_base])
                ds.last_query.close()
                ds.last_query = None
                names = [n for n, *_ in cq.profile()]
                root_fused = rootf and fused == 1 and lazy and defer   # (eager plans end in a scan with the gathers inside)
                assert any(n.startswith("root_fused") for n in names) == bool(root_fused and rootf == 1), names
                assert any(n.startswith("root_finish") for n in names) == bool(root_fused and rootf == 2), names
                if peer and os.environ.get("COLQ_PEER", "1") != "0":
                    pub = [i for i, n in enumerate(names) if n.endswith("publish")]   # own launch, or the scan's last CTA
                    assert len(pub) == 1, names
                    assert (names[pub[0]] == "peer_mask_publish") == (not tail), names
                    if root_fused:       # the COLLECT and the 51-row adjacency hop run inside the root's launch
                        assert any("+collect+csr" in n for n in names) and "peer_mask_collect+csr_pull" not in names, names
                    else:
                        assert "peer_mask_collect+csr_pull" in names, names
                    if fused == 1 and fgather:    # the final gather rides on the kernel that writes the indices
                        assert any(n.endswith("+gather") for n in names) and "peer_gather_indices" not in names, names
                    else:
                        assert "peer_gather_indices" in names, names
                    if not root_fused and lazy and defer and fused:   # the root's scan sits between the two halves of the mask exchange
                        assert pub[0] < names.index("scan_rows<1,0,lazy>") < names.index("peer_mask_collect+csr_pull"), names
                else:
                    assert "allgather_or_mask" in names and "allgather_indices" in names, names
                cq.close()
                # a query on the replicated table alone: no collective, same answer on every rank
                cq, _ = ds._translate(G.north_south_north_query())
                res = cq.execute(want_indices=True)
                assert np.array_equal(res.indices, want_nsn)
                cq.close()
                # drop this engine's tables so the next variant re-registers under the same names
                for h in set(ds._handles.values()):
                    ctx.table_destroy(h)
                checked += 1
    # ---- cross-shard hops over the CUDA-IPC mapped heap: cities and zips split by plain row ranges (not by universe), the
    #      zip -> city keys GLOBAL, so the pull hop leaves the shard; states replicated (small-mask exchange in the same plan)
    from colq.device_data import plymouth_colq_query, register_cross_shard_geography
    geo = G.build_tables(U)
    oracle = OracleDataSystem()
    G.register_geography(oracle, geo)
    oracle.execute(G.plymouth_query())
    want = oracle.last_indices.copy()
    oracle.close()
    register_cross_shard_geography(ctx, geo, world, rank)
    for lazy in (True, False):
        q = plymouth_colq_query(ctx, lazy_fk=lazy)
        for _ in range(3):   # both heap parities
            res = q.execute(want_indices=True, index_capacity=want.shape[0] + 8)
            assert res.count == want.shape[0] and np.array_equal(res.indices, want), (rank, "cross-shard", lazy)
        names = [n for n, *_ in q.profile()]
        assert "peer_bits_allgather" in names, names
        q.close()
    checked += 2
    dist.barrier()
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} variants={checked}")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
