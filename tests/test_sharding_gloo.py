"""CPU suite: the N>1 protocol on world_size-2 gloo.

Each rank holds a contiguous universe range of zips/cities and a replica of the states table.  The only exchange is the
OR-reduction of the replicated states mask after the sharded city scan; everything else is rank-local.  The oracle runs
the per-rank pieces here (test infrastructure); on GPUs the same protocol is executed by libcolq.so with NCCL.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, U, perturbed, out_dir):
    for p in (ROOT / "java-columnar-query-engine_b200", ROOT / "oracle"):
        sys.path.insert(0, str(p))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from colq import Criteria, Query, geography as G, int_half_open, str_equals
    from oracle_system import OracleDataSystem

    geo = G.build_tables(U, n_ranks=world, rank=rank, rename_plymouth_except_last_rank=perturbed)
    ds = OracleDataSystem()
    G.register_geography(ds, geo)

    # hop 1 (sharded -> replicated): states that have a city named PLYMOUTH in THIS rank's universes
    q1 = Query("states")
    q1.root_node.create_child(2).add_criteria(Criteria.StringCriteria(0, str_equals("PLYMOUTH")))
    ds.execute(q1)
    mask = np.zeros(1, dtype=np.int64)
    for s in ds.last_indices:
        mask[0] |= 1 << int(s)
    local_mask = int(mask[0])
    gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mask))          # the collective: all-gather + local OR
    full = 0
    for g in gathered:
        full |= int(g[0])

    # hops 2..4 are local given the reduced mask: adjacency on the replica, then the fused zip scan
    base = G.load_base()
    adj_off, adj_tgt = base["adj_offsets"], base["adj_targets"]
    states = [s for s in range(G.N_STATES) if any((full >> int(t)) & 1 for t in adj_tgt[adj_off[s]:adj_off[s + 1]])]
    ok_state = np.zeros(G.N_STATES, dtype=bool)
    ok_state[states] = True
    pop = geo.zips.columns()[1].ints()
    city = geo.zips.columns()[2].fk()
    st = geo.cities.columns()[1].fk()
    local = np.flatnonzero((pop >= 10_000) & (pop < 10_100) & ok_state[st[city]]) + geo.zip_row_base

    # final gather of matched indices to rank 0
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64))
    if rank == 0:
        parts = [torch.from_numpy(local.astype(np.int64))]
        for r in range(1, world):
            buf = torch.zeros(int(counts[r][0]), dtype=torch.int64)
            if buf.numel():
                dist.recv(buf, src=r)
            parts.append(buf)
        np.save(Path(out_dir) / "gathered.npy", torch.cat(parts).numpy())
    elif local.shape[0]:
        dist.send(torch.from_numpy(local.astype(np.int64)), dst=0)
    np.save(Path(out_dir) / f"mask_{rank}.npy", np.array([local_mask, full], dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("perturbed", [False, True])
def test_two_rank_protocol_matches_unsharded_oracle(tmp_path, perturbed):
    from colq import geography as G
    from oracle_system import OracleDataSystem
    U, world = 5, 2
    port = 29500 + (os.getpid() % 2000) + (1 if perturbed else 0)
    mp.spawn(_worker, args=(world, port, U, perturbed, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")

    # unsharded reference run over the concatenation of both ranks' universes
    if perturbed:
        # ranks own universes [0,3) and [3,5): only the last rank kept the PLYMOUTH rows
        geo0 = G.build_tables(U, n_ranks=world, rank=0, rename_plymouth_except_last_rank=True)
        geo1 = G.build_tables(U, n_ranks=world, rank=1, rename_plymouth_except_last_rank=True)
        m0, m1 = np.load(tmp_path / "mask_0.npy"), np.load(tmp_path / "mask_1.npy")
        assert m0[0] == 0 and m1[0] != 0 and m0[1] == m1[1] == m1[0]   # the mask existed on one rank only
        assert geo0.n_universes == 3 and geo1.n_universes == 2
    ds = OracleDataSystem()
    G.register_geography(ds, G.build_tables(U))
    ds.execute(G.plymouth_query())
    assert np.array_equal(got, ds.last_indices.astype(np.int64))   # exact copies: the perturbation moves no result row
    assert got.shape[0] == 31 * U and np.all(np.diff(got) > 0)
