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


# ---------------------------------------------------------------------------------------------------------------------
# Cross-shard hops: the protocol libcolq runs over NVLink peer memory, restated over gloo.  Tables
# are split by plain row ranges with 64-row aligned bounds, association keys stay GLOBAL:
#   pull  (parent holds the key): all-gather of the child's bits into a global bitmap, then a local bit test per key;
#   push  (child holds the key):  every rank sets bits in its own GLOBAL-sized reach bitmap, then an OR-reduce whose
#                                 slice [bounds[r], bounds[r+1]) is what rank r keeps.
def _cross_shard_worker(rank, world, port, seed, out_dir):
    for p in (ROOT / "java-columnar-query-engine_b200", ROOT / "oracle"):
        sys.path.insert(0, str(p))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from colq.local_group import even_partition
    rng = np.random.default_rng(seed)          # every rank draws the same global tables
    na, nb = 1000, 300
    a_val = rng.integers(0, 100, size=na)
    b_val = rng.integers(0, 100, size=nb)
    a_to_b = rng.integers(-1, nb, size=na)     # global keys, -1 = None
    ba, bb = even_partition(na, world), even_partition(nb, world)
    a0, a1, b0, b1 = int(ba[rank]), int(ba[rank + 1]), int(bb[rank]), int(bb[rank + 1])

    def allgather_bits(local_bits, bounds, n_global):
        m = max(int(bounds[r + 1] - bounds[r]) for r in range(world))   # ragged shards: pad to the largest
        padded = [torch.zeros(m, dtype=torch.uint8) for _ in range(world)]
        mine = torch.zeros(m, dtype=torch.uint8)
        mine[: local_bits.shape[0]] = torch.from_numpy(local_bits.astype(np.uint8))
        dist.all_gather(padded, mine)
        out = torch.cat([padded[r][: int(bounds[r + 1] - bounds[r])] for r in range(world)]).numpy().astype(bool)
        assert out.shape[0] == n_global
        return out

    # query 1 (pull): A rows with a_val < 50 whose B row has b_val < 30
    b_bits_local = b_val[b0:b1] < 30
    b_bits_global = allgather_bits(b_bits_local, bb, nb)
    keys = a_to_b[a0:a1]
    a_match = (a_val[a0:a1] < 50) & (keys >= 0) & b_bits_global[np.clip(keys, 0, nb - 1)]
    got_pull = np.flatnonzero(a_match) + a0
    # query 2 (push): B rows with b_val >= 10 that some A row with a_val >= 90 points at
    reach = np.zeros(nb, dtype=np.int32)
    src = keys[(a_val[a0:a1] >= 90) & (keys >= 0)]
    reach[src] = 1
    t = torch.from_numpy(reach)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)            # OR; rank r keeps its slice
    got_push = np.flatnonzero((t.numpy()[b0:b1] > 0) & (b_val[b0:b1] >= 10)) + b0
    np.save(Path(out_dir) / f"pull_{rank}.npy", got_pull)
    np.save(Path(out_dir) / f"push_{rank}.npy", got_push)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("seed", [3, 4])
def test_two_rank_cross_shard_protocol_matches_unsharded_oracle(tmp_path, seed):
    from colq import Criteria, InMemoryTable, Query, int_range
    from colq.in_memory import IntegerColumn
    from oracle_system import OracleDataSystem
    world = 2
    port = 29600 + (os.getpid() % 2000) + seed
    mp.spawn(_cross_shard_worker, args=(world, port, seed, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(seed)
    na, nb = 1000, 300
    a_val = rng.integers(0, 100, size=na).astype(np.int32)
    b_val = rng.integers(0, 100, size=nb).astype(np.int32)
    a_to_b = rng.integers(-1, nb, size=na).astype(np.int32)
    A = InMemoryTable.of_columns(IntegerColumn(a_val))
    B = InMemoryTable.of_columns(IntegerColumn(b_val))
    A.associate_to(B, fk=a_to_b)      # A.1 -> B ; B.1 <- A
    ds = OracleDataSystem()
    ds.register("A", A)
    ds.register("B", B)
    q = Query("A")
    q.root_node.add_criteria(Criteria.IntCriteria(0, int_range(0, 49)))
    q.root_node.create_child(1).add_criteria(Criteria.IntCriteria(0, int_range(0, 29)))
    ds.execute(q)
    got = np.concatenate([np.load(tmp_path / f"pull_{r}.npy") for r in range(world)])
    assert np.array_equal(got, ds.last_indices.astype(np.int64))
    q = Query("B")
    q.root_node.add_criteria(Criteria.IntCriteria(0, int_range(10, 1000)))
    q.root_node.create_child(1).add_criteria(Criteria.IntCriteria(0, int_range(90, 1000)))
    ds.execute(q)
    got = np.concatenate([np.load(tmp_path / f"push_{r}.npy") for r in range(world)])
    assert np.array_equal(got, ds.last_indices.astype(np.int64))
    ds.close()
