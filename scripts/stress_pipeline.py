"""Stress of COLQ_OPT_PIPELINE: random bursts of back-to-back executions of the Plymouth query, randomly
interleaved with a second query of the same context, option flips and profiled steps; every fetched result is compared
with the oracle.  One GPU: `python scripts/stress_pipeline.py`; N GPUs (sharded tables, mask exchange + gather in the
pipelined plan): `torchrun --nproc-per-node N ... scripts/stress_pipeline.py`.  Prints STRESS_OK <iterations>."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "java-columnar-query-engine_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))

from colq import _ffi, geography as G  # noqa: E402
from colq.engine import ColqContext, DataSystemColq  # noqa: E402
from oracle_system import OracleDataSystem  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", rank))
    iters = int(os.environ.get("ITERS", "300"))
    torch.cuda.set_device(local)
    ctx = ColqContext(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), world, rank)
    total = 0
    for U in (4 * world + 1, 40 * world + 3):
        oracle = OracleDataSystem()
        G.register_geography(oracle, G.build_tables(U))
        oracle.execute(G.plymouth_query())
        want = oracle.last_indices.copy()
        oracle.execute(G.north_south_north_query())
        want_nsn = oracle.last_indices.copy()
        oracle.close()
        geo = G.build_tables(U, n_ranks=world, rank=rank, rename_plymouth_except_last_rank=world > 1)
        ds = DataSystemColq(context=ctx)
        ds._tables.clear()
        G.register_geography(ds, geo, sharded=world > 1)
        ds._sync_tables()
        cq, why = ds._translate(G.plymouth_query())
        assert cq is not None, why
        other, why = ds._translate(G.north_south_north_query())
        assert other is not None, why
        rng = np.random.default_rng(1234 + U)     # the same decisions on every rank
        cap = want.shape[0] + 8
        res = cq.execute(want_indices=True, index_capacity=cap)
        assert np.array_equal(res.indices, want)
        for it in range(iters):
            burst = int(rng.integers(1, 7))
            flip = rng.random()
            if flip < 0.08:
                cq.set_option(_ffi.OPT_PIPELINE, int(rng.integers(0, 2)))
            elif flip < 0.14:
                cq.set_option(_ffi.OPT_PROFILE, int(rng.choice([0, 1, 2])))
            for _ in range(burst):
                cq.execute_async()
                if rng.random() < 0.25:
                    other.execute_async()
            if rng.random() < 0.3:
                r2 = other.execute(want_indices=True)
                assert np.array_equal(r2.indices, want_nsn), (rank, U, it, "north-south-north")
            res = cq.fetch(want_indices=True, index_capacity=cap)
            assert res.count == want.shape[0] and np.array_equal(res.indices, want), (rank, U, it, res.count, want.shape[0])
            total += burst
        cq.set_option(_ffi.OPT_PROFILE, 0)
        cq.profile_hot()
        cq.close()
        other.close()
        for h in set(ds._handles.values()):
            ctx.table_destroy(h)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if rank == 0:
        print(f"STRESS_OK world={world} executions={total}")
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
