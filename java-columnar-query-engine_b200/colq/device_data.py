"""Device-resident workload builders: the synthetic "parallel universes" tables generated straight in HBM.

PyTorch is used here only as the device-memory allocator (plumbing): tensors are created on the GPU, padded so the
TMA path can read whole 16-byte lines, and handed to libcolq.so by pointer (``colq_col_*_device``).  No query work is
done by torch.  Layout and replication rule are those of ``geography.replicate_columns``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import _ffi
from .engine import ColqContext
from .geography import N_CITIES, N_STATES, N_ZIPS, load_base, universe_range


@dataclass
class DeviceGeography:
    zips: int
    cities: int
    states: int
    n_universes: int
    u0: int
    n_zip_rows: int
    n_city_rows: int
    name_bytes: int
    tensors: Dict[str, torch.Tensor]

    def algorithmic_bytes_plymouth(self, matches: int) -> int:
        """every touched column read once + output indices."""
        return (4 * self.n_zip_rows * 2 + 4 * (self.n_city_rows + 1) + self.name_bytes + 4 * self.n_city_rows + 4 * matches)


def _padded(t: torch.Tensor, pad_elems: int) -> torch.Tensor:
    out = torch.zeros(t.numel() + pad_elems, dtype=t.dtype, device=t.device)
    out[: t.numel()] = t
    return out


def build_geography_on_device(ctx: ColqContext, n_universes: int, n_ranks: int = 1, rank: int = 0,
                              base: Optional[Dict[str, np.ndarray]] = None, device: Optional[torch.device] = None,
                              sharded: bool = False) -> DeviceGeography:
    """Runner.java:89-196 for this rank's universe range, generated in HBM and registered by pointer."""
    base = base or load_base()
    device = device or torch.device("cuda", ctx.device)
    u0, u1 = universe_range(n_universes, n_ranks, rank)
    U = u1 - u0
    nb = int(base["city_name_bytes"].shape[0])
    assert U * nb < 2 ** 32 - 64, "city-name bytes exceed the uint32 offset range; use more ranks"

    def dev(a, dtype):
        return torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype)

    u = torch.arange(U, device=device, dtype=torch.int64)
    t: Dict[str, torch.Tensor] = {}
    t["zip_code"] = _padded(dev(base["zip_code"], torch.int32).repeat(U), 16)
    t["zip_pop"] = _padded(dev(base["zip_pop"], torch.int32).repeat(U), 16)
    zc = dev(base["zip_city"], torch.int64)[None, :] + (u * N_CITIES)[:, None]
    t["zip_city"] = _padded(zc.reshape(-1).to(torch.int32), 16)
    del zc
    off = dev(base["city_name_offsets"][:-1].astype(np.int64), torch.int64)[None, :] + (u * nb)[:, None]
    off = torch.cat([off.reshape(-1), torch.tensor([U * nb], device=device, dtype=torch.int64)])
    # uint32 offsets stored in an int32 tensor (same bits); values >= 2^31 wrap, which is exactly the uint32 encoding
    t["city_name_offsets"] = _padded(((off + 2 ** 31) % 2 ** 32 - 2 ** 31).to(torch.int32), 16)
    del off
    t["city_name_bytes"] = _padded(dev(base["city_name_bytes"], torch.uint8).repeat(U), 64)
    t["city_state"] = _padded(dev(base["city_state"], torch.int32).repeat(U), 16)
    torch.cuda.synchronize(device)

    nz, nc = U * N_ZIPS, U * N_CITIES
    place = _ffi.SHARDED if sharded else _ffi.REPLICATED
    states = ctx.table_create(N_STATES, _ffi.REPLICATED, 0)
    cities = ctx.table_create(nc, place, u0 * N_CITIES)
    zips = ctx.table_create(nz, place, u0 * N_ZIPS)
    ctx.col_str(states, 0, base["state_code_offsets"], base["state_code_bytes"])
    ctx.col_str(states, 1, base["state_name_offsets"], base["state_name_bytes"])
    tt = t["city_name_offsets"], t["city_name_bytes"]
    ctx.col_str_device(cities, 0, tt[0].data_ptr(), tt[0].numel() * 4, tt[1].data_ptr(), tt[1].numel(), nc, U * nb, keepalive=tt)
    ctx.associate_fk_device(cities, 1, states, 2, t["city_state"].data_ptr(), nc, keepalive=t["city_state"])
    ctx.col_i32_device(zips, 0, t["zip_code"].data_ptr(), nz, keepalive=t["zip_code"])
    ctx.col_i32_device(zips, 1, t["zip_pop"].data_ptr(), nz, keepalive=t["zip_pop"])
    ctx.associate_fk_device(zips, 2, cities, 2, t["zip_city"].data_ptr(), nz, keepalive=t["zip_city"])
    ctx.associate_csr(states, 3, states, 4, base["adj_offsets"].astype(np.int64), base["adj_targets"])
    ctx.register("states", states)
    ctx.register("cities", cities)
    ctx.register("zips", zips)
    return DeviceGeography(zips, cities, states, U, u0, nz, nc, U * nb, t)


def plymouth_colq_query(ctx: ColqContext, lazy_fk: bool = True):
    """Runner.java:230-236 as raw C-ABI calls."""
    q = ctx.query("zips")
    q.set_option(_ffi.OPT_LAZY_FK, 1 if lazy_fk else 0)
    q.criteria_i32_range(0, 1, 10_000, 10_099)
    n = q.child(0, 2)
    n = q.child(n, 1)
    n = q.child(n, 3)
    n = q.child(n, 2)
    q.criteria_str(n, 0, 0, b"PLYMOUTH")
    return q


def north_south_north_colq_query(ctx: ColqContext):
    """Runner.java:254-259 as raw C-ABI calls."""
    q = ctx.query("states")
    q.criteria_str(0, 1, 1, b"North")
    n = q.child(0, 3)
    q.criteria_str(n, 1, 1, b"South")
    n = q.child(n, 3)
    q.criteria_str(n, 1, 1, b"North")
    return q
