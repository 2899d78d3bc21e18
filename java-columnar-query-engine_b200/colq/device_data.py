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
                              sharded: bool = False, dict_names: bool = False) -> DeviceGeography:
    """Runner.java:89-196 for this rank's universe range, generated in HBM and registered by pointer.
    ``dict_names``: the city-name column is stored dictionary-encoded (int32 codes + the distinct names)."""
    base = base or load_base()
    device = device or torch.device("cuda", ctx.device)
    u0, u1 = universe_range(n_universes, n_ranks, rank)
    U = u1 - u0
    nb = int(base["city_name_bytes"].shape[0])
    assert dict_names or U * nb < 2 ** 32 - 64, "city-name bytes exceed the uint32 offset range; use more ranks"

    def dev(a, dtype):
        return torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype)

    u = torch.arange(U, device=device, dtype=torch.int64)
    t: Dict[str, torch.Tensor] = {}
    t["zip_code"] = _padded(dev(base["zip_code"], torch.int32).repeat(U), 16)
    t["zip_pop"] = _padded(dev(base["zip_pop"], torch.int32).repeat(U), 16)
    zc = dev(base["zip_city"], torch.int64)[None, :] + (u * N_CITIES)[:, None]
    t["zip_city"] = _padded(zc.reshape(-1).to(torch.int32), 16)
    del zc
    if dict_names:
        from .engine import encode_dictionary
        from .in_memory import StringColumn
        codes1, d_off, d_bytes, _values = encode_dictionary(StringColumn(offsets=base["city_name_offsets"], data=base["city_name_bytes"]))
        t["city_name_codes"] = _padded(dev(codes1, torch.int32).repeat(U), 16)
    else:
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
    if dict_names:
        ctx.col_str_dict_device(cities, 0, t["city_name_codes"].data_ptr(), nc, d_off, d_bytes, keepalive=t["city_name_codes"])
    else:
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


# ------------------------------------------------------------------------------------------------ configs 2 and 5
def _lsr(z: torch.Tensor, k: int) -> torch.Tensor:
    """Logical shift right of int64 bit patterns (torch has no uint64 arithmetic)."""
    return (z >> k) & ((1 << (64 - k)) - 1)


def _wrap(v: int) -> int:
    """Python int -> the int64 with the same low 64 bits."""
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def splitmix64_mod_device(seed: int, start: int, n: int, m: int, device) -> torch.Tensor:
    """``splitmix64(seed, i) mod m`` for i in [start, start+n) -- bit-identical to ``geography.splitmix64`` (int64
    multiplication wraps exactly like uint64)."""
    i = torch.arange(start, start + n, device=device, dtype=torch.int64)
    z = (i + 1) * _wrap(0x9E3779B97F4A7C15) + _wrap(seed)
    z = (z ^ _lsr(z, 30)) * _wrap(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _wrap(0x94D049BB133111EB)
    z = z ^ _lsr(z, 31)
    hi, lo = _lsr(z, 32), z & 0xFFFFFFFF
    return ((hi % m) * ((1 << 32) % m) + lo % m) % m


def build_int_scan_on_device(ctx: ColqContext, n_rows: int, base=None, seed: int = 42, chunk: int = 1 << 26):
    """BASELINE config 2: ``v[i] = pops[splitmix64(42, i) mod 29353]`` as a one-column table registered as "ints"."""
    base = base or load_base()
    device = torch.device("cuda", ctx.device)
    pops = torch.from_numpy(base["zip_pop"]).to(device)
    col = torch.zeros(n_rows + 16, dtype=torch.int32, device=device)
    for s in range(0, n_rows, chunk):
        c = min(chunk, n_rows - s)
        col[s:s + c] = pops[splitmix64_mod_device(seed, s, c, N_ZIPS, device)]
    torch.cuda.synchronize(device)
    t = ctx.table_create(n_rows, _ffi.REPLICATED, 0)
    ctx.col_i32_device(t, 0, col.data_ptr(), n_rows, keepalive=col)
    ctx.register("ints", t)
    return t, col


def build_name_scan_on_device(ctx: ColqContext, n_rows: int, base=None, seed: int = 42, chunk: int = 1 << 23,
                              start: int = 0, placement: int = _ffi.REPLICATED):
    """BASELINE config 5: ``name[i] = cityNames[splitmix64(42, i) mod 25701]`` as offsets + bytes, registered as
    "names".  ``start`` / ``placement``: this rank's shard = global rows [start, start + n_rows) with shard-relative
    offsets.  Returns (table, offsets tensor, bytes tensor, idx tensor of the drawn base rows, payload bytes)."""
    base = base or load_base()
    device = torch.device("cuda", ctx.device)
    boff = torch.from_numpy(base["city_name_offsets"].astype(np.int64)).to(device)
    bbytes = torch.from_numpy(base["city_name_bytes"]).to(device)
    blen = boff[1:] - boff[:-1]
    idx = torch.empty(n_rows, dtype=torch.int32, device=device)
    off = torch.zeros(n_rows + 1 + 16, dtype=torch.int64, device=device)
    for s in range(0, n_rows, 1 << 26):
        c = min(1 << 26, n_rows - s)
        idx[s:s + c] = splitmix64_mod_device(seed, start + s, c, N_CITIES, device).to(torch.int32)
    for s in range(0, n_rows, 1 << 26):
        c = min(1 << 26, n_rows - s)
        off[s + 1:s + c + 1] = torch.cumsum(blen[idx[s:s + c].long()], 0) + off[s]
    total = int(off[n_rows].item())
    assert total < 2 ** 32 - 64
    data = torch.zeros(total + 64, dtype=torch.uint8, device=device)
    for s in range(0, n_rows, chunk):
        c = min(chunk, n_rows - s)
        ii = idx[s:s + c].long()
        lens = blen[ii]
        o0 = off[s:s + c]
        nb = int((off[s + c] - off[s]).item())
        row = torch.repeat_interleave(torch.arange(c, device=device), lens, output_size=nb)
        within = torch.arange(nb, device=device) - (o0 - off[s])[row]
        data[int(off[s].item()):int(off[s].item()) + nb] = bbytes[boff[ii][row] + within]
    off32 = ((off + 2 ** 31) % 2 ** 32 - 2 ** 31).to(torch.int32)  # uint32 bit pattern in an int32 tensor
    del off
    torch.cuda.synchronize(device)
    t = ctx.table_create(n_rows, placement, start if placement == _ffi.SHARDED else 0)
    ctx.col_str_device(t, 0, off32.data_ptr(), off32.numel() * 4, data.data_ptr(), data.numel(), n_rows, total, keepalive=(off32, data))
    ctx.register("names", t)
    return t, off32, data, idx, total


# ------------------------------------------------------------------------------------------------ cross-shard variant
def register_cross_shard_geography(ctx: ColqContext, geo, n_ranks: int, rank: int, base=None):
    """The workload of SURVEY.md 8f4 / 8e: ``geo`` (``geography.build_tables(U)``, all universes) with the city and ZIP
    tables split by PLAIN ROW RANGES -- not by universe -- so that the zip -> city keys, kept GLOBAL
    (``colq_associate_fk_global``), leave the shard at every range boundary; states replicated.  Registers this rank's
    shards under the usual names and returns their handles."""
    from .local_group import even_partition
    base = base or load_base()
    name_col = geo.cities.columns()[0]
    city_state, zip_city = geo.cities.columns()[1].fk(), geo.zips.columns()[2].fk()
    bc, bz = even_partition(geo.cities.size(), n_ranks), even_partition(geo.zips.size(), n_ranks)
    c0, c1, z0, z1 = int(bc[rank]), int(bc[rank + 1]), int(bz[rank]), int(bz[rank + 1])
    states = ctx.table_create(N_STATES, _ffi.REPLICATED, 0)
    cities = ctx.table_create(c1 - c0, _ffi.SHARDED, c0)
    zips = ctx.table_create(z1 - z0, _ffi.SHARDED, z0)
    ctx.table_partition(cities, bc)
    ctx.table_partition(zips, bz)
    ctx.col_str(states, 0, base["state_code_offsets"], base["state_code_bytes"])
    ctx.col_str(states, 1, base["state_name_offsets"], base["state_name_bytes"])
    off = name_col.offsets[c0:c1 + 1].astype(np.int64)
    ctx.col_str(cities, 0, (off - off[0]).astype(np.uint32), name_col.data[int(off[0]):int(off[-1])])
    ctx.associate_fk(cities, 1, states, 2, city_state[c0:c1])
    ctx.col_i32(zips, 0, geo.zips.columns()[0].ints()[z0:z1])
    ctx.col_i32(zips, 1, geo.zips.columns()[1].ints()[z0:z1])
    ctx.associate_fk_global(zips, 2, cities, 2, zip_city[z0:z1])
    ctx.associate_csr(states, 3, states, 4, base["adj_offsets"].astype(np.int64), base["adj_targets"])
    for name, tb in (("states", states), ("cities", cities), ("zips", zips)):
        ctx.register(name, tb)
    return states, cities, zips
