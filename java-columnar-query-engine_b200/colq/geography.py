"""The reference app's workload (``app/src/main/java/dgroomes/app/Runner.java``) restated over the compact model.

* ``load_base()``            -- the committed 1-universe fixture (tests/golden/geography.npz; produced from zips.jsonl +
                                StateData.java by tests/golden/make_fixtures.py, GeographiesLoader.java:51-85 rules)
* ``build_tables(U, ...)``   -- the three tables + three ``associateTo`` pairs of Runner.java:89-196, with the ZIP and city
                                tables replicated into U "parallel universes" (README.md:47) and optionally
                                restricted to one rank's universe range
* ``plymouth_query()``       -- Runner.java:230-236
* ``north_south_north_query()`` -- Runner.java:254-259

Schema (Runner.java:55-78): zips = [0 code int, 1 population int, 2 ->city]; cities = [0 name str, 1 ->state, 2 <-zips];
states = [0 code str, 1 name str, 2 <-cities, 3 ->adjacent states, 4 <-adjacent states].
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np

from .data_system import Criteria, Query, int_half_open, str_contains, str_equals
from .in_memory import InMemoryTable, IntegerColumn, StringColumn

FIXTURE = Path(__file__).resolve().parents[2] / "tests" / "golden" / "geography.npz"

N_ZIPS, N_CITIES, N_STATES = 29_353, 25_701, 51


def load_base(path: Optional[Path] = None) -> Dict[str, np.ndarray]:
    with np.load(path or FIXTURE) as z:
        return {k: z[k] for k in z.files}


def universe_range(n_universes: int, n_ranks: int, rank: int) -> Tuple[int, int]:
    """Contiguous universe range [u0, u1) owned by ``rank`` (remainder spread over the first ranks)."""
    q, r = divmod(n_universes, n_ranks)
    u0 = rank * q + min(rank, r)
    return u0, u0 + q + (1 if rank < r else 0)


@dataclass
class Geography:
    zips: InMemoryTable
    cities: InMemoryTable
    states: InMemoryTable
    n_universes: int          # universes held by these tables
    u0: int                   # first universe (global row base of zips = u0 * N_ZIPS, of cities = u0 * N_CITIES)

    @property
    def zip_row_base(self) -> int:
        return self.u0 * N_ZIPS

    @property
    def city_row_base(self) -> int:
        return self.u0 * N_CITIES


def replicate_columns(base: Dict[str, np.ndarray], n_universes: int) -> Dict[str, np.ndarray]:
    """Universe replication: exact copies, zip->city keys rebased by u * 25701 (shard-local)."""
    U = n_universes
    u = np.arange(U, dtype=np.int64)
    zip_city = (base["zip_city"].astype(np.int64)[None, :] + (u * N_CITIES)[:, None]).reshape(-1)
    nb = int(base["city_name_bytes"].shape[0])
    off = (base["city_name_offsets"][:-1].astype(np.int64)[None, :] + (u * nb)[:, None]).reshape(-1)
    off = np.concatenate([off, [U * nb]])
    assert U * nb < 2 ** 32, "city-name bytes exceed the uint32 offset range; shard the universes"
    return dict(
        zip_code=np.tile(base["zip_code"], U),
        zip_pop=np.tile(base["zip_pop"], U),
        zip_city=zip_city.astype(np.int32),
        city_name_offsets=off.astype(np.uint32),
        city_name_bytes=np.tile(base["city_name_bytes"], U),
        city_state=np.tile(base["city_state"], U),
    )


def build_tables(n_universes: int = 1, n_ranks: int = 1, rank: int = 0, base: Optional[Dict[str, np.ndarray]] = None,
                 rename_plymouth_except_last_rank: bool = False) -> Geography:
    """Runner.java:89-196 for this rank's universe range.

    ``rename_plymouth_except_last_rank`` builds the perturbed variant of SURVEY.md 8d: every city named PLYMOUTH is
    renamed ``PLYMOUTH_`` in all universes except those owned by the last rank, so the state mask exists on only one
    GPU before the OR-allreduce.
    """
    base = base or load_base()
    u0, u1 = universe_range(n_universes, n_ranks, rank)
    U = u1 - u0
    cols = replicate_columns(base, U)
    name_off, name_bytes = cols["city_name_offsets"], cols["city_name_bytes"]
    if rename_plymouth_except_last_rank and rank != n_ranks - 1:
        names = StringColumn(offsets=base["city_name_offsets"], data=base["city_name_bytes"]).strings()
        names = [n + "_" if n == "PLYMOUTH" else n for n in names]
        one = StringColumn(names)
        nb = int(one.data.shape[0])
        u = np.arange(U, dtype=np.int64)
        off = (one.offsets[:-1].astype(np.int64)[None, :] + (u * nb)[:, None]).reshape(-1)
        name_off = np.concatenate([off, [U * nb]]).astype(np.uint32)
        name_bytes = np.tile(one.data, U)

    states = InMemoryTable.of_columns(
        StringColumn(offsets=base["state_code_offsets"], data=base["state_code_bytes"]),
        StringColumn(offsets=base["state_name_offsets"], data=base["state_name_bytes"]))
    cities = InMemoryTable.of_columns(StringColumn(offsets=name_off, data=name_bytes))
    cities.associate_to(states, fk=cols["city_state"])                      # Runner.java:138
    zips = InMemoryTable.of_columns(IntegerColumn(cols["zip_code"]), IntegerColumn(cols["zip_pop"]))
    zips.associate_to(cities, fk=cols["zip_city"])                          # Runner.java:165
    states.associate_to(states, csr=(base["adj_offsets"].astype(np.int64), base["adj_targets"]))  # Runner.java:195
    return Geography(zips, cities, states, U, u0)


def plymouth_query() -> Query:
    """Runner.java:230-236: ZIPs with population in [10000, 10100) whose state is adjacent to a state with a city
    named exactly PLYMOUTH."""
    q = Query("zips")
    q.root_node.add_criteria(Criteria.IntCriteria(1, int_half_open(10_000, 10_100)))
    (q.root_node.create_child(2)      # zips -> cities
        .create_child(1)              # cities -> states
        .create_child(3)              # states -> adjacent states
        .create_child(2)              # states -> cities (reverse of cities.1)
        .add_criteria(Criteria.StringCriteria(0, str_equals("PLYMOUTH"))))
    return q


def north_south_north_query() -> Query:
    """Runner.java:254-259."""
    q = Query("states")
    (q.root_node.add_criteria(Criteria.StringCriteria(1, str_contains("North")))
        .create_child(3).add_criteria(Criteria.StringCriteria(1, str_contains("South")))
        .create_child(3).add_criteria(Criteria.StringCriteria(1, str_contains("North"))))
    return q


def register_geography(data_system, geo: Geography, sharded: bool = False) -> None:
    """Registration order of Runner.java:107,137,164 (tables are registered before their associations exist there;
    here they already exist, which the engines treat identically)."""
    if sharded:
        from . import _ffi
        data_system.register("states", geo.states)
        data_system.register("cities", geo.cities, placement=_ffi.SHARDED, global_row_base=geo.city_row_base)
        data_system.register("zips", geo.zips, placement=_ffi.SHARDED, global_row_base=geo.zip_row_base)
    else:
        data_system.register("states", geo.states)
        data_system.register("cities", geo.cities)
        data_system.register("zips", geo.zips)


def splitmix64(seed: int, i: np.ndarray) -> np.ndarray:
    """Counter-based generator shared by CPU and GPU data builders."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + (i.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
