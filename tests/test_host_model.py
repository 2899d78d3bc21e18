"""CPU suite: the host-side mirror of the data-system / in-memory model (no GPU, no oracle)."""
import numpy as np
import pytest

from colq import (NONE, Association, BitSet, InMemoryTable, Many, One, Query, of_columns, of_ints, of_strings,
                  str_compare_gt, str_contains, str_equals)
from colq import geography as G


def test_association_add_follows_reference():
    # Association.java:27-51
    a = Association.to_none().add(3)
    assert a == One(3)
    b = a.add(5)
    assert b == Many((3, 5))
    assert b.add(7) == Many((3, 5, 7))
    assert NONE.targets() == ()


def test_associate_to_builds_transposed_reverse_column():
    x = of_columns(of_strings("x0", "x1", "x2", "x3"))
    y = of_columns(of_strings("y0", "y1", "y2"))
    fwd = x.associate_to(y, Association.to_one(1), Association.to_many(0, 1), Association.to_none(), Association.to_one(1))
    assert x.width() == 2 and y.width() == 2            # InMemoryTable.java:48,88
    rev = y.columns()[1]
    assert fwd.reverse_associated_column() is rev and rev.reverse_associated_column() is fwd
    assert rev.associations_for_index(0) == One(1)       # InMemoryTable.java:75-82
    assert rev.associations_for_index(1) == Many((0, 1, 3))   # x ascending (:61)
    assert rev.associations_for_index(2) is NONE
    with pytest.raises(TypeError):                       # NPE at :70-71
        x.associate_to(y, Association.to_one(3), NONE, NONE, NONE)


def test_subset_keeps_all_columns_ascending_and_unremapped():
    t = of_columns(of_strings("a", "bb", "", "dddd"), of_ints(1, 2, 3, 4))
    u = of_columns(of_ints(9, 8))
    t.associate_to(u, Association.to_one(1), Association.to_none(), Association.to_one(0), Association.to_one(1))
    s = t.subset(BitSet.from_indices(np.array([3, 0]), 4))
    assert s.size() == 2 and s.width() == 3             # InMemoryTable.java:106-159
    assert s.columns()[0].strings() == ["a", "dddd"]
    assert s.columns()[1].ints().tolist() == [1, 4]
    assert s.columns()[2].associations_for_index(1) == One(1)   # indices are not remapped (:143-154)


def test_bitset_layout_is_java_util_bitset():
    b = BitSet.from_indices(np.array([0, 63, 64, 130]), 131)
    assert b.words.tolist() == [(1 << 63) | 1, 1, 4]
    assert b.to_indices().tolist() == [0, 63, 64, 130]
    assert b.cardinality() == 4 and b.get(64) and not b.get(65)


def test_query_duplicate_child_ordinal_throws():
    q = Query("t")
    q.root_node.create_child(2)
    with pytest.raises(ValueError):                      # Query.java:33-35
        q.root_node.create_child(2)


def test_structured_predicates_are_callables():
    assert str_equals("PLYMOUTH")("PLYMOUTH") and not str_equals("PLYMOUTH")("NEW PLYMOUTH")
    assert str_contains("North")("North Dakota") and not str_contains("North")("north")
    assert str_compare_gt("a")("b") and not str_compare_gt("a")("a")


def test_universe_replication_shapes(base_geography):
    geo = G.build_tables(3, base=base_geography)
    assert geo.zips.size() == 3 * G.N_ZIPS and geo.cities.size() == 3 * G.N_CITIES and geo.states.size() == G.N_STATES
    fk = geo.zips.columns()[2].fk()
    assert fk[G.N_ZIPS] == base_geography["zip_city"][0] + G.N_CITIES
    names = geo.cities.columns()[0]
    assert names.get(G.N_CITIES + 5) == names.get(5)
    assert [G.universe_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    shard = G.build_tables(10, n_ranks=4, rank=2, base=base_geography)
    assert shard.n_universes == 2 and shard.zip_row_base == 6 * G.N_ZIPS
    assert int(shard.zips.columns()[2].fk().max()) < shard.cities.size()   # keys are shard-local


def test_dictionary_encoding_host_logic():
    """engine.encode_dictionary / accept_words: what the shim does before any GPU call."""
    from colq.engine import accept_words, encode_dictionary
    from colq.in_memory import StringColumn
    col = StringColumn(["b", "a", "", "b", "é", "a", "b"])
    codes, d_off, d_bytes, values = encode_dictionary(col)
    assert values == ["b", "a", "", "é"]                       # first-appearance order
    assert codes.tolist() == [0, 1, 2, 0, 3, 1, 0]
    assert d_off.tolist() == [0, 1, 2, 2, 4] and bytes(d_bytes) == "baé".encode()
    empty = encode_dictionary(StringColumn([]))
    assert empty[0].shape == (0,) and empty[1].tolist() == [0] and empty[3] == []
    w = accept_words([True, False, True] + [False] * 61 + [True])
    assert w.dtype == np.uint64 and w.tolist() == [5, 1]
    assert accept_words([]).tolist() == [0]


def test_bench_reference_arm_contract(tmp_path):
    """`bench.py --impl reference` (the driver's CPU arm) prints ONE JSON line with the contract's keys, from rank 0
    only, and never touches a GPU."""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    cmd = [sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--reference-universes", "20"]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "plymouth_query_zip_rows_per_sec" and d["unit"] == "rows/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "plymouth_adjacency_query_10k_universes"
    assert d["same_config_as_gpu_arm"] is False and "20 of 10000 universes" in d["cpu_baseline"]["sample"]
    # the config dict is exactly the one the GPU arm prints (the driver compares them)
    assert d["config"] == {"workload": "plymouth_adjacency_query_10k_universes", "source": "BASELINE.json configs[3]; app/.../Runner.java:230-236",
                           "universes": 10_000, "zip_rows": 293_530_000, "city_rows": 257_010_000, "state_rows": 51,
                           "l2_policy": "inputs larger than L2 (no flush)"}
    # under torchrun the other ranks print nothing and exit 0
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_scan_bool_byte_to_bit_arithmetic():
    """scan_bool (csrc/colq_kernels.cuh, sb_truth16) turns 4 bytes into 4 bits with a SWAR nonzero-byte test and one
    gathering multiply.  Restated with numpy uint32 arithmetic and checked against the plain definition for every
    combination of zero / nonzero bytes and for random words: the partial products never carry into the result bits."""
    rng = np.random.default_rng(3)
    specials = np.array([0x00, 0x01, 0x7F, 0x80, 0xFF], dtype=np.uint32)
    combos = np.array([[a, b, c, d] for a in specials for b in specials for c in specials for d in specials], dtype=np.uint32)
    words = np.concatenate([combos[:, 0] | combos[:, 1] << 8 | combos[:, 2] << 16 | combos[:, 3] << 24,
                            rng.integers(0, 2 ** 32, size=200_000, dtype=np.uint64).astype(np.uint32)])
    nz = ((((words & np.uint32(0x7F7F7F7F)) + np.uint32(0x7F7F7F7F)) | words) & np.uint32(0x80808080)) >> np.uint32(7)
    got = ((nz * np.uint32(0x00204081)) >> np.uint32(21)) & np.uint32(0xF)      # uint32 multiply wraps like the GPU's
    want = np.zeros_like(words)
    for k in range(4):
        want |= (((words >> np.uint32(8 * k)) & np.uint32(0xFF)) != 0).astype(np.uint32) << np.uint32(k)
    assert np.array_equal(got, want)


def test_committed_bench_lines_keep_the_contract():
    """The bench lines committed under profiles/ for the final build carry every key the measurement contract names
    (metric / config.workload / roofline / cpu_baseline / e2e with its copy sizes / clocks / gpu_launches / gate)."""
    import json
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    for name, n_gpus in (("r02_session2_n1_driver_cmd.json", 1), ("r02_session2_n1.json", 1), ("r02_session2_n2.json", 2), ("r02_session2_n4.json", 4)):
        d = json.loads((root / "profiles" / name).read_text().strip().splitlines()[-1])
        assert d["metric"] == "plymouth_query_zip_rows_per_sec" and d["unit"] == "rows/s" and d["higher_is_better"] is True
        assert d["n_gpus"] == n_gpus and d["steps"] >= 20 and d["warmup"] >= 3 and d["scaling"] == "strong" and d["vs_baseline"] is None
        assert d["config"]["workload"] == "plymouth_adjacency_query_10k_universes" and d["config"]["universes"] == 10_000
        assert abs(d["value"] - d["config"]["zip_rows"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-9
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["ms_per_launch"] * 1e-3) / 1e9) / r["achieved"] < 1e-9
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
        assert d["gpu_launches"] == d["gpu_launches_per_step"] * d["steps"] > 0
        assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] == d["clocks"]["sm_max_mhz"]
        assert d["gate"]["exact"] is True and (d["gate"]["perturbed"] is True) == (n_gpus > 1)
        if n_gpus == 1 and d["cpu_baseline"] is not None:
            c = d["cpu_baseline"]
            assert c["kind"] == "port" and c["cores"] == 1 and c["unit"] == "rows/s" and "universes" in c["sample"]
