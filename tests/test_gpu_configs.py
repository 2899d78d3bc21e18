"""GPU suite: BASELINE.json configs[1] and configs[4] at full per-GPU size, checked through properties that do not
depend on libcolq (the generator is counter-based, so the expected row set follows from the drawn base rows), plus a
bit-exact oracle run on a prefix."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config2_int_range_scan_one_billion_rows(base_geography):
    """configs[1]: population range scan + index compaction over a synthetic 1B-row int column, one B200."""
    import torch
    from colq import Criteria, InMemoryTable, Query, geography as G, int_range
    from colq.device_data import build_int_scan_on_device, splitmix64_mod_device
    from colq.engine import ColqContext
    from colq.in_memory import IntegerColumn
    from oracle_system import OracleDataSystem

    n = 1_000_000_000
    ctx = ColqContext(0)
    _t, col = build_int_scan_on_device(ctx, n, base=base_geography)
    dev = col.device
    # the device generator is bit-identical to the numpy one
    probe = np.arange(5_000_000, 5_100_000)
    want = (G.splitmix64(42, probe) % np.uint64(G.N_ZIPS)).astype(np.int64)
    assert np.array_equal(splitmix64_mod_device(42, 5_000_000, 100_000, G.N_ZIPS, dev).cpu().numpy(), want)

    q = ctx.query("ints")
    q.criteria_i32_range(0, 0, 10_000, 10_099)
    res = q.execute(want_indices=True, index_capacity=2_000_000)
    # independent expectation: rows whose drawn base ZIP is one of the 47 with population in [10000, 10099]
    v = col[:n]
    exp = torch.nonzero((v >= 10_000) & (v <= 10_099)).flatten().to(torch.int32).cpu().numpy()
    assert res.count == exp.shape[0] and 1_500_000 < res.count < 1_700_000   # ~0.160 % of 1e9 (47 / 29353)
    assert np.array_equal(res.indices, exp)
    # other selectivities (SURVEY 8d secondary sweeps): ~1 %, ~50 %, everything, nothing
    for lo, hi in ((54_000, 2 ** 31 - 1), (0, 2_797), (-(2 ** 31), 2 ** 31 - 1), (5, 4)):
        q2 = ctx.query("ints")
        q2.criteria_i32_range(0, 0, lo, hi)
        r2 = q2.execute(want_indices=False)
        assert r2.count == int(((v >= lo) & (v <= hi)).sum().item()), (lo, hi)
        q2.close()
    # bit-exact against the oracle on a 16M-row prefix
    m = 1 << 24
    prefix = v[:m].cpu().numpy()
    oracle = OracleDataSystem()
    oracle.register("ints", InMemoryTable.of_columns(IntegerColumn(prefix)))
    oq = Query("ints")
    oq.root_node.add_criteria(Criteria.IntCriteria(0, int_range(10_000, 10_099)))
    oracle.execute(oq)
    assert np.array_equal(res.indices[res.indices < m], oracle.last_indices)
    q.close()
    ctx.close()
    del col, v
    torch.cuda.empty_cache()


def test_config5_city_name_equality_one_shard(base_geography):
    """configs[4]: city-name == "PLYMOUTH" over the synthetic offsets+bytes column; one GPU's shard of the 500M-row
    column (62.5 M rows)."""
    import torch
    from colq import Criteria, InMemoryTable, Query, str_equals
    from colq.device_data import build_name_scan_on_device
    from colq.engine import ColqContext
    from colq.in_memory import StringColumn
    from oracle_system import OracleDataSystem

    n = 62_500_000
    ctx = ColqContext(0)
    _t, off32, data, idx, total = build_name_scan_on_device(ctx, n, base=base_geography)
    names = StringColumn(offsets=base_geography["city_name_offsets"], data=base_geography["city_name_bytes"]).strings()
    ply = torch.tensor([i for i, s in enumerate(names) if s == "PLYMOUTH"], device=idx.device, dtype=torch.int32)
    assert ply.numel() == 16
    exp = torch.nonzero(torch.isin(idx, ply)).flatten().to(torch.int32).cpu().numpy()
    for op, needle, want in ((0, b"PLYMOUTH", exp),):
        q = ctx.query("names")
        q.criteria_str(0, 0, op, needle)
        res = q.execute(want_indices=True, want_bitmask=True, n_rows=n, index_capacity=100_000)
        assert res.count == want.shape[0] and 30_000 < res.count < 50_000     # ~0.062 % (16 / 25701)
        assert np.array_equal(res.indices, want)
        assert int(np.unpackbits(res.bitmask.view(np.uint8)).sum()) == res.count
        q.close()
    # contains("PLYMOUTH") must also catch NEW PLYMOUTH, PLYMOUTH MEETING, ... (5 distinct names contain it)
    cont = torch.tensor([i for i, s in enumerate(names) if "PLYMOUTH" in s], device=idx.device, dtype=torch.int32)
    q = ctx.query("names")
    q.criteria_str(0, 0, 1, b"PLYMOUTH")
    res = q.execute(want_indices=True, index_capacity=200_000)
    assert np.array_equal(res.indices, torch.nonzero(torch.isin(idx, cont)).flatten().to(torch.int32).cpu().numpy())
    q.close()
    # bit-exact against the oracle on a 4M-row prefix
    m = 1 << 22
    o = off32[: m + 1].cpu().numpy().view(np.uint32)
    b = data[: int(o[-1])].cpu().numpy()
    oracle = OracleDataSystem()
    oracle.register("names", InMemoryTable.of_columns(StringColumn(offsets=o, data=b)))
    oq = Query("names")
    oq.root_node.add_criteria(Criteria.StringCriteria(0, str_equals("PLYMOUTH")))
    oracle.execute(oq)
    assert np.array_equal(exp[exp < m], oracle.last_indices)
    ctx.close()
    del off32, data, idx
    torch.cuda.empty_cache()
