"""GPU suite (-m gpu): libcolq.so on a B200, called through the C ABI, bit-exact against the CPU oracle."""
import numpy as np
import pytest

import tck
from colq import (Association, Criteria, InMemoryTable, Query, QueryResult, geography as G, int_range, of_columns,
                  of_ints, of_strings)
from colq.data_system import StringPredicate
from colq.in_memory import IntegerColumn, StringColumn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    from colq.engine import DataSystemColq
    from oracle_system import OracleDataSystem
    made = []

    def new_gpu(lazy_fk=True, **kw):
        ds = DataSystemColq(0, lazy_fk=lazy_fk, **kw)
        made.append(ds)
        return ds

    yield new_gpu, OracleDataSystem
    for ds in made:
        ds.close()


LAZY, EAGER = dict(lazy_fk=True), dict(lazy_fk=False)
NO_DEFER = dict(lazy_fk=True, options={5: 0})            # COLQ_OPT_DEFER_CHAINS=0: FK chains inside the row scan
HOST = dict(lazy_fk=True, residency="host")              # columns stay in pinned host memory; DMA upload at the first full scan
HOST_NO_PROMOTE = dict(lazy_fk=True, residency="host", options={6: 0})   # COLQ_OPT_PROMOTE=0: always read in place
HOST_KERNEL_PROMOTE = dict(lazy_fk=True, residency="host", options={6: 1})   # =1: the scan kernel writes the HBM copy
DICT = dict(lazy_fk=True, dictionary=True)               # string columns dictionary-encoded: code lookup row scans
DICT_HOST = dict(lazy_fk=True, dictionary=True, residency="host")
LOOKBACK = dict(lazy_fk=True, options={4: 2})            # COLQ_OPT_FUSED_COMPACT=2: single-pass look-back compaction
DICT_ALL = dict(lazy_fk=True, dictionary="all")          # integer columns dictionary-encoded too
DICT_ALL_HOST = dict(lazy_fk=True, dictionary="all", residency="host")
INGEST_DEVICE = dict(lazy_fk=True, ingest="device")     # associations go up as CSRs, the GPU validates and classifies them
INGEST_DEVICE_DICT = dict(lazy_fk=True, ingest="device", dictionary=True)   # + string columns dictionary-encoded by the GPU
INGEST_DEVICE_DICT_HOST = dict(lazy_fk=True, ingest="device", dictionary=True, residency="host")
SPLIT_ROOT = dict(lazy_fk=True, options={9: 2})          # COLQ_OPT_ROOT_FUSED=2: scan_rows<..,list> + root_finish launches
SPLIT_ROOT_HOST = dict(lazy_fk=True, residency="host", options={9: 2})
UNFUSED = dict(lazy_fk=True, options={9: 0})             # COLQ_OPT_ROOT_FUSED=0: scan_rows / csr_pull / compact_fused launches (r01 plan)
ALL_VARIANTS = (LAZY, EAGER, NO_DEFER, HOST, HOST_NO_PROMOTE, HOST_KERNEL_PROMOTE, DICT, DICT_HOST, LOOKBACK, DICT_ALL, DICT_ALL_HOST, UNFUSED, SPLIT_ROOT, SPLIT_ROOT_HOST, INGEST_DEVICE, INGEST_DEVICE_DICT, INGEST_DEVICE_DICT_HOST)


def both(engines, build, queries, lazy_modes=(True, False), variants=None):
    """Build the same tables for the oracle and the GPU engine, run each query on both, compare row sets bit for bit.
    Every query runs twice per engine variant: with host-resident columns the first run streams (and promotes) them,
    the second reads the promoted HBM copies."""
    new_gpu, new_oracle = engines
    variants = variants if variants is not None else [dict(lazy_fk=lz) for lz in lazy_modes]
    oracle = new_oracle()
    build(oracle)
    want = []
    for q in queries:
        r = oracle.execute(q())
        assert isinstance(r, QueryResult.Success), getattr(r, "message", None)
        want.append((oracle.last_indices.copy(), oracle.last_words.copy(), oracle.node_cardinalities()))
    for variant, rep_i in [(v, i) for v in variants for i in range(2 if v.get("residency") == "host" else 1)]:
        lazy = variant.get("lazy_fk", True)
        if rep_i == 0:
            gpu = new_gpu(**variant)
            build(gpu)
        for q, (idx, words, cards) in zip(queries, want):
            r = gpu.execute(q())
            assert isinstance(r, QueryResult.Success), getattr(r, "message", None)
            cq = gpu.last_query
            got = cq.fetch(want_indices=True, want_bitmask=True, n_rows=words.shape[0] * 64)
            assert got.count == idx.shape[0]
            assert np.array_equal(got.indices, idx), "row index set differs from the oracle"
            assert np.array_equal(got.bitmask[: words.shape[0]], words), "BitSet words differ from the oracle"
            gc = cq.node_cardinalities()
            assert len(gc) == len(cards)
            for a, b in zip(gc, cards):
                assert a == -1 or a == b, (gc, cards)   # -1: node fused away, never materialised
            if not lazy:
                assert gc[0] == cards[0]
        if rep_i == (1 if variant.get("residency") == "host" else 0):
            gpu.close()


# ------------------------------------------------------------------ the reference's own tests + failure paths
@pytest.mark.parametrize("lazy", [True, False])
@pytest.mark.parametrize("case", tck.REFERENCE_TESTS + tck.FAILURE_TESTS + tck.EXTRA_TESTS, ids=lambda f: f.__name__)
def test_tck(engines, case, lazy):
    new_gpu, _ = engines
    case(lambda: new_gpu(lazy))


@pytest.mark.parametrize("variant", [HOST, DICT, DICT_HOST, DICT_ALL, DICT_ALL_HOST], ids=["host", "dict", "dict_host", "dict_all", "dict_all_host"])
@pytest.mark.parametrize("case", tck.REFERENCE_TESTS + tck.FAILURE_TESTS + tck.EXTRA_TESTS, ids=lambda f: f.__name__)
def test_tck_other_physical_layouts(engines, case, variant):
    """The reference's QueryTest + failure cases again, with host-resident and dictionary-encoded columns."""
    new_gpu, _ = engines
    case(lambda: new_gpu(**variant))


def test_headline_queries_one_universe(engines, expected, base_geography):
    new_gpu, _ = engines
    tck.plymouth(new_gpu, expected, base_geography)
    tck.north_south_north(new_gpu, expected, base_geography)
    tck.plymouth(lambda: new_gpu(False), expected, base_geography)


def test_opaque_lambda_is_a_failure_not_a_fallback(engines):
    new_gpu, _ = engines
    ds = new_gpu()
    ds.register("ints", of_columns(of_ints(1, 2, 3)))
    q = Query("ints")
    q.root_node.add_criteria(Criteria.IntCriteria(0, lambda i: i > 1))
    r = ds.execute(q)
    assert isinstance(r, QueryResult.Failure) and "no CPU fallback" in r.message


# ------------------------------------------------------------------ int scans: sizes around every tile boundary
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 127, 128, 129, 511, 512, 513, 4095, 4096, 4097, 32767, 32768, 32769,
                               100_003, (1 << 20) + 7])
def test_int_range_scan_sizes(engines, n):
    rng = np.random.default_rng(n)
    a = rng.integers(-50, 50, size=n, dtype=np.int32)
    b = rng.integers(np.iinfo(np.int32).min, np.iinfo(np.int32).max, size=n, dtype=np.int32, endpoint=True)

    def build(ds):
        ds.register("t", InMemoryTable.of_columns(IntegerColumn(a), IntegerColumn(b), IntegerColumn(a[::-1].copy())))

    def q1():
        q = Query("t"); q.root_node.add_criteria(Criteria.IntCriteria(0, int_range(-3, 7))); return q

    def q2():  # two predicates on two columns, extremes of the int range
        q = Query("t")
        q.root_node.add_criteria(Criteria.IntCriteria(0, int_range(-50, 49)))
        q.root_node.add_criteria(Criteria.IntCriteria(1, int_range(-(2 ** 31), 0)))
        return q

    def q3():  # three predicates: needs two fused launches
        q = Query("t")
        for o, (lo, hi) in enumerate([(-10, 30), (-(2 ** 30), 2 ** 31 - 1), (-40, 10)]):
            q.root_node.add_criteria(Criteria.IntCriteria(o, int_range(lo, hi)))
        return q

    def q4():  # empty interval and no criteria at all
        q = Query("t"); q.root_node.add_criteria(Criteria.IntCriteria(0, int_range(5, 4))); return q

    def q5():
        return Query("t")

    both(engines, build, [q1, q2, q3, q4, q5], variants=(LAZY, HOST, DICT_ALL) if n < 100_000 else (LAZY, HOST, DICT_ALL, DICT_ALL_HOST))


# ------------------------------------------------------------------ string scans: every operator, ragged lengths
def random_strings(rng, n, max_len, alphabet):
    lens = rng.integers(0, max_len + 1, size=n)
    return ["".join(rng.choice(alphabet, size=l)) for l in lens]


@pytest.mark.parametrize("n,max_len", [(0, 4), (1, 4), (1023, 6), (1024, 6), (1025, 6), (5000, 40), (70_000, 12),
                                      (400_000, 9), (3000, 300)])
def test_string_ops(engines, n, max_len):
    rng = np.random.default_rng(n * 31 + max_len)
    alphabet = np.array(list("abAB "))
    strings = random_strings(rng, n, max_len, alphabet)
    if n > 10:
        strings[3] = ""
        strings[7] = "é😀ﬁ"          # multi-byte UTF-8, incl. a supplementary code point
        strings[9] = "ab" * 700                 # longer than any ring slot: read from global memory
    col = StringColumn(strings)
    needles = ["", "a", "ab", "aB a", "abABa", "b" * 9, "é", "😀", "ﬁ", "ab" * 700]

    def build(ds):
        ds.register("s", InMemoryTable.of_columns(col))

    queries = []
    for op in range(9):
        for nd in needles:
            def mk(op=op, nd=nd):
                q = Query("s")
                q.root_node.add_criteria(Criteria.StringCriteria(0, StringPredicate(op, nd)))
                return q
            queries.append(mk)
    both(engines, build, queries, variants=(LAZY, HOST, HOST_KERNEL_PROMOTE, DICT, DICT_HOST) if n in (1025, 3000, 70_000) else (LAZY, DICT))


def test_two_string_criteria_and_int_on_one_node(engines):
    rng = np.random.default_rng(5)
    n = 9000
    s = StringColumn(random_strings(rng, n, 7, np.array(list("xyz"))))
    v = IntegerColumn(rng.integers(0, 100, size=n, dtype=np.int32))

    def build(ds):
        ds.register("t", InMemoryTable.of_columns(s, v))

    def q():
        qq = Query("t")
        qq.root_node.add_criteria(Criteria.StringCriteria(0, StringPredicate(1, "xy")))
        qq.root_node.add_criteria(Criteria.IntCriteria(1, int_range(10, 60)))
        qq.root_node.add_criteria(Criteria.StringCriteria(0, StringPredicate(2, "xz")))
        return qq

    both(engines, build, [q], variants=(LAZY, HOST, HOST_NO_PROMOTE, HOST_KERNEL_PROMOTE, DICT, DICT_HOST))


# ------------------------------------------------------------------ associations: random graphs, forward and reverse hops
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_association_graphs(engines, seed):
    rng = np.random.default_rng(seed)
    na, nb, nc = 20_011, 3_001, 97
    a_val = rng.integers(0, 1000, size=na, dtype=np.int32)
    b_val = rng.integers(0, 1000, size=nb, dtype=np.int32)
    c_val = rng.integers(0, 1000, size=nc, dtype=np.int32)
    a_to_b = rng.integers(-1, nb, size=na, dtype=np.int32)          # to-one with Nones
    b_to_c = rng.integers(-1, nc, size=nb, dtype=np.int32)
    deg = rng.integers(0, 5, size=nc)
    c_off = np.zeros(nc + 1, dtype=np.int64); np.cumsum(deg, out=c_off[1:])
    c_tgt = rng.integers(0, nc, size=int(c_off[-1]), dtype=np.int32)  # to-many self association on c
    deg2 = rng.integers(0, 3, size=nb)
    bm_off = np.zeros(nb + 1, dtype=np.int64); np.cumsum(deg2, out=bm_off[1:])
    bm_tgt = rng.integers(0, na, size=int(bm_off[-1]), dtype=np.int32)  # to-many b -> a

    def build(ds):
        A = InMemoryTable.of_columns(IntegerColumn(a_val))
        B = InMemoryTable.of_columns(IntegerColumn(b_val))
        Cc = InMemoryTable.of_columns(IntegerColumn(c_val))
        A.associate_to(B, fk=a_to_b)            # A.1 -> B ; B.1 <- A
        B.associate_to(Cc, fk=b_to_c)           # B.2 -> C ; C.1 <- B
        Cc.associate_to(Cc, csr=(c_off, c_tgt))  # C.2 -> C ; C.3 <- C
        B.associate_to(A, csr=(bm_off, bm_tgt))  # B.3 -> A (many) ; A.2 <- B
        ds.register("A", A); ds.register("B", B); ds.register("C", Cc)

    def chain_down():   # A -> B -> C -> C (forward fk, fk, csr) with criteria at root and leaf
        q = Query("A")
        q.root_node.add_criteria(Criteria.IntCriteria(0, int_range(0, 300)))
        q.root_node.create_child(1).create_child(2).create_child(2).add_criteria(Criteria.IntCriteria(0, int_range(0, 200)))
        return q

    def chain_down_no_root_pred():   # eager first hop
        q = Query("A")
        q.root_node.create_child(1).create_child(2).add_criteria(Criteria.IntCriteria(0, int_range(0, 500)))
        return q

    def chain_up():     # C <- B <- A through the reverse columns
        q = Query("C")
        q.root_node.create_child(1).create_child(1).add_criteria(Criteria.IntCriteria(0, int_range(0, 20)))
        return q

    def mid_criteria():  # criteria on every level, reverse csr hop C.3
        q = Query("C")
        q.root_node.add_criteria(Criteria.IntCriteria(0, int_range(100, 900)))
        n = q.root_node.create_child(3); n.add_criteria(Criteria.IntCriteria(0, int_range(0, 700)))
        n.create_child(1).add_criteria(Criteria.IntCriteria(0, int_range(0, 100)))
        return q

    def two_children():  # B has a to-one child and a to-many child and a reverse child
        q = Query("B")
        q.root_node.create_child(2).add_criteria(Criteria.IntCriteria(0, int_range(0, 600)))
        q.root_node.create_child(3).add_criteria(Criteria.IntCriteria(0, int_range(0, 400)))
        q.root_node.create_child(1).add_criteria(Criteria.IntCriteria(0, int_range(500, 999)))
        return q

    def reverse_of_many():  # A.2 is the reverse of the to-many B.3
        q = Query("A")
        q.root_node.create_child(2).add_criteria(Criteria.IntCriteria(0, int_range(0, 50)))
        return q

    def no_criteria_anywhere():
        q = Query("A")
        q.root_node.create_child(1).create_child(2)
        return q

    both(engines, build, [chain_down, chain_down_no_root_pred, chain_up, mid_criteria, two_children, reverse_of_many,
                          no_criteria_anywhere], variants=ALL_VARIANTS)


# ------------------------------------------------------------------ parallel universes
@pytest.mark.parametrize("U", [3, 40])
def test_plymouth_universes_match_oracle(engines, base_geography, U):
    geo = G.build_tables(U, base=base_geography)

    def build(ds):
        G.register_geography(ds, geo)

    both(engines, build, [G.plymouth_query, G.north_south_north_query], variants=ALL_VARIANTS)


def test_deferred_chains_are_planned_into_the_compaction(engines, base_geography):
    """The root's lazy FK chains move from the row scan into the fused compaction kernel (COLQ_OPT_DEFER_CHAINS)."""
    new_gpu, _ = engines
    geo = G.build_tables(2, base=base_geography)
    for opts, want_names in (({9: 1}, ["root_fused<1,1>+csr"]), ({9: 2}, ["scan_rows<1,0,list>", "root_finish<1>+csr"]), ({9: 0}, ["scan_rows<1,0,lazy>", "csr_pull", "compact_fused+chains"]),
                             ({4: 2}, ["scan_rows<1,0,lazy>", "compact_lookback+chains"]), ({5: 0}, ["scan_rows<1,1,lazy>", "compact_fused"])):
        ds = new_gpu(options=opts)
        G.register_geography(ds, geo)
        assert isinstance(ds.execute(G.plymouth_query()), QueryResult.Success)
        names = [n for n, *_ in ds.last_query.profile()]
        for w in want_names:
            assert w in names, (opts, names)
        ds.close()


def test_host_resident_columns_stream_then_promote(base_geography):
    """colq_*_host: nothing is copied at registration; the first query moves exactly the columns it scans in full over
    PCIe (h2d_bytes says how much) and leaves them in HBM, the second query moves nothing -- in both promotion modes."""
    from colq.engine import DataSystemColq
    from oracle_system import OracleDataSystem
    U = 7
    geo = G.build_tables(U, base=base_geography)
    oracle = OracleDataSystem()
    G.register_geography(oracle, geo)
    oracle.execute(G.plymouth_query())
    nz, nc = U * G.N_ZIPS, U * G.N_CITIES
    name_bytes = int(np.asarray(base_geography["city_name_bytes"]).shape[0]) * U
    for promote in (2, 1, 0):
        ds = DataSystemColq(0, residency="host", options={6: promote})
        G.register_geography(ds, geo)
        streamed = []
        for _ in range(3):
            r = ds.execute(G.plymouth_query())
            assert isinstance(r, QueryResult.Success)
            assert np.array_equal(ds.last_query.fetch(want_indices=True).indices, oracle.last_indices)
            streamed.append(int(ds.last_timing.h2d_bytes))
        touched = 4 * nz + 4 * (nc + 1) + name_bytes     # population + name offsets + name bytes, nothing else
        assert streamed[0] == touched
        assert streamed[1:] == ([0, 0] if promote else [touched, touched])
        ds.close()
    oracle.close()


def test_host_resident_fk_is_range_checked_on_walked_rows():
    """A to-one target outside the associated table is the reference's NPE at associateTo (M/InMemoryTable.java:70-71).
    Host-resident association columns are never read in bulk, so the check happens on the rows a query walks."""
    from colq import _ffi
    from colq.engine import ColqContext
    ctx = ColqContext(0)
    n = 5000
    vals = ctx.host_column(np.arange(n, dtype=np.int32), np.int32)
    fk = np.zeros(n, dtype=np.int32)
    fk[4321] = 99            # parent table has 3 rows
    fk[17] = -1              # Association.None is fine
    fkh = ctx.host_column(fk, np.int32)
    child, parent = ctx.table_create(n), ctx.table_create(3)
    ctx.col_i32(parent, 0, np.array([1, 2, 3], dtype=np.int32))
    ctx.col_i32_host(child, 0, vals)
    ctx.associate_fk_host(child, 1, parent, 1, fkh)
    ctx.register("child", child)
    ctx.register("parent", parent)
    for lo, hi, bad in ((0, 100, False), (4000, 4400, True)):
        for defer in (1, 0):
            q = ctx.query("child")
            q.set_option(_ffi.OPT_DEFER_CHAINS, defer)
            q.criteria_i32_range(0, 0, lo, hi)
            q.child(0, 1)
            if bad:
                with pytest.raises(TypeError, match="outside the associated table"):
                    q.execute()
            else:
                assert q.execute().count == hi - lo + 1 - 1   # row 17 is None
            q.close()
    # reverse direction: parent <- child push walks only the matching child rows
    q = ctx.query("parent")
    c = q.child(0, 1)
    q.criteria_i32_range(c, 0, 4321, 4321)
    with pytest.raises(TypeError, match="outside the associated table"):
        q.execute()
    q.close()
    ctx.close()


def test_three_launch_compaction_path_matches(engines, base_geography):
    """COLQ_OPT_FUSED_COMPACT: 0 = popcount / scan / write launches, 1 = cooperative two-phase kernel (default),
    2 = single-pass look-back kernel; all three must agree with the oracle."""
    from colq import _ffi
    from colq.engine import DataSystemColq
    _, new_oracle = engines
    geo = G.build_tables(5, base=base_geography)
    oracle = new_oracle()
    G.register_geography(oracle, geo)
    oracle.execute(G.plymouth_query())
    for fused in (0, 1, 2):
        ds = DataSystemColq(0, options={_ffi.OPT_FUSED_COMPACT: fused})
        G.register_geography(ds, geo)
        r = ds.execute(G.plymouth_query())
        assert isinstance(r, QueryResult.Success)
        got = ds.last_query.fetch(want_indices=True)
        assert np.array_equal(got.indices, oracle.last_indices)
        names = [n for n, *_ in ds.last_query.profile()]
        assert any(n.startswith("root_f") for n in names) == (fused == 1)   # root_fused / root_finish: the root's chains and compaction fused
        assert not any(n.startswith("compact_fused") for n in names)
        assert any(n.startswith("compact_lookback") for n in names) == (fused == 2)
        ds.close()


def test_plymouth_full_size_properties(base_geography, expected):
    """BASELINE config 4 at full size (10k universes, 293.5 M ZIP rows) generated in HBM: the result must be exactly
    {u * 29353 + r} for the 31 one-universe rows r, ascending, for both execution strategies."""
    import torch
    from colq.device_data import build_geography_on_device, plymouth_colq_query
    from colq.engine import ColqContext
    from oracle_system import OracleDataSystem
    oracle = OracleDataSystem()
    G.register_geography(oracle, G.build_tables(1, base=base_geography))
    oracle.execute(G.plymouth_query())
    rows1 = oracle.last_indices.astype(np.int64)
    U = 10_000
    ctx = ColqContext(0)
    geo = build_geography_on_device(ctx, U, base=base_geography)
    want = (np.arange(U, dtype=np.int64)[:, None] * G.N_ZIPS + rows1[None, :]).reshape(-1)
    for lazy in (True, False):
        q = plymouth_colq_query(ctx, lazy_fk=lazy)
        res = q.execute(want_indices=True, want_bitmask=True, n_rows=geo.n_zip_rows, index_capacity=31 * U)
        assert res.count == 31 * U
        assert np.array_equal(res.indices.astype(np.int64), want)
        words = res.bitmask
        assert int(np.unpackbits(words.view(np.uint8)).sum()) == 31 * U   # checksum of the mask agrees with the list
        q.close()
    del geo
    ctx.close()
    torch.cuda.empty_cache()


# ------------------------------------------------------------------ opaque lambdas over dictionary-encoded columns
def test_reference_lambdas_run_unchanged_over_dictionary_columns(engines, base_geography, expected):
    """The reference's criteria are opaque lambdas (DS/Criteria.java:17; app/.../Runner.java:236,255-259).  With
    dictionary-encoded string columns the engine evaluates them once per DISTINCT value on the host and the GPU row scan
    tests code bits: the Runner's own lambdas give the oracle's answers (the oracle runs the structured twins)."""
    new_gpu, new_oracle = engines
    geo = G.build_tables(3, base=base_geography)
    oracle = new_oracle()
    G.register_geography(oracle, geo)
    oracle.execute(G.plymouth_query())
    want_ply = oracle.last_indices.copy()
    oracle.execute(G.north_south_north_query())
    want_nsn = oracle.last_indices.copy()

    def plymouth_lambda():   # Runner.java:230-236 with "PLYMOUTH"::equals as a plain callable
        q = Query("zips")
        q.root_node.add_criteria(Criteria.IntCriteria(1, int_range(10_000, 10_099)))
        q.root_node.create_child(2).create_child(1).create_child(3).create_child(2).add_criteria(
            Criteria.StringCriteria(0, lambda s: s == "PLYMOUTH"))
        return q

    def nsn_lambda():        # Runner.java:254-259 with s -> s.contains(..)
        q = Query("states")
        q.root_node.add_criteria(Criteria.StringCriteria(1, lambda s: "North" in s))
        n = q.root_node.create_child(3)
        n.add_criteria(Criteria.StringCriteria(1, lambda s: "South" in s))
        n.create_child(3).add_criteria(Criteria.StringCriteria(1, lambda s: "North" in s))
        return q

    def plymouth_all_lambdas():   # Runner.java:230-236 verbatim: i -> i >= 10_000 && i < 10_100 and "PLYMOUTH"::equals
        q = Query("zips")
        q.root_node.add_criteria(Criteria.IntCriteria(1, lambda i: i >= 10_000 and i < 10_100))
        q.root_node.create_child(2).create_child(1).create_child(3).create_child(2).add_criteria(
            Criteria.StringCriteria(0, lambda s: s == "PLYMOUTH"))
        return q

    for variant in (DICT_ALL, DICT_ALL_HOST):   # every lambda of the app runs unchanged
        ds = new_gpu(**variant)
        G.register_geography(ds, geo)
        r = ds.execute(plymouth_all_lambdas())
        assert isinstance(r, QueryResult.Success), getattr(r, "message", None)
        assert np.array_equal(ds.last_query.fetch(want_indices=True).indices, want_ply)
        ds.close()
    ds = new_gpu(**DICT)                        # strings only: the opaque IntPredicate is still a Failure
    G.register_geography(ds, geo)
    r = ds.execute(plymouth_all_lambdas())
    assert isinstance(r, QueryResult.Failure) and "opaque IntPredicate" in r.message
    ds.close()

    for variant in (DICT, DICT_HOST):
        ds = new_gpu(**variant)
        G.register_geography(ds, geo)
        for make, want in ((plymouth_lambda, want_ply), (nsn_lambda, want_nsn)):
            r = ds.execute(make())
            assert isinstance(r, QueryResult.Success), getattr(r, "message", None)
            assert np.array_equal(ds.last_query.fetch(want_indices=True).indices, want)
        names = [n for n, *_ in ds.last_query.profile()]
        assert any(n.startswith("scan_codes") for n in names) and not any(n.startswith("scan_str") for n in names), names
        # an arbitrary predicate no structured operator covers
        q = Query("cities")
        q.root_node.add_criteria(Criteria.StringCriteria(0, lambda s: len(s) == 7 and s[::-1] < s))
        r = ds.execute(q)
        assert isinstance(r, QueryResult.Success)
        names_col = geo.cities.columns()[0]
        want = np.array([i for i in range(names_col.height()) if (lambda s: len(s) == 7 and s[::-1] < s)(names_col.get(i))], dtype=np.int32)
        assert np.array_equal(ds.last_query.fetch(want_indices=True).indices, want)
        ds.close()
    # without a dictionary an opaque lambda is still a Failure, never a CPU row scan
    ds = new_gpu()
    G.register_geography(ds, geo)
    r = ds.execute(plymouth_lambda())
    assert isinstance(r, QueryResult.Failure) and "no CPU" in r.message
    ds.close()


def test_dictionary_abi_errors():
    from colq import _ffi
    from colq.engine import ColqContext, ColqError, accept_words
    ctx = ColqContext(0)
    t = ctx.table_create(4)
    d_off = np.array([0, 1, 3], dtype=np.uint32)
    d_bytes = np.frombuffer(b"abb", dtype=np.uint8)
    with pytest.raises(ValueError, match="dictionary code outside"):
        ctx.col_str_dict(t, 0, np.array([0, 1, 2, 0], dtype=np.int32), d_off, d_bytes)
    ctx.col_str_dict(t, 0, np.array([0, 1, 1, 0], dtype=np.int32), d_off, d_bytes)
    ctx.col_str(t, 1, np.array([0, 1, 2, 3, 4], dtype=np.uint32), np.frombuffer(b"wxyz", dtype=np.uint8))
    ctx.register("t", t)
    q = ctx.query("t")
    q.criteria_str_accept(0, 0, accept_words([False, True]), 2)
    r = q.execute()
    assert r.count == 2 and list(r.indices) == [1, 2]
    q.close()
    q = ctx.query("t")
    q.criteria_str(0, 0, 0, b"a")          # structured EQ over the dictionary
    r = q.execute()
    assert list(r.indices) == [0, 3]
    q.close()
    q = ctx.query("t")
    q.criteria_str_accept(0, 0, accept_words([True, True, True]), 3)   # wrong dictionary size
    with pytest.raises(ValueError, match="accept set has 3 entries"):
        q.execute()
    q.close()
    q = ctx.query("t")
    q.criteria_str_accept(0, 1, accept_words([True] * 4), 4)           # plain string column: Failure, no CPU scan
    with pytest.raises(ColqError) as e:
        q.execute()
    assert e.value.status == _ffi.FAILURE and "dictionary-encoded" in str(e.value)
    q.close()
    ctx.close()


def test_large_dictionary_is_looked_up_in_global_memory(engines):
    """More than 262144 distinct values: the accept mask no longer fits the shared-memory staging of scan_codes."""
    n = 300_000
    col = StringColumn([f"k{i:06d}" if i % 7 else "dup" for i in range(n)])

    def build(ds):
        ds.register("s", InMemoryTable.of_columns(col))

    def mk(op, nd):
        def q():
            qq = Query("s")
            qq.root_node.add_criteria(Criteria.StringCriteria(0, StringPredicate(op, nd)))
            return qq
        return q

    both(engines, build, [mk(0, "dup"), mk(1, "9999"), mk(2, "k150000"), mk(7, "k29")], variants=(DICT, DICT_HOST))


# ------------------------------------------------------------------ result materialisation on the device
def _same_tables(a, b):
    from colq.in_memory import AssociationColumn, BooleanColumn
    assert a.width() == b.width() and a.size() == b.size()
    for ca, cb in zip(a.columns(), b.columns()):
        assert type(ca) is type(cb)
        if isinstance(ca, IntegerColumn):
            assert np.array_equal(ca.ints(), cb.ints())
        elif isinstance(ca, StringColumn):
            assert np.array_equal(ca.offsets, cb.offsets) and np.array_equal(ca.data, cb.data)
        elif isinstance(ca, BooleanColumn):
            assert np.array_equal(ca.bools(), cb.bools())
        else:
            ka, oa, ta = ca.csr()
            kb, ob, tb = cb.csr()
            assert np.array_equal(ka, kb) and np.array_equal(oa, ob) and np.array_equal(ta, tb)


@pytest.mark.parametrize("variant", [dict(), dict(residency="host"), dict(dictionary=True), dict(dictionary=True, residency="host")],
                         ids=["device", "host", "dict", "dict_host"])
def test_result_table_gathered_on_device_equals_host_subset(engines, variant):
    """Table.subset (M/InMemoryTable.java:106-159): the result table whose columns were gathered by the GPU
    (colq_result_*) equals the one the registered table builds on the host, column by column, for every column kind."""
    from colq.in_memory import BooleanColumn
    new_gpu, _ = engines
    rng = np.random.default_rng(11)
    n, m = 30_011, 257
    strings = random_strings(rng, n, 9, np.array(list("abcé ")))
    strings[5] = "x" * 3000                                     # a long row: copied by the whole warp
    vals = rng.integers(-1000, 1000, size=n, dtype=np.int32)
    flags = rng.integers(0, 2, size=n).astype(bool)
    fk = rng.integers(-1, m, size=n, dtype=np.int32)
    deg = rng.integers(0, 4, size=n)
    off = np.zeros(n + 1, dtype=np.int64); np.cumsum(deg, out=off[1:])
    tgt = rng.integers(0, m, size=int(off[-1]), dtype=np.int32)

    def build(ds):
        A = InMemoryTable.of_columns(IntegerColumn(vals), StringColumn(strings), BooleanColumn(flags))
        B = InMemoryTable.of_columns(IntegerColumn(np.arange(m, dtype=np.int32)))
        A.associate_to(B, fk=fk)                 # A.3 to-one (with Nones), B.1 reverse
        A.associate_to(B, csr=(off, tgt))        # A.4 to-many, B.2 reverse
        ds.register("A", A); ds.register("B", B)
        return A, B

    def queries():
        q1 = Query("A"); q1.root_node.add_criteria(Criteria.IntCriteria(0, int_range(-50, 120)))
        q2 = Query("A"); q2.root_node.add_criteria(Criteria.StringCriteria(1, StringPredicate(1, "ab")))
        q3 = Query("A"); q3.root_node.add_criteria(Criteria.IntCriteria(0, int_range(5000, 6000)))     # empty result
        q4 = Query("A")                                                                                 # every row
        q5 = Query("B"); q5.root_node.create_child(1).add_criteria(Criteria.IntCriteria(0, int_range(0, 3)))  # root with reverse columns
        return [q1, q2, q3, q4, q5]

    host, dev = new_gpu(**variant), new_gpu(materialize="device", **variant)
    build(host); build(dev)
    for qh, qd in zip(queries(), queries()):
        rh, rd = host.execute(qh), dev.execute(qd)
        assert isinstance(rh, QueryResult.Success) and isinstance(rd, QueryResult.Success)
        _same_tables(rh.result_set, rd.result_set)
    host.close(); dev.close()


def test_result_abi_sizes_and_errors():
    from colq import _ffi
    from colq.engine import ColqContext, ColqError
    import ctypes as C
    ctx = ColqContext(0)
    t = ctx.table_create(6)
    u = ctx.table_create(2)
    ctx.col_i32(t, 0, np.array([5, 6, 7, 8, 9, 10], dtype=np.int32))
    ctx.col_str(t, 1, np.array([0, 1, 3, 3, 6, 7, 9], dtype=np.uint32), np.frombuffer(b"abbcccdee", dtype=np.uint8))
    ctx.col_i32(u, 0, np.array([1, 2], dtype=np.int32))
    ctx.associate_csr(t, 2, u, 1, np.array([0, 2, 2, 3, 3, 4, 5], dtype=np.int64), np.array([0, 1, 1, 0, 1], dtype=np.int32))
    ctx.register("t", t)
    q = ctx.query("t")
    with pytest.raises(ColqError, match="no fetched result"):
        q.result_i32(0)
    q.criteria_i32_range(0, 0, 6, 9)
    assert q.execute().count == 4
    assert q.result_count() == 4
    assert q.result_i32(0).tolist() == [6, 7, 8, 9]
    off, data = q.result_str(1)
    assert off.tolist() == [0, 2, 2, 5, 6] and bytes(data) == b"bbcccd"
    off, tg = q.result_csr(2)
    assert off.tolist() == [0, 0, 1, 1, 2] and tg.tolist() == [1, 0]
    out = np.empty(2, dtype=np.int32)
    got = C.c_int64()
    st = ctx.lib.colq_result_i32(ctx.handle, q.handle, 0, out.ctypes.data_as(C.c_void_p), 2, C.byref(got))
    assert st == _ffi.ERR_CAPACITY and got.value == 4
    with pytest.raises(ColqError, match="not a string column"):
        q.result_str(0)
    with pytest.raises(IndexError):
        q.result_i32(7)
    q.close()
    q = ctx.query("t")          # root u: column 1 is the reverse side, no stored data
    q.close()
    q = ctx.query("u") if False else None
    ctx.register("u", u)
    q = ctx.query("u")
    q.execute()
    with pytest.raises(ColqError, match="not a stored to-many"):
        q.result_csr(1)
    q.close()
    ctx.close()


def test_global_row_indices_must_fit_int32():
    """Result rows are int32 (Java int): a shard whose global row range crosses 2^31 is refused at table creation."""
    from colq import _ffi
    from colq.engine import ColqContext
    ctx = ColqContext(0)
    ctx.table_create(1000, _ffi.SHARDED, 2 ** 31 - 1 - 1000)
    with pytest.raises(ValueError, match="int32 row-index range"):
        ctx.table_create(1000, _ffi.SHARDED, 2 ** 31 - 1000)
    ctx.close()


# ------------------------------------------------------------------ ingest on the device
@pytest.mark.parametrize("n,alphabet,max_len", [(0, "ab", 3), (1, "ab", 3), (33, "ab", 2), (5000, "abc", 3), (200_000, "abcdefgh", 4),
                                                (70_000, "xy", 12), (3000, "é😀a", 3)])
def test_device_dictionary_equals_host_dictionary(n, alphabet, max_len):
    """colq_col_str_encode: codes in first-appearance order and the distinct values, bit for bit what the host's
    encode_dictionary produces (the loop the Java shim ran in round 1)."""
    from colq.engine import ColqContext, encode_dictionary
    rng = np.random.default_rng(n + max_len)
    strings = random_strings(rng, n, max_len, np.array(list(alphabet)))
    col = StringColumn(strings)
    want_codes, want_off, want_bytes, _values = encode_dictionary(col)
    ctx = ColqContext(0)
    t = ctx.table_create(n)
    ctx.col_str(t, 0, col.offsets, col.data)
    n_dict = ctx.col_str_encode(t, 0)
    assert n_dict == want_off.shape[0] - 1
    off, data = ctx.col_dict_str(t, 0)
    assert np.array_equal(off, want_off) and np.array_equal(data, want_bytes)
    ctx.register("s", t)
    # the codes themselves: an equality query per distinct value returns exactly the rows the host codes say (a few values)
    for code in list(range(min(n_dict, 3))) + ([n_dict - 1] if n_dict > 3 else []):
        needle = bytes(want_bytes[want_off[code]:want_off[code + 1]])
        q = ctx.query("s")
        q.criteria_str(0, 0, 0, needle)
        res = q.execute(want_indices=True, index_capacity=n + 1)
        assert np.array_equal(res.indices, np.flatnonzero(want_codes == code).astype(np.int32)), code
        assert any(nm.startswith("scan_codes") for nm, *_ in q.profile())
        q.close()
    assert ctx.col_str_encode(t, 0) == n_dict     # idempotent
    ctx.close()


def test_device_dictionary_of_the_city_names(base_geography):
    from colq.engine import ColqContext, encode_dictionary
    geo = G.build_tables(40, base=base_geography)
    col = geo.cities.columns()[0]
    want_codes, want_off, want_bytes, _ = encode_dictionary(col)
    ctx = ColqContext(0)
    t = ctx.table_create(col.height())
    ctx.col_str(t, 0, col.offsets, col.data)
    assert ctx.col_str_encode(t, 0) == want_off.shape[0] - 1 == 16_584
    off, data = ctx.col_dict_str(t, 0)
    assert np.array_equal(off, want_off) and np.array_equal(data, want_bytes)
    ctx.close()


def test_global_key_associations_on_a_single_rank():
    """colq_associate_fk_global / _csr_global on a context without a communicator: the keys are global rows of the target
    table, which on one rank are its local rows, so queries behave like the local calls; a target outside the table is the
    reference's NPE (M/InMemoryTable.java:70-71) and decreasing CSR offsets are refused -- both found by the GPU's own
    validation pass, and a refused call releases its ordinals."""
    from colq import _ffi
    from colq.engine import ColqContext
    rng = np.random.default_rng(23)
    ctx = ColqContext(0)
    nx, ny = 7000, 640
    x, y = ctx.table_create(nx, _ffi.SHARDED, 0), ctx.table_create(ny, _ffi.SHARDED, 0)
    ctx.table_partition(x, [0, nx])
    ctx.table_partition(y, [0, ny])
    ctx.col_i32(x, 0, np.arange(nx, dtype=np.int32))
    ctx.col_i32(y, 0, np.arange(ny, dtype=np.int32))
    fk = rng.integers(-1, ny, size=nx, dtype=np.int32)
    bad_fk = fk.copy(); bad_fk[100] = ny
    with pytest.raises(TypeError, match="outside the associated table"):
        ctx.associate_fk_global(x, 1, y, 1, bad_fk)
    ctx.associate_fk_global(x, 1, y, 1, fk)
    deg = rng.integers(0, 4, size=nx)
    off = np.zeros(nx + 1, dtype=np.int64); np.cumsum(deg, out=off[1:])
    tg = rng.integers(0, ny, size=int(off[-1]), dtype=np.int32)
    for bad_value in (ny, -1):
        bad = tg.copy(); bad[len(bad) // 2] = bad_value
        with pytest.raises(TypeError, match="outside the associated table"):
            ctx.associate_csr_global(x, 2, y, 2, off, bad)
    off_bad = off.copy(); off_bad[5], off_bad[6] = off_bad[6] + 1, off_bad[5]
    with pytest.raises(ValueError, match="non-decreasing"):
        ctx.associate_csr_global(x, 2, y, 2, off_bad, tg)
    ctx.associate_csr_global(x, 2, y, 2, off, tg)
    ctx.register("x", x); ctx.register("y", y)
    # forward hops (pull) ...
    q = ctx.query("x")
    q.criteria_i32_range(q.child(0, 1), 0, 10, 20)
    q.criteria_i32_range(q.child(0, 2), 0, 0, 50)
    res = q.execute(want_indices=True, index_capacity=nx)
    many = np.array([np.any(tg[off[i]:off[i + 1]] <= 50) for i in range(nx)])
    assert np.array_equal(res.indices, np.flatnonzero((fk >= 10) & (fk <= 20) & many).astype(np.int32))
    q.close()
    # ... and the reverse side (push)
    q = ctx.query("y")
    q.criteria_i32_range(q.child(0, 2), 0, 0, 30)
    res = q.execute(want_indices=True, index_capacity=ny)
    want = np.unique(tg[: int(off[31])])
    assert np.array_equal(res.indices, want.astype(np.int32))
    q.close()
    ctx.close()


def test_device_association_classification():
    """colq_associate: None / One rows -> dense to-one, any Many row -> CSR; bad input is rejected by the GPU's own pass."""
    from colq.engine import ColqContext
    rng = np.random.default_rng(11)
    ctx = ColqContext(0)
    nx, ny = 10_000, 333
    x, y = ctx.table_create(nx), ctx.table_create(ny)
    ctx.col_i32(x, 0, np.arange(nx, dtype=np.int32))
    ctx.col_i32(y, 0, np.arange(ny, dtype=np.int32))
    fk = rng.integers(-1, ny, size=nx, dtype=np.int32)
    deg = (fk >= 0).astype(np.int64)
    off = np.zeros(nx + 1, dtype=np.int64); np.cumsum(deg, out=off[1:])
    assert ctx.associate(x, 1, y, 1, off, fk[fk >= 0]) is True            # all None / One
    ctx.register("x", x); ctx.register("y", y)
    q = ctx.query("x")
    c = q.child(0, 1)
    q.criteria_i32_range(c, 0, 10, 20)
    res = q.execute(want_indices=True, index_capacity=nx)
    assert np.array_equal(res.indices, np.flatnonzero((fk >= 10) & (fk <= 20)).astype(np.int32))
    assert np.array_equal(q.result_i32(1), fk[res.indices])              # the stored dense form, Nones as -1
    q.close()
    deg = rng.integers(0, 4, size=nx)
    off = np.zeros(nx + 1, dtype=np.int64); np.cumsum(deg, out=off[1:])
    tg = rng.integers(0, ny, size=int(off[-1]), dtype=np.int32)
    assert ctx.associate(x, 2, y, 2, off, tg) is False                   # some Many rows: stays a CSR
    q = ctx.query("x")
    c = q.child(0, 2)
    q.criteria_i32_range(c, 0, 0, 5)
    res = q.execute(want_indices=True, index_capacity=nx)
    want = np.array([i for i in range(nx) if np.any(tg[off[i]:off[i + 1]] <= 5)], dtype=np.int32)
    assert np.array_equal(res.indices, want)
    q.close()
    bad = tg.copy(); bad[7] = ny
    with pytest.raises(TypeError, match="NullPointerException"):           # M/InMemoryTable.java:70-71
        ctx.associate(x, 3, y, 3, off, bad)
    off_bad = off.copy(); off_bad[5], off_bad[6] = off_bad[6] + 1, off_bad[5]
    with pytest.raises(ValueError):
        ctx.associate(x, 3, y, 3, off_bad, tg)
    assert ctx.associate(x, 3, y, 3, off, tg) is False                   # the ordinals were released by the failed calls
    ctx.close()


# ------------------------------------------------------------------ regressions for the round-1 advisor findings
def test_compare_to_between_supplementary_planes(engines):
    """Two different supplementary-plane lead bytes (0xF0 vs 0xF4) must not tie: String.compareTo orders them by surrogate
    value = code-point order.  Every comparison operator, needles on both planes, device and dictionary layouts."""
    strings = ["\U00010000", "\U0010FFFF", "\U0001F600", "\U000F0000a", "\uE000", "\uFFFD", "\uD7FF", "a\U00010400", "a\U0010F400",
               "a", "", "\U0001F600\U0001F600", "é", "z\uFB01"] * 3
    col = StringColumn(strings)

    def build(ds):
        ds.register("s", InMemoryTable.of_columns(col))

    queries = []
    for needle in ("\U0001F600", "\U0010FFFF", "\U00010000", "\uE000", "a\U00010400", "", "\U000F0000"):
        for op in (2, 3, 4, 5):
            def mk(op=op, needle=needle):
                q = Query("s")
                q.root_node.add_criteria(Criteria.StringCriteria(0, StringPredicate(op, needle)))
                return q
            queries.append(mk)
    both(engines, build, queries, variants=(LAZY, DICT))


def test_empty_interval_does_not_promote_a_host_column(engines):
    """COLQ_OPT_PROMOTE=1 with a valid and an EMPTY int interval on one node: the launch degenerates to clearing the mask,
    so the host-resident column must not be switched to a never-filled HBM copy -- the next query reads real data."""
    rng = np.random.default_rng(77)
    n = 50_000
    a = IntegerColumn(rng.integers(0, 1000, size=n, dtype=np.int32))
    b = IntegerColumn(rng.integers(0, 1000, size=n, dtype=np.int32))

    def build(ds):
        ds.register("t", InMemoryTable.of_columns(a, b))

    def empty_and_valid():
        q = Query("t")
        q.root_node.add_criteria(Criteria.IntCriteria(0, int_range(10, 500)))
        q.root_node.add_criteria(Criteria.IntCriteria(1, int_range(7, 3)))      # lo > hi: matches nothing
        return q

    def valid_only():
        q = Query("t")
        q.root_node.add_criteria(Criteria.IntCriteria(0, int_range(10, 500)))
        q.root_node.add_criteria(Criteria.IntCriteria(1, int_range(3, 700)))
        return q

    both(engines, build, [empty_and_valid, valid_only, empty_and_valid, valid_only], variants=(HOST_KERNEL_PROMOTE, HOST, HOST_NO_PROMOTE))


# ------------------------------------------------------------------ boolean criteria
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 31, 32, 33, 4095, 4096, 4097, 100_003, (1 << 20) + 5])
def test_boolean_scan_sizes(engines, n):
    """scan_bool against the oracle around every word / tile boundary: the four truth tables, bytes other than 0 / 1
    never reach the column (BooleanColumn holds bools), the AND with a preceding string mask and a following int
    scan, and a boolean criterion on a child pushed through a to-one column."""
    from colq.in_memory import BooleanColumn
    rng = np.random.default_rng(1000 + n)
    flags = rng.integers(0, 2, size=n).astype(bool)
    vals = rng.integers(-100, 100, size=n, dtype=np.int32)
    strings = random_strings(rng, n, 3, np.array(list("ab")))
    m = max(1, n // 7)
    fk = rng.integers(-1, n, size=m, dtype=np.int32) if n > 0 else np.full(m, -1, dtype=np.int32)

    def build(ds):
        t = InMemoryTable.of_columns(BooleanColumn(flags), IntegerColumn(vals), StringColumn(strings))
        u = InMemoryTable.of_columns(IntegerColumn(np.arange(m, dtype=np.int32)))
        u.associate_to(t, fk=fk)                      # u.1 to-one into t (with Nones), t.3 reverse
        ds.register("t", t); ds.register("u", u)

    def truth(f, t):
        def mk():
            q = Query("t"); q.root_node.add_criteria(Criteria.BooleanCriteria(0, lambda b: t if b else f)); return q
        return mk

    def mixed():
        q = Query("t")
        q.root_node.add_criteria(Criteria.StringCriteria(2, StringPredicate(1, "a")))
        q.root_node.add_criteria(Criteria.BooleanCriteria(0, lambda b: not b))
        q.root_node.add_criteria(Criteria.IntCriteria(1, int_range(-50, 60)))
        return q

    def child_forward():
        q = Query("u")
        q.root_node.create_child(1).add_criteria(Criteria.BooleanCriteria(0, lambda b: b))
        return q

    def child_reverse():
        q = Query("t")
        q.root_node.add_criteria(Criteria.BooleanCriteria(0, lambda b: b))
        q.root_node.create_child(3).add_criteria(Criteria.IntCriteria(0, int_range(0, m // 2)))
        return q

    both(engines, build, [truth(False, True), truth(True, False), truth(True, True), truth(False, False), mixed,
                          child_forward, child_reverse], variants=(LAZY, EAGER, DICT))


# ------------------------------------------------------------------ pipelined back-to-back executions (COLQ_OPT_PIPELINE)
@pytest.mark.parametrize("U", [3, 40])
def test_pipelined_back_to_back_executions(engines, base_geography, U):
    """execute is called over and over on one plan (E/DataSystemSerialIndices.java:53).  From the second execution on the
    root kernel has left the push target zeroed (no memset in front of the string scan) and the scan is launched as a
    programmatic dependent of the previous execution's root kernel.  Every execution must return the oracle's rows: after
    a run of back-to-back launches, with another query of the same context in between, with the option off, and after
    the criteria changed."""
    new_gpu, new_oracle = engines
    geo = G.build_tables(U, base=base_geography)
    oracle = new_oracle()
    G.register_geography(oracle, geo)
    assert isinstance(oracle.execute(G.plymouth_query()), QueryResult.Success)
    want = oracle.last_indices.copy()
    assert isinstance(oracle.execute(G.north_south_north_query()), QueryResult.Success)
    want_nsn = oracle.last_indices.copy()
    for opts in ({}, {11: 0}):
        ds = new_gpu(options=opts)
        G.register_geography(ds, geo)
        assert isinstance(ds.execute(G.plymouth_query()), QueryResult.Success)
        cq = ds.last_query
        ds.last_query = None   # keep it: the next ds.execute would close it
        names = [n for n, *_ in cq.profile()]
        assert any(n.startswith("scan_str") for n in names) and any(n.startswith("root_fused") for n in names), names
        first = cq.fetch(want_indices=True)
        assert np.array_equal(first.indices, want)
        launches = []
        for _ in range(8):
            cq.execute_async()
        got = cq.fetch(want_indices=True)
        assert got.count == want.shape[0] and np.array_equal(got.indices, want)
        launches.append(int(got.timing.kernel_launches))
        assert launches[-1] == 2, launches
        assert (got.timing.gpu_ms < 0) == (opts == {}), got.timing.gpu_ms   # pipelined executions are not timed one by one
        # another query of the same context between two executions: the chain is broken, results stay right
        assert isinstance(ds.execute(G.north_south_north_query()), QueryResult.Success)
        other = ds.last_query
        for _ in range(3):
            cq.execute_async()
            other.execute_async()
        assert np.array_equal(other.fetch(want_indices=True).indices, want_nsn)
        assert np.array_equal(cq.fetch(want_indices=True).indices, want)
        cq.execute_async()
        cq.execute_async()
        assert np.array_equal(cq.fetch(want_indices=True).indices, want)
        # node cardinalities: the pushed mask does not outlive a pipelined execution (-1), the others are the oracle's
        assert isinstance(oracle.execute(G.plymouth_query()), QueryResult.Success)
        for a, b in zip(cq.node_cardinalities(), oracle.node_cardinalities()):
            assert a == -1 or a == b
        cq.close()
        ds.close()
