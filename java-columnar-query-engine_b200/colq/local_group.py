"""One host process driving several GPUs: ``ColqLocalGroup`` wraps ``colq_comm_init_local`` / ``colq_execute_group`` /
``colq_fetch_group`` (include/colq.h).

The reference engine is ONE object in ONE JVM (E/DataSystemSerialIndices.java:14-22, app/.../Runner.java:40); this is the
multi-GPU mode such a host can drive without one process per GPU: a context per GPU, rank i = the i-th context, mailboxes
reachable through ``cudaDeviceEnablePeerAccess``, the same exchange kernels as under ``torchrun``.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

from . import _ffi
from .engine import ColqContext, ColqQuery, ExecResult, _raise


class ColqLocalGroup:
    def __init__(self, devices: Sequence[int]):
        self.ctxs: List[ColqContext] = [ColqContext(d) for d in devices]
        self.lib = self.ctxs[0].lib
        arr = (C.c_void_p * len(self.ctxs))(*[c.handle for c in self.ctxs])
        st = self.lib.colq_comm_init_local(arr, len(self.ctxs))
        if st != _ffi.OK:
            msg = self.ctxs[0].last_error()
            self.close()
            _raise(st, msg or f"colq_comm_init_local failed with status {st}")

    @property
    def n_ranks(self) -> int:
        return len(self.ctxs)

    def _arrays(self, queries: Sequence[ColqQuery]):
        n = len(self.ctxs)
        assert len(queries) == n, "one query per rank"
        return (C.c_void_p * n)(*[c.handle for c in self.ctxs]), (C.c_void_p * n)(*[q.handle for q in queries])

    def execute_async(self, queries: Sequence[ColqQuery]) -> None:
        """Enqueue the query on every rank (``colq_execute_group``); nothing is waited for."""
        ca, qa = self._arrays(queries)
        st = self.lib.colq_execute_group(ca, qa, len(self.ctxs))
        if st != _ffi.OK:
            _raise(st, next((c.last_error() for c in self.ctxs if c.last_error()), f"status {st}"))

    def fetch(self, queries: Sequence[ColqQuery], want_indices: bool = True, want_bitmask: bool = False, n_rows: Optional[Sequence[int]] = None,
              index_capacity: int = 1 << 20) -> List[ExecResult]:
        """``colq_fetch_group`` (waits for all ranks, re-runs everybody if a result block was too small), then the per-rank
        copies."""
        ca, qa = self._arrays(queries)
        counts = (C.c_int64 * len(self.ctxs))()
        st = self.lib.colq_fetch_group(ca, qa, len(self.ctxs), counts)
        if st != _ffi.OK:
            _raise(st, next((c.last_error() for c in self.ctxs if c.last_error()), f"status {st}"))
        out = []
        for i, q in enumerate(queries):
            cap = max(index_capacity, int(counts[i]) + 1)
            out.append(q.fetch(want_indices=want_indices, want_bitmask=want_bitmask, n_rows=(n_rows[i] if n_rows else 0), index_capacity=cap))
        return out

    def execute(self, queries: Sequence[ColqQuery], **kw) -> List[ExecResult]:
        self.execute_async(queries)
        return self.fetch(queries, **kw)

    def close(self) -> None:
        for c in self.ctxs:
            c.close()
        self.ctxs = []


# ---------------------------------------------------------------------------------------------------------------------
import numpy as np  # noqa: E402

from .data_system import BitSet, DataSystem, Query, QueryResult, Table  # noqa: E402
from .engine import ColqError, DataSystemColq  # noqa: E402
from .in_memory import AssociationColumn, BooleanColumn, IntegerColumn, StringColumn  # noqa: E402


def even_partition(n_rows: int, n_ranks: int) -> np.ndarray:
    """Contiguous row ranges of (almost) equal size whose inner bounds are multiples of 64 rows (whole BitSet words), as
    ``colq_table_partition`` requires."""
    per = -(-n_rows // n_ranks)
    per = -(-per // 64) * 64
    return np.minimum(np.arange(n_ranks + 1, dtype=np.int64) * per, n_rows)


class DataSystemColqGroup(DataSystem):
    """``DataSystemSerialIndices`` (E/DataSystemSerialIndices.java:14-102) on N GPUs driven by ONE host object.

    ``register(name, table)`` takes the application's ordinary, whole tables.  At the first ``execute`` every table is
    split into contiguous row ranges, one per GPU (``sharded=False`` keeps a copy on every GPU instead); association
    columns keep the GLOBAL row indices the application wrote, so a hop may leave the shard in either direction
    (``colq_associate_*_global``: bitmap all-gather / OR-reduce-scatter over NVLink).  The result is the registered
    table's own ``subset`` of the matching GLOBAL rows -- exactly what the single-GPU engine and the reference return.
    """

    def __init__(self, devices: Sequence[int], lazy_fk: bool = True, options=None, default_sharded: bool = True):
        self.group = ColqLocalGroup(devices)
        self.lazy_fk = lazy_fk
        self.options = dict(options or {})
        self.default_sharded = default_sharded
        self._tables = {}
        self._sharded = {}
        self._handles = {}     # id(table) -> [colq_table per rank]
        self._bounds = {}      # id(table) -> partition bounds (sharded tables)
        self._uploaded = {}
        self._registered = {}
        self._pins = []
        self._translators = [DataSystemColq(context=c, lazy_fk=lazy_fk, options=self.options) for c in self.group.ctxs]
        self.last_queries = []

    def register(self, table_name: str, table: Table, sharded: Optional[bool] = None) -> None:
        self._tables[table_name] = table
        if sharded is not None:
            self._sharded[id(table)] = sharded
        self._pins.append(table)

    def _is_sharded(self, t: Table) -> bool:
        return self._sharded.get(id(t), self.default_sharded)

    def _rows(self, t: Table, rank: int):
        if self._is_sharded(t):
            b = self._bounds[id(t)]
            return int(b[rank]), int(b[rank + 1])
        return 0, t.size()

    def _sync_tables(self) -> None:
        n = self.group.n_ranks
        todo, seen = list(self._tables.values()), {}
        while todo:
            t = todo.pop()
            if id(t) in seen:
                continue
            seen[id(t)] = t
            for c in t.columns():
                if isinstance(c, AssociationColumn):
                    todo.append(c.associated_entity)
        for tid, t in seen.items():
            if tid in self._handles:
                continue
            self._pins.append(t)
            self._uploaded[tid] = 0
            hs = []
            if self._is_sharded(t):
                b = self._bounds[tid] = even_partition(t.size(), n)
                for r, ctx in enumerate(self.group.ctxs):
                    h = ctx.table_create(int(b[r + 1] - b[r]), _ffi.SHARDED, int(b[r]))
                    ctx.table_partition(h, b)
                    hs.append(h)
            else:
                hs = [ctx.table_create(t.size(), _ffi.REPLICATED, 0) for ctx in self.group.ctxs]
            self._handles[tid] = hs
        for tid, t in seen.items():
            cols = t.columns()
            for ordinal in range(self._uploaded[tid], len(cols)):
                c = cols[ordinal]
                for r, ctx in enumerate(self.group.ctxs):
                    h = self._handles[tid][r]
                    lo, hi = self._rows(t, r)
                    if isinstance(c, IntegerColumn):
                        ctx.col_i32(h, ordinal, c.ints()[lo:hi])
                    elif isinstance(c, StringColumn):
                        off = c.offsets[lo:hi + 1].astype(np.int64)
                        ctx.col_str(h, ordinal, (off - off[0]).astype(np.uint32), c.data[int(off[0]):int(off[-1])])
                    elif isinstance(c, BooleanColumn):
                        ctx.col_bool(h, ordinal, c.bools()[lo:hi])
        for tid, t in seen.items():
            cols = t.columns()
            for ordinal in range(self._uploaded[tid], len(cols)):
                c = cols[ordinal]
                if not (isinstance(c, AssociationColumn) and c.is_forward()):
                    continue
                y = c.associated_entity
                rev = c.reverse_associated_column()
                y_ordinal = next(i for i, yc in enumerate(y.columns()) if yc is rev)
                global_keys = self._is_sharded(y)    # the keys the application wrote ARE global row indices of y
                fk = c.fk()
                for r, ctx in enumerate(self.group.ctxs):
                    h, hy = self._handles[tid][r], self._handles[id(y)][r]
                    lo, hi = self._rows(t, r)
                    if fk is not None:
                        (ctx.associate_fk_global if global_keys else ctx.associate_fk)(h, ordinal, hy, y_ordinal, fk[lo:hi])
                    else:
                        _kind, offsets, targets = c.csr()
                        o = offsets[lo:hi + 1]
                        (ctx.associate_csr_global if global_keys else ctx.associate_csr)(h, ordinal, hy, y_ordinal, o - o[0], targets[int(o[0]):int(o[-1])])
        for tid, t in seen.items():
            self._uploaded[tid] = len(t.columns())
        for name, t in self._tables.items():
            hs = self._handles[id(t)]
            if self._registered.get(name) != hs[0]:
                for ctx, h in zip(self.group.ctxs, hs):
                    ctx.register(name, h)
                self._registered[name] = hs[0]

    def execute(self, query: Query):
        """E/DataSystemSerialIndices.java:53-102 over all GPUs."""
        if query is None:
            raise TypeError("NullPointerException: The 'query' argument must not be null")
        if query.table_name not in self._tables:
            return QueryResult.Failure(f"The query targets the table '{query.table_name}' but that table is not registered")
        table = self._tables[query.table_name]
        self._sync_tables()
        for q in self.last_queries:
            q.close()
        self.last_queries = []
        queries = []
        for tr in self._translators:
            tr._tables = self._tables
            cq, why = tr._translate(query)
            if cq is None:
                for q in queries:
                    q.close()
                return QueryResult.Failure(why)
            queries.append(cq)
        self.last_queries = queries
        try:
            results = self.group.execute(queries, want_indices=True)
        except ColqError as e:
            if e.status == _ffi.FAILURE:
                return QueryResult.Failure(str(e))
            raise
        if self._is_sharded(table):
            rows = results[0].indices            # every rank holds all ranks' GLOBAL rows, ascending
        else:
            rows = results[0].indices            # a replicated root: the same local answer on every rank
            for r in results[1:]:
                assert np.array_equal(r.indices, rows), "ranks disagree on a replicated root table"
        self.last_indices = rows
        return QueryResult.Success(table.subset(BitSet.from_indices(rows, table.size())))

    def close(self) -> None:
        for q in self.last_queries:
            q.close()
        self.last_queries = []
        self.group.close()
