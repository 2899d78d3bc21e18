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
