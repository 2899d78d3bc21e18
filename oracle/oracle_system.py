"""Python face of the CPU oracle (oracle/colq_oracle.c): ``OracleDataSystem`` has the same ``register`` / ``execute``
surface as the reference's ``DataSystemSerialIndices`` and as ``colq.DataSystemColq``, so one test body (the TCK the
reference wishes for, README.md:149-153) runs against both.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs -- never by the
product package.
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
import time
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
sys.path.insert(0, str(ORACLE_DIR.parent / "java-columnar-query-engine_b200"))

from colq.data_system import (BitSet, Criteria, IntPredicate, Query, QueryResult, StringPredicate, Table)  # noqa: E402
from colq.in_memory import AssociationColumn, BooleanColumn, IntegerColumn, StringColumn  # noqa: E402

SUCCESS, FAILURE, THROW_INDEX_OOB, THROW_NULL, THROW_ILLEGAL_STATE, THROW_ILLEGAL_ARG = range(6)

_lib = None


def build() -> Path:
    so = ORACLE_DIR / "liboracle.so"
    src = ORACLE_DIR / "colq_oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return so


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    lib = C.CDLL(str(build()))
    p, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    sig = {
        "orc_system_new": (p, []),
        "orc_system_free": (None, [p]),
        "orc_last_message": (C.c_char_p, [p]),
        "orc_table_new": (C.c_int, [p]),
        "orc_table_add_ints": (C.c_int, [p, C.c_int, p, i64]),
        "orc_table_add_strings": (C.c_int, [p, C.c_int, p, p, i64]),
        "orc_table_add_bools": (C.c_int, [p, C.c_int, p, i64]),
        "orc_table_associate": (C.c_int, [p, C.c_int, C.c_int, p, p, p, i64, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "orc_table_size": (i64, [p, C.c_int]),
        "orc_table_width": (C.c_int, [p, C.c_int]),
        "orc_register": (None, [p, C.c_char_p, C.c_int]),
        "orc_query_new": (p, [C.c_char_p]),
        "orc_query_free": (None, [p]),
        "orc_query_create_child": (C.c_int, [p, C.c_int, C.c_int]),
        "orc_query_add_int_range": (None, [p, C.c_int, C.c_int, i32, i32]),
        "orc_query_add_str": (None, [p, C.c_int, C.c_int, C.c_int, p, i32]),
        "orc_query_add_bool": (None, [p, C.c_int, C.c_int, C.c_int, C.c_int]),
        "orc_execute": (C.c_int, [p, p, C.c_int, C.POINTER(p), C.POINTER(i64), C.POINTER(p), C.POINTER(i64)]),
        "orc_last_node_cardinalities": (C.c_int, [p, C.POINTER(i64), C.c_int]),
        "orc_free": (None, [p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleDataSystem:
    """CPU restatement of DataSystemSerialIndices over the colq host model."""

    def __init__(self, n_threads: int = 1):
        self.lib = load()
        self.sys = C.c_void_p(self.lib.orc_system_new())
        self.n_threads = n_threads
        self._tables: Dict[str, Table] = {}
        self._handles: Dict[int, int] = {}
        self._uploaded: Dict[int, int] = {}
        self._pins: List[Table] = []
        self._registered: Dict[str, int] = {}
        self.last_indices: Optional[np.ndarray] = None
        self.last_words: Optional[np.ndarray] = None
        self.last_execute_seconds = 0.0   # wall time of the last orc_execute call alone (tables already resident)

    def register(self, table_name: str, table: Table, **_placement) -> None:
        self._tables[table_name] = table
        self._pins.append(table)

    def _sync_tables(self) -> None:
        todo = list(self._tables.values())
        seen: Dict[int, Table] = {}
        while todo:
            t = todo.pop()
            if id(t) in seen:
                continue
            seen[id(t)] = t
            for c in t.columns():
                if isinstance(c, AssociationColumn):
                    todo.append(c.associated_entity)
        for tid, t in seen.items():
            if tid not in self._handles:
                self._handles[tid] = self.lib.orc_table_new(self.sys)
                self._uploaded[tid] = 0
                self._pins.append(t)
        # Columns must be appended in ordinal order per table; an association appends to BOTH tables, so walk all
        # tables round-robin and add whichever column is next for its table once its prerequisites exist.
        progress = True
        while progress:
            progress = False
            for tid, t in seen.items():
                h = self._handles[tid]
                cols = t.columns()
                while self._uploaded[tid] < len(cols):
                    ordinal = self._uploaded[tid]
                    c = cols[ordinal]
                    if isinstance(c, IntegerColumn):
                        v = c.ints()
                        self.lib.orc_table_add_ints(self.sys, h, _ptr(v), v.shape[0])
                    elif isinstance(c, StringColumn):
                        self.lib.orc_table_add_strings(self.sys, h, _ptr(c.offsets), _ptr(c.data), c.height())
                    elif isinstance(c, BooleanColumn):
                        v = c.bools()
                        self.lib.orc_table_add_bools(self.sys, h, _ptr(v), v.shape[0])
                    elif isinstance(c, AssociationColumn):
                        if not c.is_forward():
                            break  # appended when its forward column is associated
                        y = c.associated_entity
                        yid = id(y)
                        rev = c.reverse_associated_column()
                        y_ordinal = next(i for i, yc in enumerate(y.columns()) if yc is rev)
                        want_y = self._uploaded[yid] + (1 if yid == tid else 0)
                        if want_y != y_ordinal:
                            break  # the other table is not there yet
                        kind, offsets, targets = c.csr()
                        xo, yo = C.c_int(), C.c_int()
                        rc = self.lib.orc_table_associate(self.sys, h, self._handles[yid], _ptr(kind), _ptr(offsets),
                                                          _ptr(targets) if targets.size else None, kind.shape[0],
                                                          C.byref(xo), C.byref(yo))
                        if rc != SUCCESS:
                            raise TypeError(self.lib.orc_last_message(self.sys).decode())
                        assert xo.value == ordinal and yo.value == y_ordinal, (xo.value, ordinal, yo.value, y_ordinal)
                        self._uploaded[yid] += 1
                        if yid == tid:
                            self._uploaded[tid] += 1
                            progress = True
                            continue
                    self._uploaded[tid] += 1
                    progress = True
        for tid, t in seen.items():
            assert self._uploaded[tid] == len(t.columns()), "could not order the association columns"
        for name, t in self._tables.items():
            h = self._handles[id(t)]
            if self._registered.get(name) != h:
                self.lib.orc_register(self.sys, name.encode(), h)
                self._registered[name] = h

    def _translate(self, query: Query):
        q = C.c_void_p(self.lib.orc_query_new(query.table_name.encode()))
        stack = [(query.root_node, 0)]
        while stack:
            node, nid = stack.pop()
            for crit in node.get_criteria():
                if isinstance(crit, Criteria.IntCriteria):
                    p = crit.integer_predicate
                    assert isinstance(p, IntPredicate), "the C oracle evaluates structured predicates"
                    self.lib.orc_query_add_int_range(q, nid, crit.ordinal, p.lo, p.hi)
                elif isinstance(crit, Criteria.BooleanCriteria):
                    p = crit.boolean_predicate
                    self.lib.orc_query_add_bool(q, nid, crit.ordinal, int(bool(p(False))), int(bool(p(True))))
                else:
                    p = crit.string_predicate
                    assert isinstance(p, StringPredicate), "the C oracle evaluates structured predicates"
                    buf = (C.c_uint8 * max(len(p.needle), 1)).from_buffer_copy(p.needle or b"\0")
                    self.lib.orc_query_add_str(q, nid, crit.ordinal, p.op, buf, len(p.needle))
            for ordinal, child in node.get_children_by_ordinal().items():
                cid = self.lib.orc_query_create_child(q, nid, ordinal)
                stack.append((child, cid))
        return q

    def execute(self, query: Query):
        if query.table_name not in self._tables:
            return QueryResult.Failure(f"The query targets the table '{query.table_name}' but that table is not registered")
        table = self._tables[query.table_name]
        self._sync_tables()
        q = self._translate(query)
        words, idx = C.c_void_p(), C.c_void_p()
        nwords, count = C.c_int64(), C.c_int64()
        try:
            t0 = time.perf_counter()
            rc = self.lib.orc_execute(self.sys, q, self.n_threads, C.byref(words), C.byref(nwords), C.byref(idx), C.byref(count))
            self.last_execute_seconds = time.perf_counter() - t0
            msg = self.lib.orc_last_message(self.sys).decode()
            if rc == FAILURE:
                return QueryResult.Failure(msg)
            if rc == THROW_INDEX_OOB:
                raise IndexError(msg)
            if rc != SUCCESS:
                raise RuntimeError(msg)
            self.last_indices = np.ctypeslib.as_array(C.cast(idx, C.POINTER(C.c_int32)), shape=(max(count.value, 1),))[: count.value].copy()
            self.last_words = np.ctypeslib.as_array(C.cast(words, C.POINTER(C.c_uint64)), shape=(max(nwords.value, 1),))[: nwords.value].copy()
        finally:
            self.lib.orc_query_free(q)
            if words:
                self.lib.orc_free(words)
            if idx:
                self.lib.orc_free(idx)
        return QueryResult.Success(table.subset(BitSet(self.last_words, table.size())))

    def node_cardinalities(self) -> List[int]:
        out = (C.c_int64 * 64)()
        n = self.lib.orc_last_node_cardinalities(self.sys, out, 64)
        return [out[i] for i in range(n)]

    def close(self) -> None:
        if self.sys:
            self.lib.orc_system_free(self.sys)
            self.sys = C.c_void_p()
