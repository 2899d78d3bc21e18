"""``DataSystemColq`` -- the B200 execution module behind the reference's ``DataSystem`` interface.

Mirrors ``DataSystemSerialIndices`` (``E = data-system-serial-indices-arrays/src/main/java/dgroomes/
data_system_serial_indices_arrays``): ``register(name, table)`` (E/DataSystemSerialIndices.java:27) and
``execute(query) -> QueryResult`` (:53).  The host side only (1) ships column arrays to HBM once, (2) translates
the ``Query`` tree into ``colq_query_*`` downcalls and (3) turns the returned row set into the result ``Table``
with the registered table's own ``subset`` -- everything else happens in ``libcolq.so`` on the GPU.

Two layers:
  * ``ColqContext``  -- thin object wrapper of the C ABI (also used by bench.py with device-resident buffers),
  * ``DataSystemColq`` -- the reference-facing class.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from .data_system import (BitSet, Criteria, DataSystem, IntPredicate, Query, QueryResult, StringPredicate, Table)
from .in_memory import AssociationColumn, BooleanColumn, IntegerColumn, StringColumn


class ColqError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


def _raise(status: int, message: str):
    """Map colq_status to the exception class the reference would throw (include/colq.h)."""
    if status == _ffi.THROW_INDEX_OOB:
        raise IndexError(message)                      # java.lang.IndexOutOfBoundsException
    if status == _ffi.THROW_NULL:
        raise TypeError("NullPointerException: " + message)
    if status == _ffi.THROW_ILLEGAL_ARG:
        raise ValueError(message)                      # java.lang.IllegalArgumentException
    raise ColqError(status, message)                   # IllegalStateException / device errors


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class ExecResult:
    count: int
    indices: Optional[np.ndarray]
    bitmask: Optional[np.ndarray]
    timing: _ffi.Timing


def encode_dictionary(col: StringColumn):
    """Dictionary-encode a string column: ``(codes int32[n], dict_offsets uint32[d+1], dict_bytes uint8[], values)``
    with the d distinct values in first-appearance order (what the Java shim does while it copies ``String[]`` into
    off-heap segments)."""
    n = col.height()
    off, data = col.offsets, col.data
    raw = data.tobytes()
    index: Dict[bytes, int] = {}
    codes = np.empty(n, dtype=np.int32)
    for i in range(n):
        b = raw[int(off[i]):int(off[i + 1])]
        code = index.get(b)
        if code is None:
            code = index[b] = len(index)
        codes[i] = code
    keys = list(index.keys())
    d_off = np.zeros(len(keys) + 1, dtype=np.uint32)
    if keys:
        np.cumsum([len(k) for k in keys], out=d_off[1:])
    d_bytes = np.frombuffer(b"".join(keys), dtype=np.uint8).copy() if keys else np.zeros(0, dtype=np.uint8)
    return codes, d_off, d_bytes, [k.decode("utf-8") for k in keys]


def accept_words(flags: Sequence[bool]) -> np.ndarray:
    """bool per dictionary entry -> java.util.BitSet words (bit d of word d >> 6)."""
    bits = np.asarray(flags, dtype=np.uint8)
    packed = np.packbits(bits, bitorder="little")
    out = np.zeros((bits.shape[0] + 63) // 64 * 8 or 8, dtype=np.uint8)
    out[: packed.shape[0]] = packed
    return out.view(np.uint64)


class ColqQuery:
    def __init__(self, ctx: "ColqContext", table_name: str):
        self.ctx = ctx
        self.handle = C.c_void_p()
        ctx._check(ctx.lib.colq_query_create(ctx.handle, table_name.encode(), C.byref(self.handle)))
        ctx._queries.add(self)

    def child(self, parent: int, ordinal: int) -> int:
        out = C.c_int()
        self.ctx._check(self.ctx.lib.colq_query_child(self.handle, parent, ordinal, C.byref(out)))
        return out.value

    def criteria_i32_range(self, node: int, ordinal: int, lo: int, hi: int) -> None:
        self.ctx._check(self.ctx.lib.colq_query_criteria_i32_range(self.handle, node, ordinal, lo, hi))

    def criteria_str(self, node: int, ordinal: int, op: int, needle: bytes) -> None:
        buf = (C.c_uint8 * max(len(needle), 1)).from_buffer_copy(needle or b"\0")
        self.ctx._check(self.ctx.lib.colq_query_criteria_str(self.handle, node, ordinal, op, buf, len(needle)))

    def criteria_str_accept(self, node: int, ordinal: int, words: np.ndarray, n_dict: int) -> None:
        w = np.ascontiguousarray(words, dtype=np.uint64)
        self.ctx._check(self.ctx.lib.colq_query_criteria_str_accept(self.handle, node, ordinal, _ptr(w), n_dict))

    def criteria_bool(self, node: int, ordinal: int, accept_false: bool, accept_true: bool) -> None:
        self.ctx._check(self.ctx.lib.colq_query_criteria_bool(self.handle, node, ordinal, int(bool(accept_false)), int(bool(accept_true))))

    def criteria_i32_accept(self, node: int, ordinal: int, words: np.ndarray, n_dict: int) -> None:
        w = np.ascontiguousarray(words, dtype=np.uint64)
        self.ctx._check(self.ctx.lib.colq_query_criteria_i32_accept(self.handle, node, ordinal, _ptr(w), n_dict))

    def set_option(self, option: int, value: int) -> None:
        self.ctx._check(self.ctx.lib.colq_query_set_option(self.handle, option, value))

    def execute(self, want_indices: bool = True, want_bitmask: bool = False, n_rows: int = 0,
                index_capacity: int = 1 << 20, pinned: bool = False) -> ExecResult:
        """``pinned``: the indices land in a pinned buffer owned by this query (what the Java shim's off-heap result
        segment is) and the returned array is a VIEW of it, valid until the next execute / close of the query."""
        return self.ctx._execute(self, want_indices, want_bitmask, n_rows, index_capacity, fetch_only=False, pinned=pinned)

    def execute_async(self) -> None:
        self.ctx._check(self.ctx.lib.colq_execute_async(self.ctx.handle, self.handle))

    def fetch(self, want_indices: bool = True, want_bitmask: bool = False, n_rows: int = 0,
              index_capacity: int = 1 << 20) -> ExecResult:
        return self.ctx._execute(self, want_indices, want_bitmask, n_rows, index_capacity, fetch_only=True)

    def profile(self) -> List[Tuple[str, float, int, int]]:
        stages = (_ffi.Stage * 64)()
        n = C.c_int()
        self.ctx._check(self.ctx.lib.colq_profile(self.handle, stages, 64, C.byref(n)))
        return [(stages[i].name.decode(), stages[i].ms, stages[i].rows, stages[i].bytes) for i in range(min(n.value, 64))]

    # -- result materialisation on the device (the value half of Table.subset, include/colq.h)
    def result_count(self) -> int:
        n = C.c_int64()
        self.ctx._check(self.ctx.lib.colq_result_count(self.ctx.handle, self.handle, C.byref(n)))
        return n.value

    def _result_fixed(self, fn, ordinal: int, dtype) -> np.ndarray:
        n = self.result_count()
        out = np.empty(max(n, 1), dtype=dtype)
        got = C.c_int64()
        self.ctx._check(fn(self.ctx.handle, self.handle, ordinal, _ptr(out), n, C.byref(got)))
        return out[: got.value]

    def result_i32(self, ordinal: int) -> np.ndarray:
        return self._result_fixed(self.ctx.lib.colq_result_i32, ordinal, np.int32)

    def result_bool(self, ordinal: int) -> np.ndarray:
        return self._result_fixed(self.ctx.lib.colq_result_bool, ordinal, np.uint8)

    def _result_var(self, fn, ordinal: int, off_dtype, elem_dtype) -> Tuple[np.ndarray, np.ndarray]:
        n, total = C.c_int64(), C.c_int64()
        st = fn(self.ctx.handle, self.handle, ordinal, None, 0, None, 0, C.byref(n), C.byref(total))   # size query
        if st not in (_ffi.OK, _ffi.ERR_CAPACITY):
            self.ctx._check(st)
        off = np.zeros(n.value + 1, dtype=off_dtype)
        data = np.empty(max(total.value, 1), dtype=elem_dtype)
        self.ctx._check(fn(self.ctx.handle, self.handle, ordinal, _ptr(off), n.value + 1, _ptr(data), total.value, C.byref(n), C.byref(total)))
        return off, data[: total.value]

    def result_str(self, ordinal: int) -> Tuple[np.ndarray, np.ndarray]:
        """(offsets uint32[count+1], UTF-8 bytes) of a string column at the matching rows."""
        return self._result_var(self.ctx.lib.colq_result_str, ordinal, np.uint32, np.uint8)

    def result_csr(self, ordinal: int) -> Tuple[np.ndarray, np.ndarray]:
        """(offsets int64[count+1], targets int32) of a stored to-many association column at the matching rows."""
        return self._result_var(self.ctx.lib.colq_result_csr, ordinal, np.int64, np.int32)

    def profile_hot(self) -> Tuple[str, float, int, int, int]:
        """(name, mean ms, rows, algorithmic bytes, samples) of the dominant launch since the last call (OPT_PROFILE=2)."""
        st = _ffi.Stage()
        n = C.c_int()
        self.ctx._check(self.ctx.lib.colq_profile_hot(self.handle, C.byref(st), C.byref(n)))
        return st.name.decode(), st.ms, st.rows, st.bytes, n.value

    def node_cardinalities(self) -> List[int]:
        out = (C.c_int64 * 64)()
        n = C.c_int()
        self.ctx._check(self.ctx.lib.colq_node_cardinalities(self.ctx.handle, self.handle, out, 64, C.byref(n)))
        return [out[i] for i in range(min(n.value, 64))]

    def close(self) -> None:
        if self.handle and self.ctx.handle:
            self.ctx.lib.colq_query_destroy(self.handle)
        self.handle = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class ColqContext:
    """Object wrapper of one ``colq_ctx``."""

    def __init__(self, device: int = 0):
        self.lib = _ffi.load()
        if self.lib.colq_abi_version() != _ffi.ABI_VERSION:
            raise RuntimeError("libcolq.so ABI version mismatch")
        self.handle = C.c_void_p()
        st = self.lib.colq_create(device, C.byref(self.handle))
        if st != _ffi.OK:
            raise ColqError(st, f"colq_create(device={device}) failed with status {st}: no usable sm_100 GPU "
                                "(libcolq has no CPU fallback)")
        self.device = device
        self._keepalive: List[object] = []
        self._host_buffers: Dict[int, int] = {}      # address of a host_alloc array -> pointer to free
        self._host_views: Dict[int, np.ndarray] = {}  # address of a host_column view -> its padded raw buffer
        self._queries = weakref.WeakSet()   # colq_destroy frees a context's queries: close them first

    # -- plumbing
    def last_error(self) -> str:
        return (self.lib.colq_last_error(self.handle) or b"").decode("utf-8", "replace")

    def _check(self, status: int) -> None:
        if status != _ffi.OK:
            _raise(status, self.last_error())

    def set_stream(self, cuda_stream: int) -> None:
        self._check(self.lib.colq_set_stream(self.handle, C.c_void_p(cuda_stream)))

    def synchronize(self) -> None:
        self._check(self.lib.colq_synchronize(self.handle))

    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        self._check(self.lib.colq_comm_unique_id(self.handle, buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes, n_ranks: int, rank: int) -> None:
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.lib.colq_comm_init(self.handle, buf, n_ranks, rank))

    # -- tables
    def table_create(self, n_rows: int, placement: int = _ffi.REPLICATED, global_row_base: int = 0) -> int:
        out = C.c_int32()
        self._check(self.lib.colq_table_create(self.handle, n_rows, placement, global_row_base, C.byref(out)))
        return out.value

    def register(self, name: str, table: int) -> None:
        self._check(self.lib.colq_register(self.handle, name.encode(), table))

    def table_destroy(self, table: int) -> None:
        self._check(self.lib.colq_table_destroy(self.handle, table))

    def col_i32(self, table: int, ordinal: int, values: np.ndarray) -> None:
        v = np.ascontiguousarray(values, dtype=np.int32)
        self._check(self.lib.colq_col_i32(self.handle, table, ordinal, _ptr(v), v.shape[0]))

    def col_str(self, table: int, ordinal: int, offsets: np.ndarray, data: np.ndarray) -> None:
        o = np.ascontiguousarray(offsets, dtype=np.uint32)
        d = np.ascontiguousarray(data, dtype=np.uint8)
        self._check(self.lib.colq_col_str(self.handle, table, ordinal, _ptr(o), _ptr(d), o.shape[0] - 1, d.shape[0]))

    def col_i32_dict(self, table: int, ordinal: int, codes: np.ndarray, dict_values: np.ndarray) -> None:
        c = np.ascontiguousarray(codes, dtype=np.int32)
        d = np.ascontiguousarray(dict_values, dtype=np.int32)
        self._check(self.lib.colq_col_i32_dict(self.handle, table, ordinal, _ptr(c), c.shape[0], _ptr(d), d.shape[0]))

    def col_i32_dict_host(self, table: int, ordinal: int, codes: np.ndarray, dict_values: np.ndarray,
                          capacity_bytes: Optional[int] = None, n: Optional[int] = None) -> None:
        n = codes.shape[0] if n is None else n
        cap = self._capacity(codes) if capacity_bytes is None else capacity_bytes
        d = np.ascontiguousarray(dict_values, dtype=np.int32)
        self._keepalive.append(codes)
        self._check(self.lib.colq_col_i32_dict_host(self.handle, table, ordinal, _ptr(codes), cap, n, _ptr(d), d.shape[0]))

    def col_str_dict(self, table: int, ordinal: int, codes: np.ndarray, dict_offsets: np.ndarray, dict_bytes: np.ndarray) -> None:
        c = np.ascontiguousarray(codes, dtype=np.int32)
        o = np.ascontiguousarray(dict_offsets, dtype=np.uint32)
        d = np.ascontiguousarray(dict_bytes, dtype=np.uint8)
        self._check(self.lib.colq_col_str_dict(self.handle, table, ordinal, _ptr(c), c.shape[0], _ptr(o), _ptr(d), o.shape[0] - 1, d.shape[0]))

    def col_str_dict_device(self, table: int, ordinal: int, codes_ptr: int, n: int, dict_offsets: np.ndarray, dict_bytes: np.ndarray,
                            keepalive=None) -> None:
        o = np.ascontiguousarray(dict_offsets, dtype=np.uint32)
        d = np.ascontiguousarray(dict_bytes, dtype=np.uint8)
        self._keepalive.append(keepalive)
        self._check(self.lib.colq_col_str_dict_device(self.handle, table, ordinal, C.c_void_p(codes_ptr), n, _ptr(o), _ptr(d),
                                                      o.shape[0] - 1, d.shape[0]))

    def col_str_dict_host(self, table: int, ordinal: int, codes: np.ndarray, dict_offsets: np.ndarray, dict_bytes: np.ndarray,
                          capacity_bytes: Optional[int] = None, n: Optional[int] = None) -> None:
        n = codes.shape[0] if n is None else n
        cap = self._capacity(codes) if capacity_bytes is None else capacity_bytes
        o = np.ascontiguousarray(dict_offsets, dtype=np.uint32)
        d = np.ascontiguousarray(dict_bytes, dtype=np.uint8)
        self._keepalive.append(codes)
        self._check(self.lib.colq_col_str_dict_host(self.handle, table, ordinal, _ptr(codes), cap, n, _ptr(o), _ptr(d),
                                                    o.shape[0] - 1, d.shape[0]))

    def col_bool(self, table: int, ordinal: int, values: np.ndarray) -> None:
        v = np.ascontiguousarray(values, dtype=np.uint8)
        self._check(self.lib.colq_col_bool(self.handle, table, ordinal, _ptr(v), v.shape[0]))

    def col_i32_device(self, table: int, ordinal: int, dev_ptr: int, n: int, keepalive=None) -> None:
        self._keepalive.append(keepalive)
        self._check(self.lib.colq_col_i32_device(self.handle, table, ordinal, C.c_void_p(dev_ptr), n))

    def col_str_device(self, table: int, ordinal: int, off_ptr: int, off_cap: int, bytes_ptr: int, bytes_cap: int, n: int,
                       n_bytes: int, keepalive=None) -> None:
        self._keepalive.append(keepalive)
        self._check(self.lib.colq_col_str_device(self.handle, table, ordinal, C.c_void_p(off_ptr), off_cap,
                                                 C.c_void_p(bytes_ptr), bytes_cap, n, n_bytes))

    # -- pinned host memory and host-resident columns (include/colq.h "Host-resident columns")
    def host_alloc(self, nbytes: int) -> np.ndarray:
        """A pinned, device-mapped off-heap buffer as a uint8 array (the Java shim wraps the same pointer in a
        MemorySegment).  Freed by ``host_free`` or when the context closes."""
        p = C.c_void_p()
        self._check(self.lib.colq_host_alloc(self.handle, int(nbytes), C.byref(p)))
        n = max(int(nbytes), 16)
        a = np.ctypeslib.as_array((C.c_uint8 * n).from_address(p.value))
        self._host_buffers[a.ctypes.data] = p.value
        return a

    def host_free(self, buf: np.ndarray) -> None:
        p = self._host_buffers.pop(buf.ctypes.data)
        self._check(self.lib.colq_host_free(self.handle, C.c_void_p(p)))

    def host_register(self, buf: np.ndarray) -> None:
        self._check(self.lib.colq_host_register(self.handle, C.c_void_p(buf.ctypes.data), buf.nbytes))

    def host_unregister(self, buf: np.ndarray) -> None:
        self._check(self.lib.colq_host_unregister(self.handle, C.c_void_p(buf.ctypes.data)))

    def host_column(self, values: np.ndarray, dtype, pad_bytes: int = 64) -> np.ndarray:
        """Copy ``values`` into a fresh pinned buffer padded for whole-line reads; returns the typed view (length n)."""
        v = np.ascontiguousarray(values, dtype=dtype)
        raw = self.host_alloc((v.nbytes + 15) // 16 * 16 + pad_bytes)
        raw[: v.nbytes] = v.view(np.uint8).reshape(-1)
        raw[v.nbytes:] = 0
        out = raw[: v.nbytes].view(dtype)
        self._host_views[out.ctypes.data] = raw
        return out

    def _capacity(self, a: np.ndarray) -> int:
        raw = self._host_views.get(a.ctypes.data)
        return int(raw.nbytes) if raw is not None else int(a.nbytes)

    def col_i32_host(self, table: int, ordinal: int, values: np.ndarray, capacity_bytes: Optional[int] = None, n: Optional[int] = None) -> None:
        """``values`` must live in pinned memory (``host_column`` / ``host_alloc`` / a pinned torch tensor)."""
        n = values.shape[0] if n is None else n
        cap = self._capacity(values) if capacity_bytes is None else capacity_bytes
        self._keepalive.append(values)
        self._check(self.lib.colq_col_i32_host(self.handle, table, ordinal, _ptr(values), cap, n))

    def col_str_host(self, table: int, ordinal: int, offsets: np.ndarray, data: np.ndarray, n: int, n_bytes: int,
                     offsets_capacity: Optional[int] = None, bytes_capacity: Optional[int] = None) -> None:
        oc = self._capacity(offsets) if offsets_capacity is None else offsets_capacity
        bc = self._capacity(data) if bytes_capacity is None else bytes_capacity
        self._keepalive.append((offsets, data))
        self._check(self.lib.colq_col_str_host(self.handle, table, ordinal, _ptr(offsets), oc, _ptr(data), bc, n, n_bytes))

    def associate_fk_host(self, x: int, x_ordinal: int, y: int, y_ordinal: int, fk: np.ndarray,
                          capacity_bytes: Optional[int] = None, n: Optional[int] = None) -> None:
        n = fk.shape[0] if n is None else n
        cap = self._capacity(fk) if capacity_bytes is None else capacity_bytes
        self._keepalive.append(fk)
        self._check(self.lib.colq_associate_fk_host(self.handle, x, x_ordinal, y, y_ordinal, _ptr(fk), cap, n))

    def associate_fk(self, x: int, x_ordinal: int, y: int, y_ordinal: int, fk: np.ndarray) -> None:
        f = np.ascontiguousarray(fk, dtype=np.int32)
        self._check(self.lib.colq_associate_fk(self.handle, x, x_ordinal, y, y_ordinal, _ptr(f), f.shape[0]))

    def associate_fk_device(self, x: int, x_ordinal: int, y: int, y_ordinal: int, dev_ptr: int, n: int, keepalive=None) -> None:
        self._keepalive.append(keepalive)
        self._check(self.lib.colq_associate_fk_device(self.handle, x, x_ordinal, y, y_ordinal, C.c_void_p(dev_ptr), n))

    def associate_csr(self, x: int, x_ordinal: int, y: int, y_ordinal: int, offsets: np.ndarray, targets: np.ndarray) -> None:
        o = np.ascontiguousarray(offsets, dtype=np.int64)
        t = np.ascontiguousarray(targets, dtype=np.int32)
        self._check(self.lib.colq_associate_csr(self.handle, x, x_ordinal, y, y_ordinal, _ptr(o), _ptr(t), o.shape[0] - 1,
                                                t.shape[0]))

    # -- ingest on the device (include/colq.h "Ingest on the device")
    def associate(self, x: int, x_ordinal: int, y: int, y_ordinal: int, offsets: np.ndarray, targets: np.ndarray) -> bool:
        """``x.associateTo(y, ...)`` from a CSR; the GPU validates it and picks dense to-one vs CSR.  Returns True when the
        column was stored as the dense to-one form."""
        o = np.ascontiguousarray(offsets, dtype=np.int64)
        t = np.ascontiguousarray(targets, dtype=np.int32)
        is_fk = C.c_int()
        self._check(self.lib.colq_associate(self.handle, x, x_ordinal, y, y_ordinal, _ptr(o), _ptr(t), o.shape[0] - 1, t.shape[0], C.byref(is_fk)))
        return bool(is_fk.value)

    def associate_device(self, x: int, x_ordinal: int, y: int, y_ordinal: int, off_ptr: int, tgt_ptr: int, n: int, nnz: int, keepalive=None) -> bool:
        self._keepalive.append(keepalive)
        is_fk = C.c_int()
        self._check(self.lib.colq_associate_device(self.handle, x, x_ordinal, y, y_ordinal, C.c_void_p(off_ptr), C.c_void_p(tgt_ptr), n, nnz, C.byref(is_fk)))
        return bool(is_fk.value)

    def col_str_encode(self, table: int, ordinal: int) -> int:
        """Dictionary-encode a registered plain string column in place, on the GPU; returns the number of distinct values."""
        n = C.c_int64()
        self._check(self.lib.colq_col_str_encode(self.handle, table, ordinal, C.byref(n)))
        return n.value

    def col_dict_str(self, table: int, ordinal: int) -> Tuple[np.ndarray, np.ndarray]:
        """(offsets uint32[n_dict+1], bytes) of a dictionary-encoded column's distinct values."""
        n, nb = C.c_int64(), C.c_int64()
        st = self.lib.colq_col_dict_str(self.handle, table, ordinal, None, 0, None, 0, C.byref(n), C.byref(nb))
        if st not in (_ffi.OK, _ffi.ERR_CAPACITY):
            self._check(st)
        off = np.zeros(n.value + 1, dtype=np.uint32)
        data = np.zeros(max(nb.value, 1), dtype=np.uint8)
        self._check(self.lib.colq_col_dict_str(self.handle, table, ordinal, _ptr(off), n.value + 1, _ptr(data), nb.value, C.byref(n), C.byref(nb)))
        return off, data[: nb.value]

    # -- cross-shard associations: global targets into a sharded table (include/colq.h)
    def table_partition(self, table: int, bounds: Sequence[int]) -> None:
        b = np.ascontiguousarray(bounds, dtype=np.int64)
        self._check(self.lib.colq_table_partition(self.handle, table, _ptr(b), b.shape[0] - 1))

    def associate_fk_global(self, x: int, x_ordinal: int, y: int, y_ordinal: int, fk: np.ndarray) -> None:
        f = np.ascontiguousarray(fk, dtype=np.int32)
        self._check(self.lib.colq_associate_fk_global(self.handle, x, x_ordinal, y, y_ordinal, _ptr(f), f.shape[0]))

    def associate_csr_global(self, x: int, x_ordinal: int, y: int, y_ordinal: int, offsets: np.ndarray, targets: np.ndarray) -> None:
        o = np.ascontiguousarray(offsets, dtype=np.int64)
        t = np.ascontiguousarray(targets, dtype=np.int32)
        self._check(self.lib.colq_associate_csr_global(self.handle, x, x_ordinal, y, y_ordinal, _ptr(o), _ptr(t), o.shape[0] - 1, t.shape[0]))

    def query(self, table_name: str) -> ColqQuery:
        return ColqQuery(self, table_name)

    def _execute(self, q: ColqQuery, want_indices: bool, want_bitmask: bool, n_rows: int, index_capacity: int,
                 fetch_only: bool, pinned: bool = False) -> ExecResult:
        fn = self.lib.colq_fetch if fetch_only else self.lib.colq_execute
        count = C.c_int64()
        timing = _ffi.Timing()
        bitmask = np.zeros((n_rows + 63) // 64, dtype=np.uint64) if want_bitmask else None
        cap = max(int(index_capacity), 1)
        while True:
            if want_indices and pinned:
                buf = getattr(q, "_pinned_idx", None)
                if buf is None or buf.shape[0] < cap:
                    if buf is not None:
                        self.host_free(q._pinned_raw)
                    q._pinned_raw = self.host_alloc(cap * 4 + 64)
                    q._pinned_idx = buf = q._pinned_raw[: cap * 4].view(np.int32)
                idx = buf
            else:
                idx = np.empty(cap, dtype=np.int32) if want_indices else None
            st = fn(self.handle, q.handle, _ptr(bitmask), 0 if bitmask is None else bitmask.shape[0], _ptr(idx),
                    0 if idx is None else cap, C.byref(count), C.byref(timing))
            if st == _ffi.ERR_CAPACITY and want_indices and count.value > cap:
                cap = int(count.value)  # the true count was reported: retry the copy with a big enough buffer
                fn = self.lib.colq_fetch
                continue
            self._check(st)
            break
        return ExecResult(count.value, None if idx is None else idx[: count.value], bitmask, timing)

    def close(self) -> None:
        if self.handle:
            for q in list(self._queries):
                q.close()
            self._keepalive.clear()
            self._host_views.clear()
            for p in list(self._host_buffers.values()):
                self.lib.colq_host_free(self.handle, C.c_void_p(p))
            self._host_buffers.clear()
            self.lib.colq_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class DataSystemColq(DataSystem):
    """The reference-facing engine: same two methods as ``DataSystemSerialIndices``."""

    def __init__(self, device: int = 0, lazy_fk: bool = True, context: Optional[ColqContext] = None,
                 options: Optional[Dict[int, int]] = None, residency: str = "device", dictionary: bool = False,
                 materialize: str = "host", ingest: str = "host"):
        """``residency``: "device" copies every column to HBM at the first ``execute`` (default); "host" keeps int,
        string and to-one association columns in pinned off-heap buffers that the kernels stream in place over PCIe
        (only what a query touches moves; fully scanned columns are promoted to HBM by that first scan)."""
        if residency not in ("device", "host"):
            raise ValueError(residency)
        self.residency = residency
        # dictionary=True: string columns are stored dictionary-encoded; every string criterion -- structured or an
        # opaque lambda like the reference's -- is evaluated per DISTINCT value and the GPU row scan tests code bits.
        # dictionary="all": integer columns too, so opaque IntPredicate lambdas run as well (worth it when values repeat)
        if dictionary not in (False, True, "all"):
            raise ValueError(dictionary)
        self.dictionary = dictionary
        # materialize="device": the result Table's int / string / stored association columns are gathered on the GPU
        # (colq_result_*) instead of the registered table's host-side subset
        if materialize not in ("host", "device"):
            raise ValueError(materialize)
        self.materialize = materialize
        # ingest="device": the shim ships flat arrays only -- string columns are dictionary-encoded by the GPU
        # (colq_col_str_encode, when dictionary is set) and every association goes up as a CSR that the GPU validates and
        # classifies into dense to-one vs to-many (colq_associate); ingest="host": the round-1 host loops
        if ingest not in ("host", "device"):
            raise ValueError(ingest)
        self.ingest = ingest
        self._dict_values: Dict[Tuple[int, int], List[str]] = {}   # (id(table), ordinal) -> distinct values
        self.ctx = context or ColqContext(device)
        self.lazy_fk = lazy_fk
        self.options = dict(options or {})   # colq_option -> value, applied to every query
        self._tables: Dict[str, Table] = {}                 # private final Map<String, Table> tables (E/...:18)
        self._placement: Dict[int, Tuple[int, int]] = {}    # id(table) -> (placement, global_row_base)
        self._handles: Dict[int, int] = {}                  # id(table) -> colq_table
        self._uploaded: Dict[int, int] = {}                 # id(table) -> number of columns already on the device
        self._registered_handle: Dict[str, int] = {}
        self._pins: List[Table] = []
        self.last_timing: Optional[_ffi.Timing] = None
        self.last_query: Optional[ColqQuery] = None

    def register(self, table_name: str, table: Table, placement: int = _ffi.REPLICATED, global_row_base: int = 0) -> None:
        """E/DataSystemSerialIndices.java:27-29.  Like the reference this only records the reference: the app registers
        tables BEFORE ``associateTo`` appends their association columns (app/.../Runner.java:107 vs :138), so the
        device upload happens at the first ``execute`` by walking the table graph."""
        self._tables[table_name] = table
        self._placement[id(table)] = (placement, global_row_base)
        self._pins.append(table)

    # -- ingest: Table graph -> HBM
    def _sync_tables(self) -> None:
        # discover every table reachable through association columns (identity-keyed, cycle-safe)
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
                placement, base = self._placement.get(tid, (_ffi.REPLICATED, 0))
                self._handles[tid] = self.ctx.table_create(t.size(), placement, base)
                self._uploaded[tid] = 0
                self._pins.append(t)
        # scalar columns first, then associations (both ends must exist)
        for tid, t in seen.items():
            h = self._handles[tid]
            cols = t.columns()
            for ordinal in range(self._uploaded[tid], len(cols)):
                c = cols[ordinal]
                host = self.residency == "host" and t.size() > 0
                if isinstance(c, IntegerColumn) and self.dictionary == "all":
                    values, codes = np.unique(c.ints(), return_inverse=True)
                    self._dict_values[(tid, ordinal)] = [int(v) for v in values]
                    if host:
                        self.ctx.col_i32_dict_host(h, ordinal, self.ctx.host_column(codes.astype(np.int32), np.int32), values)
                    else:
                        self.ctx.col_i32_dict(h, ordinal, codes.astype(np.int32), values)
                elif isinstance(c, IntegerColumn):
                    if host:
                        self.ctx.col_i32_host(h, ordinal, self.ctx.host_column(c.ints(), np.int32))
                    else:
                        self.ctx.col_i32(h, ordinal, c.ints())
                elif isinstance(c, StringColumn) and self.dictionary and self.ingest == "device":
                    if host:
                        off = self.ctx.host_column(c.offsets, np.uint32)
                        dat = self.ctx.host_column(c.data, np.uint8)
                        self.ctx.col_str_host(h, ordinal, off, dat, t.size(), int(c.data.shape[0]))
                    else:
                        self.ctx.col_str(h, ordinal, c.offsets, c.data)
                    self.ctx.col_str_encode(h, ordinal)
                    d_off, d_bytes = self.ctx.col_dict_str(h, ordinal)
                    self._dict_values[(tid, ordinal)] = StringColumn(offsets=d_off, data=d_bytes).strings()
                elif isinstance(c, StringColumn) and self.dictionary:
                    codes, d_off, d_bytes, values = encode_dictionary(c)
                    self._dict_values[(tid, ordinal)] = values
                    if host:
                        self.ctx.col_str_dict_host(h, ordinal, self.ctx.host_column(codes, np.int32), d_off, d_bytes)
                    else:
                        self.ctx.col_str_dict(h, ordinal, codes, d_off, d_bytes)
                elif isinstance(c, StringColumn):
                    if host:
                        off = self.ctx.host_column(c.offsets, np.uint32)
                        dat = self.ctx.host_column(c.data, np.uint8)
                        self.ctx.col_str_host(h, ordinal, off, dat, t.size(), int(c.data.shape[0]))
                    else:
                        self.ctx.col_str(h, ordinal, c.offsets, c.data)
                elif isinstance(c, BooleanColumn):
                    self.ctx.col_bool(h, ordinal, c.bools())
        for tid, t in seen.items():
            h = self._handles[tid]
            cols = t.columns()
            for ordinal in range(self._uploaded[tid], len(cols)):
                c = cols[ordinal]
                if isinstance(c, AssociationColumn) and c.is_forward():
                    y = c.associated_entity
                    rev = c.reverse_associated_column()
                    y_ordinal = next(i for i, yc in enumerate(y.columns()) if yc is rev)
                    if self.ingest == "device":
                        _kind, offsets, targets = c.csr()
                        self.ctx.associate(h, ordinal, self._handles[id(y)], y_ordinal, offsets, targets)
                        continue
                    fk = c.fk()
                    if fk is not None and self.residency == "host" and t.size() > 0:
                        self.ctx.associate_fk_host(h, ordinal, self._handles[id(y)], y_ordinal, self.ctx.host_column(fk, np.int32))
                    elif fk is not None:
                        self.ctx.associate_fk(h, ordinal, self._handles[id(y)], y_ordinal, fk)
                    else:
                        _kind, offsets, targets = c.csr()
                        self.ctx.associate_csr(h, ordinal, self._handles[id(y)], y_ordinal, offsets, targets)
        for tid, t in seen.items():
            self._uploaded[tid] = len(t.columns())
        for name, t in self._tables.items():
            h = self._handles[id(t)]
            if self._registered_handle.get(name) != h:
                self.ctx.register(name, h)
                self._registered_handle[name] = h

    # -- Query -> colq_query
    def _translate(self, query: Query) -> Tuple[Optional[ColqQuery], Optional[str]]:
        cq = self.ctx.query(query.table_name)
        cq.set_option(_ffi.OPT_LAZY_FK, 1 if self.lazy_fk else 0)
        for opt, val in self.options.items():
            cq.set_option(opt, val)
        stack = [(query.root_node, 0, self._tables[query.table_name])]
        while stack:
            node, nid, node_table = stack.pop()
            for crit in node.get_criteria():
                if isinstance(crit, Criteria.IntCriteria):
                    p = crit.integer_predicate
                    values = self._dict_values.get((id(node_table), crit.ordinal)) if node_table is not None else None
                    if not isinstance(p, IntPredicate) and values is not None and callable(p):
                        # the reference's opaque IntPredicate (DS/Criteria.java:19): run it once per distinct value
                        cq.criteria_i32_accept(nid, crit.ordinal, accept_words([bool(p(v)) for v in values]), len(values))
                        continue
                    if not isinstance(p, IntPredicate):
                        cq.close()
                        return None, ("The criterion on ordinal %d is an opaque IntPredicate lambda; the GPU engine only runs "
                                      "structured predicates (colq.data_system.int_range & co.) and has no CPU fallback." % crit.ordinal)
                    cq.criteria_i32_range(nid, crit.ordinal, p.lo, p.hi)
                elif isinstance(crit, Criteria.StringCriteria):
                    p = crit.string_predicate
                    values = self._dict_values.get((id(node_table), crit.ordinal)) if node_table is not None else None
                    if not isinstance(p, StringPredicate) and values is not None and callable(p):
                        # the reference's opaque Predicate<String> (DS/Criteria.java:17): run it once per distinct value
                        cq.criteria_str_accept(nid, crit.ordinal, accept_words([bool(p(v)) for v in values]), len(values))
                        continue
                    if not isinstance(p, StringPredicate):
                        cq.close()
                        return None, ("The criterion on ordinal %d is an opaque Predicate<String> lambda; the GPU engine only runs "
                                      "structured predicates (colq.data_system.str_equals & co.) over plain string columns and has no CPU "
                                      "fallback; construct DataSystemColq(dictionary=True) to run opaque string predicates per distinct value." % crit.ordinal)
                    cq.criteria_str(nid, crit.ordinal, p.op, p.needle)
                elif isinstance(crit, Criteria.BooleanCriteria):
                    # a Predicate<Boolean> has two inputs: evaluate the lambda on both, the GPU tests the truth table
                    p = crit.boolean_predicate
                    cq.criteria_bool(nid, crit.ordinal, bool(p(False)), bool(p(True)))
                else:
                    raise TypeError(f"not a Criteria: {crit!r}")
            for ordinal, child in node.get_children_by_ordinal().items():
                child_table = None
                if node_table is not None and 0 <= ordinal < len(node_table.columns()):
                    col = node_table.columns()[ordinal]
                    if isinstance(col, AssociationColumn):
                        child_table = col.associated_entity
                stack.append((child, cq.child(nid, ordinal), child_table))
        return cq, None

    def execute(self, query: Query):
        """E/DataSystemSerialIndices.java:53-102."""
        if query is None:
            raise TypeError("NullPointerException: The 'query' argument must not be null")  # E/Verifier.java:41
        if query.table_name not in self._tables:  # (:54-57)
            return QueryResult.Failure(f"The query targets the table '{query.table_name}' but that table is not registered")
        table = self._tables[query.table_name]
        self._sync_tables()
        cq, why = self._translate(query)
        if cq is None:
            return QueryResult.Failure(why)
        try:
            res = cq.execute(want_indices=True, want_bitmask=False, n_rows=table.size())
        except ColqError as e:
            cq.close()
            if e.status == _ffi.FAILURE:
                return QueryResult.Failure(str(e))
            raise
        self.last_timing = res.timing
        if self.last_query is not None:
            self.last_query.close()
        self.last_query = cq
        placement, base = self._placement.get(id(table), (_ffi.REPLICATED, 0))
        if placement == _ffi.SHARDED:
            # with a communicator colq_execute returns ALL ranks' global indices (include/colq.h); the result Table of
            # this rank is built from its own rows only
            mine = res.indices[(res.indices >= base) & (res.indices < base + table.size())]
            local = mine - base
        else:
            local = res.indices
        matching_rows = BitSet.from_indices(local, table.size())
        if self.materialize == "device" and placement != _ffi.SHARDED:
            return QueryResult.Success(self._materialize_on_device(table, cq, local))
        # table.subset(executionContext.matchingRows()) (:100): the registered table builds the result itself
        return QueryResult.Success(table.subset(matching_rows))

    def _materialize_on_device(self, table: Table, cq: ColqQuery, rows: np.ndarray) -> Table:
        """M/InMemoryTable.java:106-159 with the per-column row copies done by the GPU (colq_result_*); columns the
        engine holds no data for (the reverse side of an association) are subset by the host table's own column."""
        from .in_memory import InMemoryTable
        out = []
        idx = rows.astype(np.int64)
        for ordinal, c in enumerate(table.columns()):
            if isinstance(c, IntegerColumn):
                out.append(IntegerColumn(cq.result_i32(ordinal)))
            elif isinstance(c, StringColumn):
                off, data = cq.result_str(ordinal)
                out.append(StringColumn(offsets=off, data=data))
            elif isinstance(c, BooleanColumn):
                out.append(BooleanColumn(cq.result_bool(ordinal).astype(bool)))
            else:
                out.append(c.take(idx))   # association columns keep their host form (indices un-remapped, no reverse link)
        return InMemoryTable(out)

    def close(self) -> None:
        if self.last_query is not None:
            self.last_query.close()
            self.last_query = None
        self.ctx.close()
