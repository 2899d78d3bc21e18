"""Mirror of the reference's array-backed physical model (``data-model-in-memory``).

``M = data-model-in-memory/src/main/java/dgroomes/in_memory`` in the citations.  Same construction API
(``of_columns`` / ``of_ints`` / ``of_strings`` / ``associate_to`` / ``subset``) as M/InMemoryTable.java and
M/InMemoryColumn.java, but every column is held in a compact numpy form (int32 arrays, offsets+UTF-8 bytes,
kind/offsets/targets for ``Association[]``) so that the same objects scale to the 10k-universe workload and can
be handed to ``libcolq.so`` (and to the CPU oracle) without per-row Python objects.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .data_system import NONE, Association, BitSet, Column, Many, One, Table


def pack_strings(strings: Sequence[str]):
    blobs = [s.encode("utf-8") for s in strings]
    offsets = np.zeros(len(blobs) + 1, dtype=np.uint32)
    if blobs:
        np.cumsum([len(b) for b in blobs], out=offsets[1:])
    data = np.frombuffer(b"".join(blobs), dtype=np.uint8).copy()
    return offsets, data


class InMemoryColumn(Column):
    """M/InMemoryColumn.java:19 (sealed: Boolean | Integer | String | Association)."""

    @staticmethod
    def of_ints(*ints: int) -> "IntegerColumn":
        return IntegerColumn(np.array(ints, dtype=np.int32))

    @staticmethod
    def of_strings(*strings: str) -> "StringColumn":
        return StringColumn(list(strings))


class IntegerColumn(InMemoryColumn):
    """M/InMemoryColumn.java:46-62."""

    def __init__(self, ints):
        self._ints = np.ascontiguousarray(ints, dtype=np.int32)

    def ints(self) -> np.ndarray:
        return self._ints

    def height(self) -> int:
        return int(self._ints.shape[0])

    def take(self, idx: np.ndarray) -> "IntegerColumn":
        return IntegerColumn(self._ints[idx])


class BooleanColumn(InMemoryColumn):
    """M/InMemoryColumn.java:28-44 (``where`` is unimplemented in the reference; criteria on it fail)."""

    def __init__(self, bools):
        self._bools = np.ascontiguousarray(bools, dtype=np.uint8)

    def bools(self) -> np.ndarray:
        return self._bools

    def height(self) -> int:
        return int(self._bools.shape[0])

    def take(self, idx: np.ndarray) -> "BooleanColumn":
        return BooleanColumn(self._bools[idx])


class StringColumn(InMemoryColumn):
    """M/InMemoryColumn.java:64-80, stored as n+1 uint32 offsets + UTF-8 bytes."""

    def __init__(self, strings: Optional[Sequence[str]] = None, *, offsets=None, data=None):
        if strings is not None:
            self.offsets, self.data = pack_strings(strings)
        else:
            self.offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
            self.data = np.ascontiguousarray(data, dtype=np.uint8)

    def height(self) -> int:
        return int(self.offsets.shape[0]) - 1

    def get(self, i: int) -> str:
        return bytes(self.data[int(self.offsets[i]):int(self.offsets[i + 1])]).decode("utf-8")

    def strings(self) -> List[str]:
        return [self.get(i) for i in range(self.height())]

    def take(self, idx: np.ndarray) -> "StringColumn":
        idx = np.asarray(idx, dtype=np.int64)
        starts = self.offsets[idx].astype(np.int64)
        lens = self.offsets[idx + 1].astype(np.int64) - starts
        new_off = np.zeros(idx.shape[0] + 1, dtype=np.uint32)
        np.cumsum(lens, out=new_off[1:])
        total = int(new_off[-1])
        # gather byte ranges: position p of output row j reads data[starts[j] + p]
        if total:
            row_of = np.repeat(np.arange(idx.shape[0]), lens)
            within = np.arange(total) - np.repeat(new_off[:-1].astype(np.int64), lens)
            data = self.data[starts[row_of] + within]
        else:
            data = np.zeros(0, dtype=np.uint8)
        return StringColumn(offsets=new_off, data=data)


class AssociationColumn(InMemoryColumn):
    """M/InMemoryColumn.java:85-138, with ``Association[]`` flattened to (kind, offsets, targets).

    kind[i]: 0 None, 1 One, 2 Many.  ``fk`` is set when every row is None/One (dense to-one form, -1 = None).
    A reverse column created by ``associate_to`` is the transpose of its forward column and is only
    materialised on first access (the 10k-universe tables never need it on the host).
    """

    def __init__(self, associated_entity: "InMemoryTable", *, kind=None, offsets=None, targets=None, fk=None,
                 transpose_of: Optional["AssociationColumn"] = None, n_rows: Optional[int] = None):
        self.associated_entity = associated_entity
        self._reverse: Optional[AssociationColumn] = None
        self._transpose_of = transpose_of
        self._fk = None
        self._kind = self._offsets = self._targets = None
        if transpose_of is not None:
            self._n = int(n_rows)
        elif fk is not None:
            self._fk = np.ascontiguousarray(fk, dtype=np.int32)
            self._n = int(self._fk.shape[0])
        else:
            self._kind = np.ascontiguousarray(kind, dtype=np.uint8)
            self._offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            self._targets = np.ascontiguousarray(targets, dtype=np.int32)
            self._n = int(self._kind.shape[0])

    # -- construction helpers
    @staticmethod
    def from_associations(associated_entity, associations: Sequence[Association]) -> "AssociationColumn":
        kinds = np.zeros(len(associations), dtype=np.uint8)
        offsets = np.zeros(len(associations) + 1, dtype=np.int64)
        tg: List[int] = []
        for i, a in enumerate(associations):
            if a is None:
                raise RuntimeError("Found a null association")  # M/InMemoryTable.java:66
            if isinstance(a, One):
                kinds[i] = 1
            elif isinstance(a, Many):
                kinds[i] = 2
            tg.extend(a.targets())
            offsets[i + 1] = len(tg)
        return AssociationColumn(associated_entity, kind=kinds, offsets=offsets, targets=np.array(tg, dtype=np.int32))

    # -- Column
    def height(self) -> int:
        return self._n

    # -- AssociationColumn interface (DS/AssociationColumn.java:5-13)
    def set_reverse_associated_column(self, col: "AssociationColumn") -> None:
        if self._reverse is not None:  # M/InMemoryColumn.java:115-117
            raise RuntimeError("reverseAssociatedColumn is already set")
        self._reverse = col

    def reverse_associated_column(self) -> "AssociationColumn":
        if self._reverse is None:  # M/InMemoryColumn.java:123-125
            raise RuntimeError("reverseAssociatedColumn was never set")
        return self._reverse

    def associations_for_index(self, i: int) -> Association:
        kind, offsets, targets = self.csr()
        k = int(kind[i])
        if k == 0:
            return NONE
        t = targets[int(offsets[i]):int(offsets[i + 1])]
        return One(int(t[0])) if k == 1 else Many(tuple(int(v) for v in t))

    # -- compact views
    def is_forward(self) -> bool:
        return self._transpose_of is None

    def forward_column(self) -> "AssociationColumn":
        return self if self._transpose_of is None else self._transpose_of

    def fk(self) -> Optional[np.ndarray]:
        """Dense to-one form (-1 = None) or None if some row is ``Many``."""
        if self._fk is not None:
            return self._fk
        if self._transpose_of is not None and self._kind is None:
            self._materialise_transpose()
        if self._kind is not None and not (self._kind == 2).any():
            fk = np.full(self._n, -1, dtype=np.int32)
            ones = np.flatnonzero(self._kind == 1)
            fk[ones] = self._targets[self._offsets[ones]]
            return fk
        return None

    def csr(self):
        """(kind, offsets, targets) numpy arrays."""
        if self._kind is None:
            if self._fk is not None:
                valid = self._fk >= 0
                self._kind = valid.astype(np.uint8)
                self._offsets = np.zeros(self._n + 1, dtype=np.int64)
                np.cumsum(valid, out=self._offsets[1:])
                self._targets = self._fk[valid]
            else:
                self._materialise_transpose()
        return self._kind, self._offsets, self._targets

    def _materialise_transpose(self) -> None:
        # M/InMemoryTable.java:55-82: y -> [x...] with x ascending, None/One/Many by list length
        fkind, foff, ftgt = self._transpose_of.csr()
        lens = np.diff(foff)  # kind 0 rows own an empty range in every construction path
        xs = np.repeat(np.arange(fkind.shape[0], dtype=np.int32), lens)
        ys = ftgt
        order = np.argsort(ys, kind="stable")
        counts = np.bincount(ys, minlength=self._n).astype(np.int64)
        self._offsets = np.zeros(self._n + 1, dtype=np.int64)
        np.cumsum(counts, out=self._offsets[1:])
        self._targets = xs[order].astype(np.int32)
        self._kind = np.minimum(counts, 2).astype(np.uint8)

    def take(self, idx: np.ndarray) -> "AssociationColumn":
        """Row subset with UN-remapped indices and no reverse link (M/InMemoryTable.java:143-154)."""
        idx = np.asarray(idx, dtype=np.int64)
        if self._fk is not None:
            return AssociationColumn(self.associated_entity, fk=self._fk[idx])
        kind, offsets, targets = self.csr()
        lens = (offsets[idx + 1] - offsets[idx])
        new_off = np.zeros(idx.shape[0] + 1, dtype=np.int64)
        np.cumsum(lens, out=new_off[1:])
        total = int(new_off[-1])
        if total:
            row_of = np.repeat(np.arange(idx.shape[0]), lens)
            within = np.arange(total) - np.repeat(new_off[:-1], lens)
            tg = targets[offsets[idx][row_of] + within]
        else:
            tg = np.zeros(0, dtype=np.int32)
        return AssociationColumn(self.associated_entity, kind=kind[idx], offsets=new_off, targets=tg)


class InMemoryTable(Table):
    """M/InMemoryTable.java:16-160."""

    def __init__(self, columns: List[InMemoryColumn]):
        self._columns = columns

    def columns(self) -> List[InMemoryColumn]:
        return self._columns

    @staticmethod
    def of_columns(*columns: InMemoryColumn) -> "InMemoryTable":
        return InMemoryTable(list(columns))

    def associate_to(self, associated_entity: "InMemoryTable", *associations: Association, fk=None, csr=None) -> AssociationColumn:
        """``x.associateTo(y, associations...)`` (M/InMemoryTable.java:44-90).

        Besides the reference's ``Association...`` varargs the compact forms ``fk=int32[n]`` (-1 = None) and
        ``csr=(offsets, targets)`` are accepted for large tables.
        """
        if fk is not None:
            col = AssociationColumn(associated_entity, fk=fk)
            tg, bad_lo = col._fk, -1
        elif csr is not None:
            offsets, targets = csr
            offsets = np.asarray(offsets, dtype=np.int64)
            lens = np.diff(offsets)
            col = AssociationColumn(associated_entity, kind=np.minimum(lens, 2).astype(np.uint8), offsets=offsets, targets=targets)
            tg, bad_lo = col._targets, 0
        else:
            col = AssociationColumn.from_associations(associated_entity, associations)
            tg, bad_lo = col._targets, 0
        ysize = associated_entity.size()
        if tg.size and (int(tg.min()) < bad_lo or int(tg.max()) >= ysize):
            # yIndexToXAssociations.get(yIndex) is null for an index outside the associated table (:70-71)
            raise TypeError("NullPointerException: association target outside the associated table")
        self._columns.append(col)  # (:48)
        reverse = AssociationColumn(self, transpose_of=col, n_rows=ysize)
        reverse.set_reverse_associated_column(col)  # (:84)
        col.set_reverse_associated_column(reverse)  # (:85)
        associated_entity._columns.append(reverse)  # (:88)
        return col

    def size(self) -> int:
        return self._columns[0].height()  # M/InMemoryTable.java:92-101: length of column 0

    def subset(self, matching_rows: BitSet) -> "InMemoryTable":
        """M/InMemoryTable.java:106-159: every column pruned to the set bits, ascending row order."""
        idx = matching_rows.to_indices().astype(np.int64)
        return InMemoryTable([c.take(idx) for c in self._columns])
