"""Host-side mirror of the reference's ``data-system`` API (the drop-in boundary).

Same names, argument meaning and error behaviour as the Java interfaces so that the parity tests read like
the reference's own ``QueryTest``.  Citations are relative to the reference checkout, with
``DS = data-system/src/main/java/dgroomes/data_system``.

The one deliberate difference: the reference's criteria carry opaque Java lambdas
(``Predicate<String>`` / ``IntPredicate``, DS/Criteria.java:17-19) which no GPU can run.  This module ships
*structured* predicate objects that are still plain callables (so they fit an unchanged ``Criteria`` record and
work on the CPU oracle), and the engine recognises them by type; an opaque callable yields
``QueryResult.Failure`` -- there is no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Sequence, Union

import numpy as np

INT_MIN = -(2 ** 31)
INT_MAX = 2 ** 31 - 1


# --------------------------------------------------------------------------------------- Association
class Association:
    """DS/Association.java:6-52 -- ``None | One(idx) | Many(indices)`` with ``add``."""

    @staticmethod
    def to_none() -> "AssociationNone":
        return NONE

    @staticmethod
    def to_one(idx: int) -> "One":
        return One(int(idx))

    @staticmethod
    def to_many(*indices: int) -> "Many":
        return Many(tuple(int(i) for i in indices))

    def add(self, idx: int) -> "Association":  # pragma: no cover - overridden
        raise NotImplementedError

    def targets(self) -> tuple:
        raise NotImplementedError


class AssociationNone(Association):
    def add(self, idx: int) -> Association:  # DS/Association.java:32-34
        return One(int(idx))

    def targets(self) -> tuple:
        return ()

    def __repr__(self) -> str:
        return "None"


@dataclass(frozen=True)
class One(Association):
    idx: int

    def add(self, idx: int) -> Association:  # DS/Association.java:39-41
        return Many((self.idx, int(idx)))

    def targets(self) -> tuple:
        return (self.idx,)


@dataclass(frozen=True)
class Many(Association):
    indices: tuple

    def add(self, idx: int) -> Association:  # DS/Association.java:46-50
        return Many(tuple(self.indices) + (int(idx),))

    def targets(self) -> tuple:
        return tuple(self.indices)


NONE = AssociationNone()


# --------------------------------------------------------------------------------------- structured predicates
class IntPredicate:
    """A structured ``java.util.function.IntPredicate``: the closed interval ``lo <= v <= hi``."""

    def __init__(self, lo: int, hi: int):
        self.lo = max(int(lo), INT_MIN)
        self.hi = min(int(hi), INT_MAX)

    def __call__(self, v: int) -> bool:
        return self.lo <= v <= self.hi

    test = __call__

    def __repr__(self) -> str:
        return f"IntRange[{self.lo}, {self.hi}]"


def int_range(lo: int, hi: int) -> IntPredicate:
    """``i -> lo <= i && i <= hi`` (closed)."""
    return IntPredicate(lo, hi)


def int_between_exclusive(lo: int, hi: int) -> IntPredicate:
    """``pop -> pop > lo && pop < hi`` (QueryTest.java:89)."""
    return IntPredicate(lo + 1, hi - 1)


def int_half_open(lo: int, hi: int) -> IntPredicate:
    """``i -> i >= lo && i < hi`` (app/.../Runner.java:231)."""
    return IntPredicate(lo, hi - 1)


def int_greater_than(x: int) -> IntPredicate:
    """``i -> i > x`` (QueryTest.java:43)."""
    return IntPredicate(x + 1, INT_MAX)


def int_less_than(x: int) -> IntPredicate:
    return IntPredicate(INT_MIN, x - 1)


def int_equals(x: int) -> IntPredicate:
    return IntPredicate(x, x)


# operator codes: the colq_str_op enum of include/colq.h
STR_EQ, STR_CONTAINS, STR_CMP_GT, STR_CMP_LT, STR_CMP_GE, STR_CMP_LE, STR_NE, STR_STARTS_WITH, STR_ENDS_WITH = range(9)


def _java_compare_to(a: str, b: str) -> int:
    """``String.compareTo``: UTF-16 code-unit order."""
    ea, eb = a.encode("utf-16-be", "surrogatepass"), b.encode("utf-16-be", "surrogatepass")
    return (ea > eb) - (ea < eb)


class StringPredicate:
    """A structured ``Predicate<String>``: one operator and one constant."""

    _IMPL = {
        STR_EQ: lambda s, x: s == x,
        STR_NE: lambda s, x: s != x,
        STR_CONTAINS: lambda s, x: x in s,
        STR_CMP_GT: lambda s, x: _java_compare_to(s, x) > 0,
        STR_CMP_LT: lambda s, x: _java_compare_to(s, x) < 0,
        STR_CMP_GE: lambda s, x: _java_compare_to(s, x) >= 0,
        STR_CMP_LE: lambda s, x: _java_compare_to(s, x) <= 0,
        STR_STARTS_WITH: lambda s, x: s.startswith(x),
        STR_ENDS_WITH: lambda s, x: s.endswith(x),
    }

    def __init__(self, op: int, value: str):
        self.op = int(op)
        self.value = value
        self.needle = value.encode("utf-8")

    def __call__(self, s: str) -> bool:
        return bool(self._IMPL[self.op](s, self.value))

    test = __call__

    def __repr__(self) -> str:
        return f"StringPredicate(op={self.op}, {self.value!r})"


def str_equals(x: str) -> StringPredicate:
    """``"X"::equals`` (Runner.java:236)."""
    return StringPredicate(STR_EQ, x)


def str_contains(x: str) -> StringPredicate:
    """``s -> s.contains("X")`` (Runner.java:255)."""
    return StringPredicate(STR_CONTAINS, x)


def str_compare_gt(x: str) -> StringPredicate:
    """``s -> s.compareTo("X") > 0`` (QueryTest.java:124)."""
    return StringPredicate(STR_CMP_GT, x)


def str_compare_lt(x: str) -> StringPredicate:
    """``s -> s.compareTo("X") < 0`` (QueryTest.java:125)."""
    return StringPredicate(STR_CMP_LT, x)


def str_compare_ge(x: str) -> StringPredicate:
    return StringPredicate(STR_CMP_GE, x)


def str_compare_le(x: str) -> StringPredicate:
    return StringPredicate(STR_CMP_LE, x)


def str_not_equals(x: str) -> StringPredicate:
    return StringPredicate(STR_NE, x)


def str_starts_with(x: str) -> StringPredicate:
    return StringPredicate(STR_STARTS_WITH, x)


def str_ends_with(x: str) -> StringPredicate:
    return StringPredicate(STR_ENDS_WITH, x)


# --------------------------------------------------------------------------------------- Criteria / Query / QueryResult
class Criteria:
    """DS/Criteria.java:10-20 (sealed: IntCriteria | StringCriteria).

    BooleanCriteria is the SURVEY.md 8(f4) extension: the reference declares BooleanColumnFilterable.where(Predicate<Boolean>)
    (DS/ColumnFilterable.java:20-22) but has no criterion for it and its Verifier refuses boolean columns
    (E/Verifier.java:82-84).  Int / string criteria on a boolean column still answer that Failure."""

    @dataclass(frozen=True)
    class StringCriteria:
        ordinal: int
        string_predicate: Callable[[str], bool]

    @dataclass(frozen=True)
    class IntCriteria:
        ordinal: int
        integer_predicate: Callable[[int], bool]

    @dataclass(frozen=True)
    class BooleanCriteria:
        ordinal: int
        boolean_predicate: Callable[[bool], bool]


class Query:
    """DS/Query.java:17-54."""

    class Node:
        def __init__(self) -> None:
            self._children_by_ordinal: Dict[int, "Query.Node"] = {}
            self._criteria: List[object] = []

        def create_child(self, ordinal: int) -> "Query.Node":
            if ordinal in self._children_by_ordinal:  # DS/Query.java:33-35
                raise ValueError(f"A child already exists at ordinal {ordinal}")
            child = Query.Node()
            self._children_by_ordinal[ordinal] = child
            return child

        def get_children_by_ordinal(self) -> Dict[int, "Query.Node"]:
            return dict(self._children_by_ordinal)

        def add_criteria(self, criteria) -> "Query.Node":
            self._criteria.append(criteria)
            return self

        def get_criteria(self) -> List[object]:
            return self._criteria

    def __init__(self, table_name: str):
        self.table_name = table_name
        self.root_node = Query.Node()


class QueryResult:
    """DS/QueryResult.java:3-9."""

    @dataclass
    class Success:
        result_set: "Table"

    @dataclass
    class Failure:
        message: str


# --------------------------------------------------------------------------------------- Table / Column interfaces
class Column:
    """DS/Column.java:6-17."""

    def height(self) -> int:
        raise NotImplementedError


class Table:
    """DS/Table.java:15-35."""

    def columns(self) -> Sequence[Column]:
        raise NotImplementedError

    def width(self) -> int:
        return len(self.columns())

    def size(self) -> int:
        raise NotImplementedError

    def subset(self, matching_rows: "BitSet") -> "Table":
        raise NotImplementedError


class DataSystem:
    """DS/DataSystem.java:15-32."""

    def execute(self, query: Query) -> Union[QueryResult.Success, QueryResult.Failure]:
        raise NotImplementedError


# --------------------------------------------------------------------------------------- BitSet
class BitSet:
    """``java.util.BitSet`` over little-endian uint64 words: bit i <-> words[i >> 6] & (1 << (i & 63))."""

    def __init__(self, words: np.ndarray, nbits: int):
        self.words = np.ascontiguousarray(words, dtype=np.uint64)
        self.nbits = int(nbits)

    @staticmethod
    def from_indices(indices: np.ndarray, nbits: int) -> "BitSet":
        words = np.zeros((nbits + 63) // 64, dtype=np.uint64)
        idx = np.asarray(indices, dtype=np.int64)
        np.bitwise_or.at(words, idx >> 6, np.uint64(1) << (idx & 63).astype(np.uint64))
        return BitSet(words, nbits)

    def cardinality(self) -> int:
        return int(np.unpackbits(self.words.view(np.uint8)).sum())

    def to_indices(self) -> np.ndarray:
        bits = np.unpackbits(self.words.view(np.uint8), bitorder="little")
        return np.flatnonzero(bits[: self.nbits]).astype(np.int32)

    def get(self, i: int) -> bool:
        return bool((int(self.words[i >> 6]) >> (i & 63)) & 1)
