"""colq -- host-side mirror of the reference's data-system API over libcolq.so (B200, sm_100a).

Importing this package never touches the GPU; ``DataSystemColq()`` / ``ColqContext()`` do, and fail loudly when
``lib/libcolq.so`` or an sm_100 device is missing (there is no CPU fallback in the product path).
"""
from .data_system import (NONE, Association, BitSet, Criteria, DataSystem, IntPredicate, Many, One, Query, QueryResult,
                          StringPredicate, Table, int_between_exclusive, int_equals, int_greater_than, int_half_open,
                          int_less_than, int_range, str_compare_ge, str_compare_gt, str_compare_le, str_compare_lt,
                          str_contains, str_ends_with, str_equals, str_not_equals, str_starts_with)
from .in_memory import (AssociationColumn, BooleanColumn, InMemoryColumn, InMemoryTable, IntegerColumn, StringColumn)

of_columns = InMemoryTable.of_columns
of_ints = InMemoryColumn.of_ints
of_strings = InMemoryColumn.of_strings


def __getattr__(name):  # lazy: engine pulls in ctypes + the shared library only when used
    if name in ("DataSystemColq", "ColqContext", "ColqQuery", "ColqError"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
