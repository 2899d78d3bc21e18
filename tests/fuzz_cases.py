"""Random schemas and query trees shared by the GPU fuzz suite (test_gpu_fuzz.py: libcolq.so vs the oracle) and the CPU
semantics suite (test_oracle_semantics.py: the oracle vs a row-object model of the Java engine).

Random tables with int and string columns, random to-one (with Nones) and to-many associations between them (self
associations included), and random query trees that walk forward and reverse association columns with criteria on any
level -- random instances of exactly the semantics of E/DataSystemSerialIndices.java:53-102.
"""
import numpy as np

from colq import Criteria, InMemoryTable, Query, int_range
from colq.data_system import StringPredicate
from colq.in_memory import IntegerColumn, StringColumn

SIZES = [1, 37, 1000, 5003, 20_000]
WORDS = ["", "a", "ab", "abc", "north", "South", "North Dakota", "é", "PLYMOUTH", "NEW PLYMOUTH", "x" * 40]


def make_case(seed, size_choices=None):
    """-> (build(data_system), [query factories]); `size_choices` replaces SIZES (the CPU suite uses small tables)."""
    rng = np.random.default_rng(seed)
    n_tables = int(rng.integers(2, 5))
    sizes = [int(rng.choice(size_choices or SIZES)) for _ in range(n_tables)]
    scalar = []   # per table: list of ("int", array) / ("str", list)
    for n in sizes:
        cols = [("int", rng.integers(-20, 20, size=n, dtype=np.int32))]
        if rng.random() < 0.7:
            cols.append(("str", [WORDS[i] for i in rng.integers(0, len(WORDS), size=n)]))
        if rng.random() < 0.4:
            cols.append(("int", rng.integers(0, 1000, size=n, dtype=np.int32)))
        scalar.append(cols)
    assocs = []   # (x, y, kind, payload)
    for _ in range(int(rng.integers(1, 5))):
        x, y = int(rng.integers(0, n_tables)), int(rng.integers(0, n_tables))
        if rng.random() < 0.6:
            assocs.append((x, y, "fk", rng.integers(-1, sizes[y], size=sizes[x], dtype=np.int32)))
        else:
            deg = rng.integers(0, 4, size=sizes[x])
            off = np.zeros(sizes[x] + 1, dtype=np.int64)
            np.cumsum(deg, out=off[1:])
            assocs.append((x, y, "csr", (off, rng.integers(0, sizes[y], size=int(off[-1]), dtype=np.int32))))
    # column layout per table after the associateTo calls: scalar columns, then association columns in call order
    layout = [[(k, (int(v.min()), int(v.max())) if k == "int" and len(v) else None) for k, v in cols] for cols in scalar]
    for x, y, _kind, _p in assocs:
        layout[x].append(("assoc", y))
        layout[y].append(("assoc", x))

    def build(ds):
        tables = []
        for cols in scalar:
            tables.append(InMemoryTable.of_columns(*[IntegerColumn(v) if k == "int" else StringColumn(v) for k, v in cols]))
        for x, y, kind, payload in assocs:
            if kind == "fk":
                tables[x].associate_to(tables[y], fk=payload)
            else:
                tables[x].associate_to(tables[y], csr=payload)
        for i, t in enumerate(tables):
            ds.register(f"t{i}", t)

    def random_query(qseed):
        def make():
            r = np.random.default_rng(qseed)
            root = int(r.integers(0, n_tables))
            q = Query(f"t{root}")
            budget = [int(r.integers(1, 7))]

            def fill(node, table, depth):
                for ordinal, (kind, target) in enumerate(layout[table]):
                    if kind == "int" and r.random() < 0.4:
                        vmin, vmax = target
                        lo = int(r.integers(vmin - 3, vmax + 1))
                        node.add_criteria(Criteria.IntCriteria(ordinal, int_range(lo, lo + int(r.integers(0, max(vmax - vmin, 1))))))
                    elif kind == "str" and r.random() < 0.4:
                        node.add_criteria(Criteria.StringCriteria(ordinal, StringPredicate(int(r.integers(0, 9)), WORDS[int(r.integers(0, len(WORDS)))])))
                    elif kind == "assoc" and depth < 3 and budget[0] > 0 and r.random() < 0.55:
                        budget[0] -= 1
                        fill(node.create_child(ordinal), target, depth + 1)

            fill(q.root_node, root, 0)
            return q
        return make

    return build, [random_query(seed * 1000 + i) for i in range(6)]
