"""Test infrastructure: a row-object model of the reference engine, independent of oracle/colq_oracle.c.

The C oracle restates the Java engine over flat arrays and uint64 words.  This module restates it a second time the way
the Java itself is written -- one ``Association`` object per row fetched with ``associations_for_index`` from the REVERSE
column that ``associate_to`` cross-linked, predicates as callables over a row index, ``BitSet`` as a Python set of ints --
so that the two restatements can be compared on random schemas (tests/test_oracle_semantics.py).  It also carries the
declarative reading of the result, which is what the GPU planner's single post-order pass relies on:

    match(node, r) = AND(criteria of node on row r)  and  for every child c: exists r' in assoc_c(r): match(c, r')

Citations: E = data-system-serial-indices-arrays/src/main/java/dgroomes/data_system_serial_indices_arrays.
Pure-Python row loops: small tables only.  Imported by tests only.
"""
from collections import deque

import numpy as np

from colq import Criteria, QueryResult
from colq.data_system import BitSet
from colq.in_memory import AssociationColumn, BooleanColumn, IntegerColumn, StringColumn


class _Node:
    """E/ExecutionContext.java:38-58."""

    def __init__(self, table, parent, association_to_parent):
        self.table = table
        self.parent = parent
        self.association_to_parent = association_to_parent
        self.column_predicates = []
        self.child_nodes = []
        self.matching_bits = set()

    def create_child_node(self, association_to_child):  # E/ExecutionContext.java:64-68
        child = _Node(association_to_child.associated_entity, self, association_to_child.reverse_associated_column())
        self.child_nodes.append(child)
        return child

    def filter_self(self):  # E/ExecutionContext.java:79-94
        n = self.table.size()
        if not self.column_predicates:
            self.matching_bits = set(range(n))
            return
        self.matching_bits = {i for i in range(n) if all(p(i) for p in self.column_predicates)}

    def filter_parent(self):  # E/ExecutionContext.java:100-122
        if self.parent is None:
            return
        by_association = set()
        for i in range(self.table.size()):
            if i not in self.matching_bits:
                continue
            by_association.update(self.association_to_parent.associations_for_index(i).targets())
        self.parent.matching_bits &= by_association


def _where(column, criterion):
    """E/Verifier.java:71-90 + M/InMemoryColumn.java:53-56,71-74: (failure message | None, row-index predicate)."""
    if isinstance(column, StringColumn):
        if not isinstance(criterion, Criteria.StringCriteria):
            return "The column is a string column but the criterion is not a string predicate.", None
        strings = column.strings()
        return None, (lambda i, p=criterion.string_predicate: bool(p(strings[i])))
    if isinstance(column, IntegerColumn):
        if not isinstance(criterion, Criteria.IntCriteria):
            return "The column is an integer column but the criterion is not an integer predicate.", None
        ints = column.ints()
        return None, (lambda i, p=criterion.integer_predicate: bool(p(int(ints[i]))))
    if isinstance(column, BooleanColumn):
        if not isinstance(criterion, Criteria.BooleanCriteria):  # the reference stops here for every criterion
            return "Boolean columns are not supported yet.", None
        bools = column.bools()
        return None, (lambda i, p=criterion.boolean_predicate: bool(p(bool(bools[i]))))
    return "Association columns can't be matched on with a scalar criteria.", None


class JavaModelDataSystem:
    """E/DataSystemSerialIndices.java:14-102 over the colq host model, row objects and Python sets."""

    def __init__(self):
        self.tables = {}
        self.last_indices = None
        self.nodes_in_creation_order = []

    def register(self, table_name, table, **_placement):
        self.tables[table_name] = table

    def _verify(self, query, table):  # E/Verifier.java:40-111
        root = _Node(table, None, None)
        self.nodes_in_creation_order = [root]
        to_visit = deque([(query.root_node, root)])
        while to_visit:
            query_node, node = to_visit.popleft()
            columns = node.table.columns()
            for criterion in query_node.get_criteria():
                ordinal = criterion.ordinal
                if len(columns) < ordinal:  # sic (:62)
                    return f"The query ordinal '{ordinal}' is out of bounds for the table with {len(columns)} columns", None
                if ordinal < 0 or ordinal >= len(columns):
                    raise IndexError(f"Index {ordinal} out of bounds for length {len(columns)}")
                message, predicate = _where(columns[ordinal], criterion)
                if message:
                    return message, None
                node.column_predicates.append(predicate)
            for ordinal, child_query_node in query_node.get_children_by_ordinal().items():
                if ordinal < 0 or ordinal >= len(columns):
                    raise IndexError(f"Index {ordinal} out of bounds for length {len(columns)}")
                column = columns[ordinal]
                if not isinstance(column, AssociationColumn):
                    return (f"The column at ordinal {ordinal} is not an association column. It is a "
                            f"dgroomes.in_memory.InMemoryColumn${type(column).__name__}"), None   # column.getClass().getName() (:103)
                child = node.create_child_node(column)
                self.nodes_in_creation_order.append(child)
                to_visit.append((child_query_node, child))
        return None, root

    def execute(self, query):  # E/DataSystemSerialIndices.java:53-102
        if query.table_name not in self.tables:
            return QueryResult.Failure(f"The query targets the table '{query.table_name}' but that table is not registered")
        table = self.tables[query.table_name]
        message, root = self._verify(query, table)
        if message:
            return QueryResult.Failure(message)
        leaves = deque()   # push / pop at the head (:75, :86, :93-97)
        nodes = deque([root])
        while nodes:
            node = nodes.popleft()
            node.filter_self()
            if not node.child_nodes:
                leaves.appendleft(node)
            else:
                nodes.extend(node.child_nodes)   # addAll appends at the tail (:88)
        while leaves:
            leaf = leaves.popleft()
            leaf.filter_parent()
            if leaf.parent is not None:
                leaves.appendleft(leaf.parent)
        self.last_indices = sorted(root.matching_bits)
        matching_rows = BitSet.from_indices(np.array(self.last_indices, dtype=np.int64), table.size())
        return QueryResult.Success(table.subset(matching_rows))   # (:100-101)

    def node_cardinalities(self):
        return [len(n.matching_bits) for n in self.nodes_in_creation_order]


def declarative_matches(tables, query):
    """The fixed point the leaf-to-root walks reach, read off the query tree directly: the rows of
    the root table for which every criterion holds and every child subtree has at least one matching associated row."""

    def match_set(table, query_node):
        columns = table.columns()
        predicates = []
        for criterion in query_node.get_criteria():
            message, predicate = _where(columns[criterion.ordinal], criterion)
            assert message is None, message
            predicates.append(predicate)
        rows = {r for r in range(table.size()) if all(p(r) for p in predicates)}
        for ordinal, child_query_node in query_node.get_children_by_ordinal().items():
            column = columns[ordinal]   # the association FROM this table's rows TO the child's
            child_rows = match_set(column.associated_entity, child_query_node)
            rows = {r for r in rows if any(t in child_rows for t in column.associations_for_index(r).targets())}
        return rows

    return sorted(match_set(tables[query.table_name], query.root_node))
