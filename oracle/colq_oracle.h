/*
 * colq_oracle.h -- CPU restatement of the reference's serial-indices query engine.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or the timed CPU baseline.  libcolq.so never links or calls it.
 *
 * Parity status: PINNED to every known-answer vector the reference's own tests hold for this
 * path (the 5 QueryTest cases and the TheTest loader cardinalities; see tests/test_oracle_goldens.py).
 * The two headline queries (Plymouth, North/South/North) are only logged by the reference
 * (app/.../Runner.java:246,269), never asserted, and no JVM exists in this image to run it, so
 * their expected outputs are ORACLE-DERIVED goldens (labelled as such in tests/golden/).
 *
 * All citations are relative to /root/reference/.
 *   E  = data-system-serial-indices-arrays/src/main/java/dgroomes/data_system_serial_indices_arrays
 *   M  = data-model-in-memory/src/main/java/dgroomes/in_memory
 *   DS = data-system/src/main/java/dgroomes/data_system
 */
#ifndef COLQ_ORACLE_H
#define COLQ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* execute() outcomes */
enum {
    ORC_SUCCESS = 0,            /* QueryResult.Success                      (DS/QueryResult.java:4) */
    ORC_FAILURE = 1,            /* QueryResult.Failure(message)             (DS/QueryResult.java:7) */
    ORC_THROW_INDEX_OOB = 2,    /* java.lang.IndexOutOfBoundsException      (E/Verifier.java:67,100) */
    ORC_THROW_NULL = 3,         /* java.lang.NullPointerException           (E/Verifier.java:41-42; M/InMemoryTable.java:70) */
    ORC_THROW_ILLEGAL_STATE = 4,/* java.lang.IllegalStateException          (M/InMemoryColumn.java:122-126) */
    ORC_THROW_ILLEGAL_ARG = 5   /* java.lang.IllegalArgumentException       (DS/Query.java:33-35) */
};

/* string predicate operators: structured stand-ins for the reference's Predicate<String> lambdas */
enum {
    ORC_STR_EQ = 0,          /* "X"::equals                 (app/.../Runner.java:236; QueryTest.java:169,194,316-320) */
    ORC_STR_CONTAINS = 1,    /* s -> s.contains("X")        (Runner.java:255,257,259) */
    ORC_STR_CMP_GT = 2,      /* s -> s.compareTo("X") > 0   (QueryTest.java:124) */
    ORC_STR_CMP_LT = 3,      /* s -> s.compareTo("X") < 0   (QueryTest.java:125) */
    ORC_STR_CMP_GE = 4,
    ORC_STR_CMP_LE = 5,
    ORC_STR_NE = 6,
    ORC_STR_STARTS_WITH = 7,
    ORC_STR_ENDS_WITH = 8
};

typedef struct orc_system orc_system;   /* E/DataSystemSerialIndices.java:14 */
typedef struct orc_query orc_query;     /* DS/Query.java:17 */

orc_system *orc_system_new(void);
void orc_system_free(orc_system *s);
const char *orc_last_message(const orc_system *s);

/* InMemoryTable.ofColumns() with no columns yet (M/InMemoryTable.java:32-35); returns table id */
int orc_table_new(orc_system *s);
/* append columns (returns the new ordinal, <0 on error) */
int orc_table_add_ints(orc_system *s, int table, const int32_t *v, int64_t n);            /* M/InMemoryColumn.java:46 */
int orc_table_add_strings(orc_system *s, int table, const uint32_t *offsets,               /* M/InMemoryColumn.java:64 */
                          const uint8_t *bytes, int64_t n);
int orc_table_add_bools(orc_system *s, int table, const uint8_t *v, int64_t n);           /* M/InMemoryColumn.java:28 */
/*
 * x.associateTo(y, associations) (M/InMemoryTable.java:44-90).  Association[] is given flattened:
 * kind[i] in {0 None, 1 One, 2 Many}, targets of row i = targets[offsets[i] .. offsets[i+1]).
 * Appends the forward column to x and the transposed (reverse) column to y; returns ORC_SUCCESS or
 * ORC_THROW_NULL (a target outside [0, y.size) hits `yIndexToXAssociations.get(yIndex)` == null, :70).
 */
int orc_table_associate(orc_system *s, int x, int y, const uint8_t *kind, const int64_t *offsets,
                        const int32_t *targets, int64_t n, int *x_ordinal, int *y_ordinal);
int64_t orc_table_size(const orc_system *s, int table);                                    /* M/InMemoryTable.java:92-101 */
int orc_table_width(const orc_system *s, int table);

void orc_register(orc_system *s, const char *name, int table);                             /* E/DataSystemSerialIndices.java:27 */

/* Query construction (DS/Query.java:17-54). Node 0 is the root. */
orc_query *orc_query_new(const char *table_name);
void orc_query_free(orc_query *q);
int orc_query_create_child(orc_query *q, int parent_node, int ordinal);  /* new node id, or -ORC_THROW_ILLEGAL_ARG */
void orc_query_add_int_range(orc_query *q, int node, int ordinal, int32_t lo, int32_t hi); /* closed [lo,hi] */
void orc_query_add_str(orc_query *q, int node, int ordinal, int op, const uint8_t *needle, int32_t len);

/*
 * DataSystemSerialIndices.execute (E/DataSystemSerialIndices.java:53-102).
 * On ORC_SUCCESS: *out_words receives a malloc'd array of ceil(size/64) little-endian BitSet words
 * (the root node's matchingBits, E/ExecutionContext.java:27-29), *out_indices a malloc'd ascending
 * int32 row list (the order InMemoryTable.subset copies rows in, M/InMemoryTable.java:106-159) and
 * *out_count its length.  Free both with orc_free().  nthreads==1 is the literal serial engine;
 * nthreads>1 splits the row loops over OpenMP threads (same results; bench baseline only).
 */
int orc_execute(orc_system *s, const orc_query *q, int nthreads, uint64_t **out_words, int64_t *out_nwords,
                int32_t **out_indices, int64_t *out_count);
/* cardinality of every execution node's bitset after the last successful execute, in creation (BFS) order */
int orc_last_node_cardinalities(const orc_system *s, int64_t *out, int cap);
void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
