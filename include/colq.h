/*
 * colq.h -- C ABI of libcolq.so, the B200-native (sm_100a) execution module for the
 * dgroomes/java-columnar-query-engine `data-system` API.
 *
 * The reference has no FFI of its own: its engine is the Java class DataSystemSerialIndices and the
 * drop-in boundary is the `data-system` interface set.  A sibling Gradle module (`data-system-b200`,
 * source under java-columnar-query-engine_b200/java/) implements those interfaces and binds exactly
 * these symbols with java.lang.foreign downcalls.  Every entry point below names
 * the reference code it stands in for.  Citations are relative to the reference checkout:
 *   E  = data-system-serial-indices-arrays/src/main/java/dgroomes/data_system_serial_indices_arrays
 *   M  = data-model-in-memory/src/main/java/dgroomes/in_memory
 *   DS = data-system/src/main/java/dgroomes/data_system
 *
 * Conventions
 *   - plain C, no torch/CUDA types in any signature; pointers are HOST pointers unless a name ends in
 *     `_device`; sizes are int64_t; every function returns a colq_status (0 = OK).
 *   - colq_last_error(ctx) returns the message of the last non-OK status on that context (the text a
 *     QueryResult.Failure carries, DS/QueryResult.java:7).
 *   - host buffers are BORROWED for the duration of the call; the library copies them to HBM and owns
 *     the device memory until colq_destroy.  Exception: the `*_host` variants keep borrowing pinned buffers
 *     (see "Host-resident columns").
 *   - bitmasks use the java.util.BitSet word layout: row i <-> words[i >> 6] & (1L << (i & 63)),
 *     little-endian uint64, so BitSet.valueOf(LongBuffer) wraps the output directly.
 *   - one context per process and GPU; calls on one context must not overlap (the Java shim holds a lock).
 *   - there is NO CPU fallback: if no sm_100-class device is usable, colq_create fails.
 */
#ifndef COLQ_H
#define COLQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COLQ_ABI_VERSION 2

typedef enum colq_status {
    COLQ_OK = 0,
    /* -> QueryResult.Failure(colq_last_error())  (E/DataSystemSerialIndices.java:54-57, E/Verifier.java:62-104) */
    COLQ_FAILURE = 1,
    /* -> java.lang.IndexOutOfBoundsException: the reference's unchecked `columns().get(ordinal)` (E/Verifier.java:67,100) */
    COLQ_THROW_INDEX_OOB = 2,
    /* -> java.lang.NullPointerException (E/Verifier.java:41-42; M/InMemoryTable.java:70-71 for an association target outside the associated table) */
    COLQ_THROW_NULL = 3,
    /* -> java.lang.IllegalStateException (M/InMemoryColumn.java:122-126 and misuse of this ABI) */
    COLQ_THROW_ILLEGAL_STATE = 4,
    /* -> java.lang.IllegalArgumentException (DS/Query.java:33-35 duplicate child ordinal; bad enum values) */
    COLQ_THROW_ILLEGAL_ARG = 5,
    /* CUDA / NCCL / allocation error; message holds the driver text */
    COLQ_ERR_DEVICE = 6,
    /* caller-provided output capacity too small; *out_count still holds the true count */
    COLQ_ERR_CAPACITY = 7
} colq_status;

/* where a table's rows live when the context is part of a multi-GPU communicator */
typedef enum colq_placement {
    COLQ_REPLICATED = 0,   /* every rank holds all rows (the 51-row states table) */
    COLQ_SHARDED = 1       /* this rank holds a contiguous row range [global_row_base, +n_rows) (zips, cities) */
} colq_placement;

/* structured stand-ins for the reference's opaque Predicate<String> lambdas (DS/Criteria.java:17) */
typedef enum colq_str_op {
    COLQ_STR_EQ = 0,           /* "X"::equals                (app/.../Runner.java:236) */
    COLQ_STR_CONTAINS = 1,     /* s -> s.contains("X")       (Runner.java:255,257,259) */
    COLQ_STR_CMP_GT = 2,       /* s -> s.compareTo("X") > 0  (QueryTest.java:124); UTF-16 code-unit order */
    COLQ_STR_CMP_LT = 3,       /* s -> s.compareTo("X") < 0  (QueryTest.java:125) */
    COLQ_STR_CMP_GE = 4,
    COLQ_STR_CMP_LE = 5,
    COLQ_STR_NE = 6,
    COLQ_STR_STARTS_WITH = 7,
    COLQ_STR_ENDS_WITH = 8
} colq_str_op;

/* execution strategy knobs (colq_query_set_option) */
typedef enum colq_option {
    /* 1 (default): a criteria-free node reached through a to-one foreign key is evaluated lazily, only at
       the rows its parent still needs (fused FK chain gather); 0: always materialise every node's bitmask. */
    COLQ_OPT_LAZY_FK = 0,
    /* 1: record one CUDA event pair per kernel so colq_profile() can report per-stage times (adds launch gaps);
       2: one event pair per execution, around the launch with the most algorithmic bytes only, kept for every
          execution since the last colq_profile_hot() (up to 4096) -- cheap enough to leave on inside a timed region */
    COLQ_OPT_PROFILE = 1,
    /* value 2 is unused (ABI 1 reserved it for CUDA-graph replay, which never existed: a step is now two or three launches,
       so there is nothing left for a graph to save; colq_query_set_option rejects it) */
    /* 1 (default): multi-GPU exchanges (state-mask OR, final index gather) run as own kernels that store into the
       peers' HBM over NVLink (CUDA-IPC mailboxes); 0: NCCL all-gathers */
    COLQ_OPT_PEER_EXCHANGE = 3,
    /* 1 (default): one cooperative two-phase compaction launch (popcount, grid barrier, ordered write);
       2: single-pass compaction with decoupled look-back (one ordinary launch, the mask is read once; measured slower
       on B200); 0: popcount / scan / write as three launches */
    COLQ_OPT_FUSED_COMPACT = 4,
    /* 1 (default): the root node's lazy FK chains are walked by the fused compaction kernel for the rows that
       survived the root's predicates (the row scan stays a pure coalesced stream); 0: inside the row scan.
       Needs COLQ_OPT_LAZY_FK and COLQ_OPT_FUSED_COMPACT >= 1. */
    COLQ_OPT_DEFER_CHAINS = 5,
    /* what happens when a launch is about to read a host-resident column (colq_*_host) IN FULL:
       2 (default): the copy engine brings that column to HBM first (cudaMemcpyAsync on the query's stream), the
          launch and all later queries read the HBM copy;
       1: the scan kernel reads the pinned host memory in place and writes the HBM copy as a side effect;
       0: always read in place, never keep a copy (tables larger than HBM).
       Sparsely walked columns (lazy FK chains) are read in place in every mode. */
    COLQ_OPT_PROMOTE = 6,
    /* 1 (default): the multi-GPU final gather rides on the kernel that writes the indices (root_fused / compact_fused): every
       index is also stored into all ranks' mailbox slots over NVLink, the last block publishes the flags; the slots are
       concatenated when the host fetches the indices.  0: two dedicated peer_gather launches per execution. */
    COLQ_OPT_FUSED_GATHER = 7,
    /* 1 (default): the multi-GPU mask PUBLISH is done by the last CTA of the scan kernel that produced the mask instead
       of a separate one-block launch */
    COLQ_OPT_TAIL_PUBLISH = 8,
    /* 1 (default): a root node that ends in an int-predicate scan runs as ONE persistent launch -- predicate scan, deferred
       to-one chains, a tiny to-many hop feeding them (e.g. the 51-row state adjacency, with the multi-GPU mask COLLECT),
       ordered compaction and the final gather (root_fused_kernel); 0: scan_rows / csr_pull / compact_fused launches;
       2: two launches -- the ordinary non-persistent scan_rows, which additionally lists every 512-row chunk's
       survivors, then root_finish_kernel (chains, hop, ordered write, gather) over those lists.
       Needs COLQ_OPT_FUSED_COMPACT == 1. */
    COLQ_OPT_ROOT_FUSED = 9,
    /* 1 (default): a fused final gather whose plan already synchronises the ranks once per execution (a mask or bitmap
       exchange) only PUBLISHES its flags; they are awaited when the host fetches the result.  0: every execution ends with
       a wait for all ranks' flags. */
    COLQ_OPT_LAZY_GATHER_WAIT = 10,
    /* 1 (default): back-to-back executions of the same query are pipelined (execute is called over and over on one plan,
       E/DataSystemSerialIndices.java:53): the root kernel leaves the small push target behind its folded hop zeroed, so
       the next execution needs no memset, and that execution's first scan is launched as a programmatic dependent of the
       root kernel -- it streams its (immutable) column while the root kernel drains and touches shared state only after
       griddepcontrol.wait.  An event between two executions would undo the overlap, so executions of such a plan are not
       timed one by one: colq_timing.gpu_ms is -1 for them.  0: every execution starts after the previous one has finished. */
    COLQ_OPT_PIPELINE = 11
} colq_option;

typedef struct colq_ctx colq_ctx;       /* one DataSystem instance  (E/DataSystemSerialIndices.java:14-22) */
typedef struct colq_query colq_query;   /* one Query                (DS/Query.java:17-25) */
typedef int32_t colq_table;             /* table handle, valid for the owning context */

typedef struct colq_timing {
    double gpu_ms;          /* CUDA-event time of the whole kernel(+collective) pipeline of the last execute
                               (-1: not measured, the plan pipelines back-to-back executions -- COLQ_OPT_PIPELINE) */
    int32_t kernel_launches;/* kernels of this library launched by the last execute */
    int32_t collectives;    /* NCCL calls issued by the last execute */
    int64_t h2d_bytes;      /* bytes the last execute streamed host->device: host-resident columns it scanned in full */
    int64_t d2h_bytes;      /* bytes copied device->host by the last execute (count, bitmask, indices) */
} colq_timing;

typedef struct colq_stage {
    char name[48];          /* kernel name */
    double ms;              /* CUDA-event duration (only with COLQ_OPT_PROFILE) */
    int64_t rows;           /* rows the launch covered */
    int64_t bytes;          /* ALGORITHMIC bytes of the launch: every input element read once + outputs written */
} colq_stage;

/* ---- lifecycle ------------------------------------------------------------------------------------ */

int colq_abi_version(void);
/* identifies the build: "<first 16 hex digits of the SHA-256 over the csrc/ files and this header> <nvcc arch flags>" -- lets a host
   (and the driver's smoke run) check that the library it loaded was built from the sources next to it */
const char *colq_build_id(void);
/* new DataSystemSerialIndices() (E/DataSystemSerialIndices.java:20-22). device = CUDA ordinal. */
colq_status colq_create(int device, colq_ctx **out_ctx);
/* also destroys every query still alive on the context */
colq_status colq_destroy(colq_ctx *ctx);
const char *colq_last_error(const colq_ctx *ctx);
/* run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL restores the context's own stream */
colq_status colq_set_stream(colq_ctx *ctx, void *cuda_stream);
colq_status colq_get_stream(colq_ctx *ctx, void **out_cuda_stream);
colq_status colq_synchronize(colq_ctx *ctx);
/* Freed device buffers (dropped tables, query scratch) are parked in a per-device cache and re-used instead of going
   through cudaFree / cudaMalloc (milliseconds each, device-synchronising). colq_trim returns them to the driver;
   the last colq_destroy on a device does it implicitly. Cache limit: COLQ_CACHE_GB (default 24). */
colq_status colq_trim(colq_ctx *ctx);

/* ---- multi-GPU: one process per GPU, host bootstraps the communicator ---------------- */

/* 128-byte ncclUniqueId produced on rank 0; the host ships it to the other ranks by any side channel */
colq_status colq_comm_unique_id(colq_ctx *ctx, uint8_t out_id[128]);
colq_status colq_comm_init(colq_ctx *ctx, const uint8_t id[128], int n_ranks, int rank);
colq_status colq_comm_info(const colq_ctx *ctx, int *out_n_ranks, int *out_rank);
/*
 * The same communicator for ONE host process that drives all GPUs -- the shape of the reference, where the engine is a
 * single object in a single JVM (E/DataSystemSerialIndices.java:14-22; app/.../Runner.java:40): ctxs[i] (one context per
 * GPU, created with colq_create on distinct devices) becomes rank i of n_ranks.  The mailboxes are made mutually
 * reachable with cudaDeviceEnablePeerAccess instead of CUDA IPC, NCCL is not involved at all, and the exchange kernels
 * are the same.  Because those kernels wait for one another ACROSS GPUs, the host must enqueue the work of every rank
 * before it fetches any rank's result: call colq_execute_group (or colq_execute_async on every context), then
 * colq_fetch per context.  A plain colq_execute on one context of a local group would wait for peers that were never
 * launched (the kernels give up after 4 s: COLQ_ERR_DEVICE).
 */
colq_status colq_comm_init_local(colq_ctx **ctxs, int n_ranks);
/* colq_execute_async(ctxs[i], queries[i]) for i in [0, n); stops at the first non-OK status and returns it */
colq_status colq_execute_group(colq_ctx **ctxs, colq_query **queries, int n);
/* Waits for every rank and writes the match counts (global count for a sharded root) to out_counts[n] (nullable).  If a
   rank's result block was too small, the query is first re-run on ALL ranks with a larger block -- the step a
   one-process-per-GPU host performs inside colq_fetch on every rank at once.  Afterwards colq_fetch(ctxs[i], ...) copies
   rank i's bitmask / indices out without waiting for anybody. */
colq_status colq_fetch_group(colq_ctx **ctxs, colq_query **queries, int n, int64_t *out_counts);

/* ---- tables and columns: InMemoryTable.ofColumns / associateTo (M/InMemoryTable.java:32-35,44-90) --- */

/* n_rows = this rank's rows. For COLQ_SHARDED tables global_row_base is added to emitted row indices. */
colq_status colq_table_create(colq_ctx *ctx, int64_t n_rows, colq_placement placement, int64_t global_row_base,
                              colq_table *out_table);
/* DataSystemSerialIndices.register (E/DataSystemSerialIndices.java:27-29): HashMap.put, the last put wins */
colq_status colq_register(colq_ctx *ctx, const char *table_name, colq_table table);

/* IntegerColumn(int[] ints) (M/InMemoryColumn.java:46) at `ordinal` (must be the next free ordinal or an unset one) */
colq_status colq_col_i32(colq_ctx *ctx, colq_table table, int ordinal, const int32_t *values, int64_t n);
/* StringColumn(String[] strings) (M/InMemoryColumn.java:64) as n+1 shard-relative uint32 offsets + UTF-8 bytes */
colq_status colq_col_str(colq_ctx *ctx, colq_table table, int ordinal, const uint32_t *offsets, const uint8_t *bytes,
                         int64_t n, int64_t n_bytes);
/* BooleanColumn(boolean[] bools) (M/InMemoryColumn.java:28): registered so ordinals line up; any criterion on it
   is a Failure exactly as in the reference (E/Verifier.java:82-84) */
colq_status colq_col_bool(colq_ctx *ctx, colq_table table, int ordinal, const uint8_t *values, int64_t n);
/* zero-copy variants: adopt device buffers the host already owns (e.g. torch tensors). 16-byte aligned; the
   buffers must outlive the table; `*_capacity` is the allocation size in bytes (the TMA path reads whole
   16-byte lines inside it). */
colq_status colq_col_i32_device(colq_ctx *ctx, colq_table table, int ordinal, const void *values_device, int64_t n);
colq_status colq_col_str_device(colq_ctx *ctx, colq_table table, int ordinal, const void *offsets_device,
                                int64_t offsets_capacity, const void *bytes_device, int64_t bytes_capacity,
                                int64_t n, int64_t n_bytes);

/*
 * Host-resident columns: the off-heap MemorySegment the Java shim fills IS the column (north_star: "off-heap
 * MemorySegment column buffers").  Like DataSystemSerialIndices.register (E/DataSystemSerialIndices.java:27-29,
 * which only keeps a reference) nothing is copied at registration: the buffer is BORROWED until colq_table_destroy
 * / colq_destroy and the kernels read it in place over PCIe (pinned, device-mapped memory).  A query therefore moves
 * only the bytes it touches -- the columns it scans, plus single sectors of the foreign-key columns it walks lazily --
 * and the first scan that streams a whole column also leaves a copy in HBM (COLQ_OPT_PROMOTE), so the next query on
 * it runs at HBM speed.
 *   - buffers must be pinned: allocate them with colq_host_alloc (Java: MemorySegment.ofAddress(p).reinterpret(n))
 *     or pin an Arena allocation with colq_host_register; 16-byte aligned;
 *   - `*_capacity` is the usable size of the buffer in bytes: int / association columns need n*4 rounded up to 16,
 *     string offsets (n+1)*4 rounded up to 16, string bytes n_bytes rounded up to 16 plus 32 (the TMA path reads
 *     whole 16-byte lines);
 *   - to-one targets of colq_associate_fk_host are range-checked on the rows a query walks, not at registration: a
 *     target outside the associated table makes colq_execute / colq_fetch return COLQ_THROW_NULL (the reference's
 *     NPE at associateTo, M/InMemoryTable.java:70-71).
 */
colq_status colq_host_alloc(colq_ctx *ctx, int64_t bytes, void **out_ptr);
colq_status colq_host_free(colq_ctx *ctx, void *ptr);
colq_status colq_host_register(colq_ctx *ctx, void *ptr, int64_t bytes);
colq_status colq_host_unregister(colq_ctx *ctx, void *ptr);
colq_status colq_col_i32_host(colq_ctx *ctx, colq_table table, int ordinal, const int32_t *values_pinned,
                              int64_t capacity_bytes, int64_t n);
colq_status colq_col_str_host(colq_ctx *ctx, colq_table table, int ordinal, const uint32_t *offsets_pinned,
                              int64_t offsets_capacity, const uint8_t *bytes_pinned, int64_t bytes_capacity, int64_t n,
                              int64_t n_bytes);

/*
 * Dictionary-encoded StringColumn: row i holds the string dictionary[codes[i]]; the
 * dictionary is an ordinary offsets + bytes column of n_dict DISTINCT values (host buffers, copied).  Every string
 * criterion is evaluated once per distinct value -- by the string kernel over the dictionary for the structured
 * predicates, or by the HOST for an opaque Predicate<String> lambda (colq_query_criteria_str_accept) -- and the row
 * scan (4 bytes per row instead of offsets + bytes) only tests bit `code` of the resulting n_dict-bit mask.  This is
 * how the reference's unchanged lambdas (app/.../Runner.java:236,255-259) run without a CPU row scan.
 * _device adopts an int32 code buffer already in HBM; _host borrows a pinned one (see "Host-resident columns").
 */
/* the same for an IntegerColumn: int32 codes + the n_dict distinct int32 values; a closed-interval criterion is evaluated
   over the distinct values, an opaque IntPredicate lambda by the host (colq_query_criteria_i32_accept) */
colq_status colq_col_i32_dict(colq_ctx *ctx, colq_table table, int ordinal, const int32_t *codes, int64_t n,
                              const int32_t *dict_values, int64_t n_dict);
colq_status colq_col_i32_dict_host(colq_ctx *ctx, colq_table table, int ordinal, const int32_t *codes_pinned,
                                   int64_t capacity_bytes, int64_t n, const int32_t *dict_values, int64_t n_dict);
colq_status colq_col_str_dict(colq_ctx *ctx, colq_table table, int ordinal, const int32_t *codes, int64_t n,
                              const uint32_t *dict_offsets, const uint8_t *dict_bytes, int64_t n_dict, int64_t n_dict_bytes);
colq_status colq_col_str_dict_device(colq_ctx *ctx, colq_table table, int ordinal, const void *codes_device, int64_t n,
                                     const uint32_t *dict_offsets, const uint8_t *dict_bytes, int64_t n_dict,
                                     int64_t n_dict_bytes);
colq_status colq_col_str_dict_host(colq_ctx *ctx, colq_table table, int ordinal, const int32_t *codes_pinned,
                                   int64_t capacity_bytes, int64_t n, const uint32_t *dict_offsets,
                                   const uint8_t *dict_bytes, int64_t n_dict, int64_t n_dict_bytes);

/*
 * x.associateTo(y, associations) (M/InMemoryTable.java:44-90): creates the forward AssociationColumn on x at
 * x_ordinal AND its reverse (transposed) column on y at y_ordinal, cross-linked (:83-85).  Only the forward data
 * is stored; hops through the reverse column are executed as a push through the forward data, which is the same
 * relation because the reference builds the reverse column as the transpose (:55-82).
 *   _fk : every row is Association.One(fk[i]) or Association.None (fk[i] == -1)        (DS/Association.java:27-43)
 *   _csr: row i is associated to targets[offsets[i] .. offsets[i+1]) (None / One / Many) (DS/Association.java:27-51)
 * For two COLQ_SHARDED tables the indices are shard-local. A target outside [0, y.n_rows) is
 * COLQ_THROW_NULL like the reference's NPE (:70-71).
 */
colq_status colq_associate_fk(colq_ctx *ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal,
                              const int32_t *fk, int64_t n);
colq_status colq_associate_csr(colq_ctx *ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal,
                               const int64_t *offsets, const int32_t *targets, int64_t n, int64_t nnz);
colq_status colq_associate_fk_device(colq_ctx *ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal,
                                     const void *fk_device, int64_t n);
/* host-resident variant of colq_associate_fk (see "Host-resident columns" above) */
colq_status colq_associate_fk_host(colq_ctx *ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal,
                                   const int32_t *fk_pinned, int64_t capacity_bytes, int64_t n);

/*
 * Ingest on the device.  The reference classifies and transposes Association[] objects and copies
 * String[] columns on the host at load time (M/InMemoryTable.java:44-90, app/.../Runner.java:89-196); here the flat arrays
 * are shipped as they are and kernels do the rest:
 *   colq_associate       x.associateTo(y, associations) from a CSR (row i -> targets[offsets[i] .. offsets[i+1])): offsets
 *                        order, target range (COLQ_THROW_NULL like :70-71) and the largest degree are computed in one pass
 *                        on the GPU; when every row is Association.None or Association.One the column is stored as the
 *                        dense to-one form (-1 = None), otherwise as the CSR.  *out_is_fk (nullable) says which.
 *   colq_col_str_encode  turns an already registered plain string column (colq_col_str / _device / _host) into a
 *                        dictionary-encoded one IN PLACE: hash insert with the first row of each distinct value as its
 *                        representative, byte-exact verification, codes in first-appearance order.  Afterwards the column
 *                        behaves exactly like one registered with colq_col_str_dict.  COLQ_ERR_CAPACITY when the column has
 *                        more than ~45 M distinct values (the column is left unchanged).
 *   colq_col_dict_str    reads the distinct values of a dictionary-encoded column back (n_dict + 1 offsets + UTF-8 bytes):
 *                        the host evaluates an opaque Predicate<String> lambda on these (colq_query_criteria_str_accept).
 */
colq_status colq_associate(colq_ctx *ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal, const int64_t *offsets,
                           const int32_t *targets, int64_t n, int64_t nnz, int *out_is_fk);
colq_status colq_associate_device(colq_ctx *ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal,
                                  const void *offsets_device, const void *targets_device, int64_t n, int64_t nnz, int *out_is_fk);
colq_status colq_col_str_encode(colq_ctx *ctx, colq_table table, int ordinal, int64_t *out_n_dict);
colq_status colq_col_dict_str(colq_ctx *ctx, colq_table table, int ordinal, uint32_t *out_offsets, int64_t offsets_capacity,
                              uint8_t *out_bytes, int64_t bytes_capacity, int64_t *out_n_dict, int64_t *out_n_bytes);

/*
 * Cross-shard associations: the general case of filterParent (E/ExecutionContext.java:100-122), where
 * an association between two tables that are BOTH sharded -- or from a replicated table into a sharded one -- points at
 * rows of other ranks.  The targets are then GLOBAL row indices of the associated table:
 *   colq_table_partition   declares how a COLQ_SHARDED table is split: bounds[r] .. bounds[r+1] are rank r's global rows
 *                          (bounds[0] = 0, every bound below the total a multiple of 64 so that a shard is whole BitSet
 *                          words; this rank's entry must match its colq_table_create).  Same array on every rank.
 *   colq_associate_*_global  like colq_associate_fk / _csr with targets in [0, global rows of y) (-1 = None for _fk).
 * A hop through such a column exchanges a bitmap over y's (or x's) GLOBAL rows through the peer-mapped heap behind the
 * mailboxes: all-gather of the child's bits for a pull, OR-reduce-scatter of the reach bits for a push (own kernels over
 * NVLink; needs the peer-memory exchange -- there is no NCCL path for these).  On a single rank they are ordinary
 * associations.  Heap size per execution parity: COLQ_PEER_HEAP_MB (default 128 MB = 1 G rows of bitmap).
 */
colq_status colq_table_partition(colq_ctx *ctx, colq_table table, const int64_t *bounds, int n_ranks);
colq_status colq_associate_fk_global(colq_ctx *ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal,
                                     const int32_t *fk_global, int64_t n);
colq_status colq_associate_csr_global(colq_ctx *ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal,
                                      const int64_t *offsets, const int32_t *targets_global, int64_t n, int64_t nnz);

/* Release a table's device memory and its registrations (the reference leaves this to the garbage collector).
   Association columns of other tables that pointed at it become unset. The handle stays reserved. */
colq_status colq_table_destroy(colq_ctx *ctx, colq_table table);
colq_status colq_table_size(const colq_ctx *ctx, colq_table table, int64_t *out_rows);   /* Table.size()  DS/Table.java:29 */
colq_status colq_table_width(const colq_ctx *ctx, colq_table table, int *out_columns);   /* Table.width() DS/Table.java:22 */

/* ---- queries: Query / Query.Node / Criteria (DS/Query.java:17-54, DS/Criteria.java:10-20) ----------- */

/* new Query(tableName); node 0 is rootNode. The name is resolved at execute time, like the reference (:54-57). */
colq_status colq_query_create(colq_ctx *ctx, const char *table_name, colq_query **out_query);
colq_status colq_query_destroy(colq_query *query);
/* Query.Node.createChild(ordinal) (DS/Query.java:31-38): COLQ_THROW_ILLEGAL_ARG on a duplicate ordinal */
colq_status colq_query_child(colq_query *query, int parent_node, int ordinal, int *out_node);
/* addCriteria(new Criteria.IntCriteria(ordinal, v -> lo <= v && v <= hi)) (DS/Criteria.java:19) */
colq_status colq_query_criteria_i32_range(colq_query *query, int node, int ordinal, int32_t lo, int32_t hi);
/* addCriteria(new Criteria.StringCriteria(ordinal, <op needle>)) (DS/Criteria.java:17) */
colq_status colq_query_criteria_str(colq_query *query, int node, int ordinal, colq_str_op op, const uint8_t *needle,
                                    int32_t needle_len);
/* addCriteria(new Criteria.StringCriteria(ordinal, <any Predicate<String>>)) over a dictionary-encoded column: the host
   evaluated the predicate on each of the n_dict dictionary entries; bit d of accept_words (BitSet layout) = result for
   entry d.  COLQ_FAILURE at execute when the column is not dictionary-encoded. */
colq_status colq_query_criteria_str_accept(colq_query *query, int node, int ordinal, const uint64_t *accept_words,
                                           int64_t n_dict);
/* addCriteria(new Criteria.IntCriteria(ordinal, <any IntPredicate>)) over a dictionary-encoded int column (DS/Criteria.java:19) */
colq_status colq_query_criteria_i32_accept(colq_query *query, int node, int ordinal, const uint64_t *accept_words,
                                           int64_t n_dict);
/* A criterion over a BooleanColumn (M/InMemoryColumn.java:28-44).  The reference declares the column kind and its
   BooleanColumnFilterable.where(Predicate<Boolean>) (DS/ColumnFilterable.java:20-22) but its Verifier answers Failure
   for it (E/Verifier.java:82-84, "not supported yet") and Criteria has no boolean member (DS/Criteria.java:10-20): this
   is the SURVEY.md 8(f4) extension.  A Predicate<Boolean> has a two-entry truth table, so the host evaluates the lambda
   on FALSE and on TRUE and passes both answers; a row matches when accept_{its value} != 0 (a byte != 0 is TRUE).
   Int / string criteria on a boolean column keep the reference's Failure. */
colq_status colq_query_criteria_bool(colq_query *query, int node, int ordinal, int accept_false, int accept_true);
colq_status colq_query_set_option(colq_query *query, colq_option option, int value);

/*
 * DataSystemSerialIndices.execute (E/DataSystemSerialIndices.java:53-102): verify/link (E/Verifier.java:40-111),
 * per-node predicate scan (E/ExecutionContext.java:79-94), leaf-to-root association pruning (:100-122) and the
 * ascending index half of Table.subset (M/InMemoryTable.java:121-131), all on the device.
 *   out_bitmask : nullable; receives ceil(rows/64) words of the ROOT node's matching bits (this rank's rows)
 *   out_indices : nullable; receives the matching row indices, ascending (global indices for a sharded root;
 *                 with a communicator, every rank receives all ranks' indices concatenated in rank order)
 *   out_count   : number of matching rows (global count when the root table is sharded over a communicator)
 * Returns COLQ_OK or one of the statuses documented on colq_status.
 */
colq_status colq_execute(colq_ctx *ctx, colq_query *query, uint64_t *out_bitmask, int64_t bitmask_capacity_words,
                         int32_t *out_indices, int64_t indices_capacity, int64_t *out_count, colq_timing *out_timing);
/*
 * Same pipeline without any host synchronisation or result copy: kernels and collectives (including the final
 * gather of matched indices) are only enqueued on the context's stream; results stay in HBM until colq_fetch.
 * Used to time K back-to-back executions with CUDA events.
 */
colq_status colq_execute_async(colq_ctx *ctx, colq_query *query);
colq_status colq_fetch(colq_ctx *ctx, colq_query *query, uint64_t *out_bitmask, int64_t bitmask_capacity_words,
                       int32_t *out_indices, int64_t indices_capacity, int64_t *out_count, colq_timing *out_timing);
/*
 * Result materialisation -- the value half of Table.subset(BitSet) (M/InMemoryTable.java:106-159).  The reference
 * re-scans every column for the set bits; here the ascending index list of the last colq_execute / colq_fetch is
 * already in HBM, so one column of the result is a device-side gather (O(matches)) copied to the host in compact form.
 * Rows are this rank's matching rows of the query's ROOT table (colq_result_count), ascending.
 *   _i32  : IntegerColumn values, or the targets of a stored to-one AssociationColumn (-1 = None; indices are NOT
 *           re-mapped, exactly like the reference :143-154)
 *   _bool : BooleanColumn values (0 / 1)
 *   _str  : StringColumn (plain or dictionary-encoded) as count+1 uint32 offsets + UTF-8 bytes
 *   _csr  : a stored to-many AssociationColumn as count+1 int64 offsets + int32 targets (un-remapped)
 * The reverse side of an association has no stored data (it is the transpose of its peer): COLQ_FAILURE; the host
 * subsets its own copy.  Too small a buffer: COLQ_ERR_CAPACITY with the out-counts set, so the caller can size and
 * call again (passing NULL buffers is the size query).
 */
colq_status colq_result_count(colq_ctx *ctx, colq_query *query, int64_t *out_rows);
colq_status colq_result_i32(colq_ctx *ctx, colq_query *query, int ordinal, int32_t *out_values, int64_t capacity,
                            int64_t *out_count);
colq_status colq_result_bool(colq_ctx *ctx, colq_query *query, int ordinal, uint8_t *out_values, int64_t capacity,
                             int64_t *out_count);
colq_status colq_result_str(colq_ctx *ctx, colq_query *query, int ordinal, uint32_t *out_offsets, int64_t offsets_capacity,
                            uint8_t *out_bytes, int64_t bytes_capacity, int64_t *out_count, int64_t *out_n_bytes);
colq_status colq_result_csr(colq_ctx *ctx, colq_query *query, int ordinal, int64_t *out_offsets, int64_t offsets_capacity,
                            int32_t *out_targets, int64_t targets_capacity, int64_t *out_count, int64_t *out_nnz);

/* per-kernel stages of the last execute of `query` (times only with COLQ_OPT_PROFILE); returns the stage count */
colq_status colq_profile(const colq_query *query, colq_stage *out_stages, int capacity, int *out_n_stages);
/* COLQ_OPT_PROFILE == 2: synchronises, then reports the dominant launch of `query` -- name, rows, algorithmic bytes and
   the MEAN CUDA-event duration over the *out_samples executions recorded since the previous call -- and resets. */
colq_status colq_profile_hot(colq_query *query, colq_stage *out_stage, int *out_samples);
/* cardinality of every execution node's bitmask after the last execute, in the reference's node creation (BFS)
   order; -1 for nodes that were fused away and never materialised. Debug / parity aid. */
colq_status colq_node_cardinalities(colq_ctx *ctx, const colq_query *query, int64_t *out, int capacity, int *out_n);

#ifdef __cplusplus
}
#endif
#endif /* COLQ_H */
