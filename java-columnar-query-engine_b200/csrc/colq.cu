// colq.cu -- libcolq.so: context, device-resident tables, query verifier, planner/executor and the C ABI.
//
// The structure follows the reference's execute() (E/DataSystemSerialIndices.java:53-102):
//   verify/link (E/Verifier.java:40-111)  ->  per-node self filter (E/ExecutionContext.java:79-94)
//   ->  leaf-to-root association pruning (:100-122)  ->  ascending subset indices (M/InMemoryTable.java:121-131)
// but evaluates the node tree bottom-up in ONE post-order pass (the reference's repeated leaf-to-root walks reach
// the same fixed point because every step is a monotone AND), and fuses work across nodes:
//   * a hop through a forward to-one column is PULLED (parent row tests its child's bit) inside the parent's scan,
//   * chains of criteria-free to-one hops are walked lazily only for rows that survived the parent's predicates,
//   * a hop through a reverse column is PUSHED by the epilogue of the child's own scan kernel,
//   * the ROOT node -- predicate scan, its chains, a tiny to-many hop feeding them, the ordered compaction and the
//     multi-GPU gather -- runs as one persistent launch (root_fused_kernel), started as a programmatic dependent of the
//     kernel in front of it,
//   * in a multi-GPU communicator (one process per GPU over CUDA IPC, or ONE process driving all GPUs over peer access) a
//     push from a sharded child into a replicated parent is followed by an OR of the (tiny) parent mask over NVLink peer
//     memory: published by the last CTA of the producing scan, collected inside the root's launch, as fence-free
//     flag-in-data words; hops between tables that are both sharded exchange whole bitmaps (all-gather / OR-reduce-scatter)
//     through a peer-mapped heap; the matched indices are written straight into every rank's mailbox by the kernel that
//     produces them.
// Columns may live in HBM or stay in pinned host memory (colq_*_host; moved on first touch only), string / int columns
// may be dictionary-encoded (predicates evaluated per distinct value, rows tested by code lookup), and the load-time work
// (dictionary building, association classification and validation) runs on the device too (colq_ingest.cuh).
// There is no CPU fallback anywhere in this file.
//
// E = data-system-serial-indices-arrays/src/main/java/dgroomes/data_system_serial_indices_arrays
// M = data-model-in-memory/src/main/java/dgroomes/in_memory, DS = data-system/src/main/java/dgroomes/data_system
#include "../../include/colq.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "colq_kernels.cuh"
#include "colq_ingest.cuh"
#include "colq_nccl.h"

using namespace colq;

// =====================================================================================================
// internal types
// =====================================================================================================

namespace {

// Freed device buffers are parked here and handed out again (best fit, at most 25 % larger than asked): cudaMalloc /
// cudaFree of multi-GB column buffers cost milliseconds to tens of milliseconds and synchronise the device, which
// would dominate a cold query over host-resident tables (DESIGN.md "e2e").  All work of a context is enqueued on one
// stream, so a block that is re-issued is only touched by work ordered after its previous user.
struct DeviceCache {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;
    size_t cached_bytes = 0;
    size_t limit_bytes = (size_t)((getenv("COLQ_CACHE_GB") ? atof(getenv("COLQ_CACHE_GB")) : 24.0) * (double)((size_t)1 << 30));
    int live_contexts = 0;

    void* take(size_t bytes, size_t* block_bytes) {
        std::lock_guard<std::mutex> g(mu);
        auto it = free_blocks.lower_bound(bytes);
        if (it == free_blocks.end() || it->first > bytes + bytes / 4 + 4096) return nullptr;
        void* p = it->second;
        *block_bytes = it->first;
        cached_bytes -= it->first;
        free_blocks.erase(it);
        return p;
    }
    bool park(void* p, size_t bytes) {
        std::lock_guard<std::mutex> g(mu);
        if (cached_bytes + bytes > limit_bytes) return false;
        free_blocks.emplace(bytes, p);
        cached_bytes += bytes;
        return true;
    }
    void trim() {
        std::lock_guard<std::mutex> g(mu);
        for (auto& kv : free_blocks) cudaFree(kv.second);
        free_blocks.clear();
        cached_bytes = 0;
    }
};
DeviceCache& device_cache() {
    static DeviceCache caches[64];
    int dev = 0;
    cudaGetDevice(&dev);
    return caches[dev & 63];
}

struct DevBuf {
    void* ptr = nullptr;
    size_t bytes = 0;
    bool owned = false;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            ptr = o.ptr; bytes = o.bytes; owned = o.owned;
            o.ptr = nullptr; o.bytes = 0; o.owned = false;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (owned && ptr && !device_cache().park(ptr, bytes)) cudaFree(ptr);
        ptr = nullptr; bytes = 0; owned = false;
    }
};

enum ColKind { COL_UNSET = 0, COL_I32, COL_STR, COL_BOOL, COL_ASSOC };

struct Column {
    ColKind kind = COL_UNSET;
    int64_t n = 0;
    DevBuf data;      // i32 values | string bytes | to-one fk
    DevBuf offsets;   // string offsets (u32, n+1) | CSR offsets (i64, n+1)
    DevBuf targets;   // CSR targets
    int64_t n_bytes = 0;         // string payload bytes
    int64_t nnz = 0;             // CSR edges
    int64_t bytes_capacity = 0;  // usable allocation of `data` for strings (multiple of 16)
    int64_t max_tile_bytes = 0;  // largest payload of a 1024-row scan tile (sizes the TMA ring slot)
    // association columns (M/InMemoryColumn.java:85-138)
    bool forward = false;  // true: this column owns the data; false: reverse column = transpose of the peer
    bool is_fk = false;    // forward data is a dense to-one array (else CSR)
    int peer_table = -1;   // associatedEntity
    int peer_ordinal = -1; // reverseAssociatedColumn
    // host-resident columns (colq_*_host): `data` / `offsets` point at pinned host memory that the kernels read in
    // place over PCIe.  The first scan that streams the whole column also fills `promoted*` (HBM copies), which then
    // replace the host pointers: later queries run at HBM speed.  Sparsely walked columns (lazy FK chains) stay put.
    // dictionary-encoded string column (colq_col_str_dict): `data` holds int32 codes, the n_dict DISTINCT values live
    // in an ordinary (offsets, bytes) string column of their own.  A predicate is evaluated once per distinct value.
    std::unique_ptr<Column> dict;
    bool global_targets = false;  // association targets are GLOBAL rows of a sharded target table (cross-shard hops)
    bool host_resident = false;
    bool fk_validated = true;  // false: to-one targets are range-checked on the rows a query walks, not at ingest
    DevBuf promoted, promoted_offsets;
};

struct Table {
    int64_t n_rows = 0;
    colq_placement placement = COLQ_REPLICATED;
    int64_t row_base = 0;
    std::vector<int64_t> part;  // colq_table_partition: global row bounds of every rank's shard (n_ranks + 1 entries)
    std::vector<Column> cols;
    int64_t global_rows() const { return part.empty() ? n_rows : part.back(); }
};

struct Crit {
    int ordinal = 0;
    bool is_str = false;
    int32_t lo = 0, hi = 0;
    int op = 0;
    std::vector<uint8_t> needle;
    DevBuf needle_dev;
    // an opaque Predicate<String> over a dictionary-encoded column: the host evaluated it per dictionary entry
    bool is_accept = false;
    int64_t accept_n = 0;
    DevBuf accept_dev;
    // a Predicate<Boolean>: its truth table (colq_query_criteria_bool)
    bool is_bool = false;
    bool accept_false = false, accept_true = false;
};

struct QNode {
    std::vector<Crit> crit;
    std::vector<std::pair<int, int>> children;  // (ordinal, node)
};

// one kernel launch (or collective / memset) of a planned query
enum OpKind { K_SCAN_ROWS, K_SCAN_CODES, K_SCAN_STR, K_CSR_PULL, K_PUSH_BITS, K_AND, K_FILL, K_ZERO, K_ALLGATHER_OR, K_POPC, K_SCAN_COUNTS, K_COMPACT, K_GATHER, K_COMPACT_FUSED, K_COMPACT_LOOKBACK, K_ROOT_FUSED, K_ROOT_FINISH, K_PEER_BITS_ALLGATHER, K_PEER_BITS_REDUCE, K_PEER_MASK_PUBLISH, K_PEER_MASK_COLLECT, K_PEER_GATHER, K_SCAN_BOOL };

struct Op {
    OpKind kind;
    int node = -1;
    int np = 0, ng = 0;
    bool eager = false;
    ScanCodesParams codes{};
    ScanBoolParams boolp{};
    bool never = false;  // an empty int interval: the launch degenerates to clearing the mask
    ScanRowsParams rows{};
    ScanStrParams str{};
    CsrPullParams csr{};
    PushBitsParams pushb{};
    // generic
    u32* dst = nullptr;
    const u32* src = nullptr;
    int64_t n_words = 0, n_rows = 0, n_alloc_words = 0;
    u32* gathered = nullptr;
    u32* block_counts = nullptr;
    u64* block_offsets = nullptr;
    u64* total = nullptr;
    int32_t* out_idx = nullptr;
    int64_t capacity = 0, row_base = 0, n_blocks = 0;
    CompactFusedParams cfused{};
    CompactLookbackParams clook{};
    RootFusedParams rfused{};
    PeerBitsParams pbits{};
    PeerMaskParams pmask{};
    PeerGatherParams pgather{};
    bool tail_publish = false;  // K_SCAN_STR / K_SCAN_CODES: the last CTA publishes the pushed mask to the peers
    bool dict_scan = false;  // K_SCAN_ROWS over the distinct values of a dictionary-encoded int column
    bool list = false;       // K_SCAN_ROWS: the LIST instantiation in front of a K_ROOT_FINISH (per-chunk survivor lists)
    bool gather = false;  // K_COMPACT_FUSED: the peer-memory final gather is fused into this launch
    int publish_op = -1;  // K_PEER_MASK_COLLECT / a csr_pull with a fused collect: index of the matching publish
    // launch shape for scan_str
    int grid = 0;
    size_t smem = 0;
    // accounting
    const char* name = "";
    int64_t acct_rows = 0, acct_bytes = 0;
};

struct XNode {
    int table = -1;
    int parent = -1;
    int parent_ordinal = -1;
    std::vector<const Crit*> preds;
    std::vector<std::pair<int, int>> children;  // (ordinal on this table, xnode)
    u32* bits = nullptr;   // final bitmask if materialised
    bool all_ones = false;
    bool fused = false;    // folded into a lazy FK chain, never materialised
};

struct Pool {
    std::vector<DevBuf> bufs;
    size_t cursor = 0;
    void reset() { cursor = 0; }
};

}  // namespace

struct colq_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::vector<Table> tables;
    std::map<std::string, int> registry;
    std::vector<colq_query*> queries;  // live queries; destroyed with the context
    std::string err;
    // communicator
    NcclApi nccl;
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0;
    bool str_attr_set = false;
    // peer-memory mailboxes (CUDA IPC over NVLink); ok == false -> NCCL collectives on the data path
    struct PeerBox {
        bool ok = false;
        bool ipc = true;   // peer pointers come from cudaIpcOpenMemHandle (one process per GPU); false: one process drives
                           // all GPUs and the pointers are the peers' own allocations (colq_comm_init_local)
        void* local = nullptr;
        size_t bytes = 0;
        void* peer_ptr[MAX_RANKS] = {};
        uint8_t** d_peers = nullptr;
        u32* d_done = nullptr;
        u32* d_status = nullptr;
        u64 mask_epoch = 0, gather_epoch = 0;
        int64_t slot_cap = 0;
        size_t slot_bytes = 0;
        // peer-mapped heap behind the mailbox: global bitmaps of cross-shard hops, two halves used alternately by
        // successive executions (a rank can be at most one execution ahead of a peer)
        size_t heap_off = 0, heap_half = 0;
        u64 heap_step = 0;
    } peer;
    // pinned [header | first FETCH_SPEC indices] staging of colq_fetch: small results come back with ONE sync.  Owned by
    // the context, allocated once: cudaHostAlloc / cudaFreeHost per query cost up to hundreds of ms on some hosts
    void* h_stage = nullptr;
    int compact_grid[2 * (CF_MAX_GATHER + 1)] = {};  // co-resident grid of compact_fused_kernel<NG, GATHER>
    int root_fused_grid[SR_MAX_PRED + 1] = {};       // resident CTAs of root_fused_kernel<NP> on this device
    int root_finish_grid = 0;                        // resident CTAs of root_finish_kernel
    u32* d_tile_counters = nullptr;                  // scan_str's tile-claim counter pair (zero between launches)
    colq_query* chain_query = nullptr;               // the last thing enqueued on the stream was this query's pipelining root kernel
    std::map<std::pair<int, size_t>, int> str_occupancy;  // (kernel mode, dynamic smem bytes) -> resident CTAs per SM
};

struct colq_query {
    colq_ctx* ctx = nullptr;
    std::string table_name;
    std::vector<QNode> nodes;
    int opt_lazy = 1, opt_profile = 0, opt_peer = 1, opt_fused_compact = 1, opt_defer = 1, opt_promote = 2, opt_fused_gather = 1, opt_tail_publish = 1, opt_root_fused = 1, opt_lazy_gather_wait = 1, opt_pipeline = 1;
    u32* clean_ready = nullptr;   // the push target the last execution's root kernel left zeroed (COLQ_OPT_PIPELINE)
    int64_t clean_words = 0;      //   ... and how many words of it
    bool timed = true;            // ev_start / ev_stop were recorded for the last execution
    std::vector<GatherD> deferred;  // root-node FK chains resolved by the compaction kernel instead of the row scan
    int own_begin = -1, own_end = -1;  // root-node scan ops that depend on no child (hoistable behind a mask publish)
    std::vector<Column*> pending_promotions;  // host-resident columns whose HBM copy this execution fills
    bool lazy_oob = false;  // the plan walks a to-one column that was not range-checked at ingest
    size_t heap_cursor = 0;  // bytes of the peer heap half this plan uses
    bool heap_used = false;
    int heap_parity = 0;
    int64_t promoted_bytes = 0;
    DevBuf barrier_buf;  // {arrival count, generation} of the cooperative compaction kernel
    DevBuf rf_state_buf;  // root_fused_kernel: [ticket, finished counters | pad to 64 B | one u64 state per CTA]
    int64_t rf_state_ctas = 0;
    u32 rf_epoch = 0;
    // the final gather of the last plan left every rank's indices in the mailbox slots (written by the kernel that
    // produced them); the concatenation into one list runs when the host fetches the indices
    u64* rf_dbg = nullptr;
    int rf_dbg_ctas = 0;
    bool gather_lazy = false;
    PeerGatherParams lazy_pg{};
    DevBuf lookback_buf;  // [2 counters | pad | tile states] of the single-pass compaction kernel (zeroed once)
    int64_t lookback_tiles = 0;
    u32 lookback_epoch = 0;
    // execution state
    Pool pool;
    std::vector<XNode> xnodes;
    std::vector<Op> ops;
    int root_table = -1;
    u32* root_bits = nullptr;
    // result block in HBM: [u64 count | 8 B pad | idx_capacity int32 row indices]; grows on demand, outside the pool
    DevBuf idx_buf;
    u64* d_total = nullptr;
    int32_t* d_idx = nullptr;
    int64_t idx_capacity = 0;
    int64_t want_idx_capacity = 0;  // 0 = pick a default at first execute
    // multi-GPU final gather: all ranks' result blocks, then their valid prefixes concatenated in rank order
    bool gathered = false;
    bool gather_is_peer = false;   // the final gather of the last plan runs over peer memory (fused or two launches)
    int64_t gather_block_cap = 0;  // indices one rank can contribute to it
    DevBuf gather_buf, gout_buf, ginfo_buf;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    std::vector<cudaEvent_t> stage_ev;
    std::vector<colq_stage> stages;
    // COLQ_OPT_PROFILE == 2: one event pair per execution around the launch with the most algorithmic bytes, kept
    // for every step of a timed region and averaged by colq_profile_hot (bench.py's live roofline figure)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> hot_ring;
    size_t hot_used = 0;
    colq_stage hot_stage{};
    colq_timing timing{};
    bool executed = false;
    int64_t local_count = -1;     // this rank's matching rows after the last fetch (-1: not fetched yet)
    DevBuf mat_a, mat_b;          // scratch of the result-materialisation gathers
};

namespace {

// =====================================================================================================
// error helpers
// =====================================================================================================

colq_status fail(colq_ctx* ctx, colq_status st, const char* fmt, ...) {
    char buf[768];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return st;
}

#define CU(ctx, expr)                                                                                         \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return fail((ctx), COLQ_ERR_DEVICE, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                        __LINE__, cudaGetErrorString(_e));                                                    \
    } while (0)

#define NC(ctx, expr)                                                                                   \
    do {                                                                                                \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != 0)                                                                                    \
            return fail((ctx), COLQ_ERR_DEVICE, "NCCL error %d at %s:%d: %s", _r, __FILE__, __LINE__,   \
                        (ctx)->nccl.GetErrorString ? (ctx)->nccl.GetErrorString(_r) : "?");             \
    } while (0)

#define ST(expr)                              \
    do {                                      \
        colq_status _s = (expr);              \
        if (_s != COLQ_OK) return _s;         \
    } while (0)

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// bitmask allocation: whole compaction tiles (131072 rows) so every kernel can store / vector-load full lines
inline int64_t bitmap_alloc_words(int64_t n_rows) { return round_up(std::max<int64_t>(n_rows, 1), 32 * CF_WORDS_PER_TILE) / 32; }
inline int64_t bitmap_words(int64_t n_rows) { return (n_rows + 31) / 32; }

colq_status dev_alloc(colq_ctx* ctx, DevBuf& b, size_t bytes) {
    // re-growing a live buffer: work already enqueued on this context's stream may still use the old block, and the
    // cache could hand it to ANOTHER context (another stream) right away -- drain first (rare: first executions only)
    if (b.owned && b.ptr) CU(ctx, cudaStreamSynchronize(ctx->stream));
    b.release();
    bytes = std::max<size_t>(bytes, 16);
    DeviceCache& cache = device_cache();
    size_t got = 0;
    void* p = cache.take(bytes, &got);
    if (!p) {
        got = bytes;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {  // out of memory: give the parked blocks back and retry once
            cudaGetLastError();
            cache.trim();
            CU(ctx, cudaMalloc(&p, bytes));
        }
    }
    b.ptr = p; b.bytes = got; b.owned = true;
    return COLQ_OK;
}

colq_status pool_alloc(colq_query* q, size_t bytes, void** out) {
    Pool& pl = q->pool;
    if (pl.cursor == pl.bufs.size()) pl.bufs.emplace_back();
    DevBuf& b = pl.bufs[pl.cursor++];
    if (b.bytes < bytes) ST(dev_alloc(q->ctx, b, bytes));
    *out = b.ptr;
    return COLQ_OK;
}

Table* get_table(colq_ctx* ctx, colq_table t) {
    if (t < 0 || (size_t)t >= ctx->tables.size()) return nullptr;
    return &ctx->tables[t];
}

colq_status slot_for(colq_ctx* ctx, colq_table t, int ordinal, int64_t n, Column** out) {
    Table* tb = get_table(ctx, t);
    if (!tb) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown table handle %d", t);
    if (ordinal < 0) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "negative column ordinal %d", ordinal);
    if (n != tb->n_rows)
        return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "column height %lld does not match the table's %lld rows", (long long)n,
                    (long long)tb->n_rows);
    if ((size_t)ordinal >= tb->cols.size()) tb->cols.resize(ordinal + 1);
    if (tb->cols[ordinal].kind != COL_UNSET)
        return fail(ctx, COLQ_THROW_ILLEGAL_STATE, "column %d of table %d is already set", ordinal, t);
    *out = &tb->cols[ordinal];
    return COLQ_OK;
}

const char* java_class_name(ColKind k) {
    switch (k) {
        case COL_I32: return "dgroomes.in_memory.InMemoryColumn$IntegerColumn";
        case COL_STR: return "dgroomes.in_memory.InMemoryColumn$StringColumn";
        case COL_BOOL: return "dgroomes.in_memory.InMemoryColumn$BooleanColumn";
        default: return "dgroomes.in_memory.InMemoryColumn$AssociationColumn";
    }
}

// =====================================================================================================
// verify: E/Verifier.java:40-111 (same traversal order, same messages)
// =====================================================================================================

colq_status verify(colq_query* q) {
    colq_ctx* ctx = q->ctx;
    auto it = ctx->registry.find(q->table_name);
    if (it == ctx->registry.end())  // E/DataSystemSerialIndices.java:54-57
        return fail(ctx, COLQ_FAILURE, "The query targets the table '%s' but that table is not registered",
                    q->table_name.c_str());
    q->root_table = it->second;
    q->xnodes.clear();
    q->xnodes.reserve(q->nodes.size());
    std::deque<std::pair<int, int>> to_visit;  // (query node, execution node): add = tail, pop = head (:49-54)
    q->xnodes.emplace_back();
    q->xnodes[0].table = q->root_table;
    to_visit.emplace_back(0, 0);
    while (!to_visit.empty()) {
        auto [qi, xi] = to_visit.front();
        to_visit.pop_front();
        const QNode& qn = q->nodes[qi];
        const Table& tb = ctx->tables[q->xnodes[xi].table];
        const int width = (int)tb.cols.size();
        for (const Crit& c : qn.crit) {
            if (width < c.ordinal)  // sic: `<` (:62); ordinal == width falls through to columns().get()
                return fail(ctx, COLQ_FAILURE, "The query ordinal '%d' is out of bounds for the table with %d columns",
                            c.ordinal, width);
            if (c.ordinal < 0 || c.ordinal >= width)  // columns().get(ordinal) (:67)
                return fail(ctx, COLQ_THROW_INDEX_OOB, "Index %d out of bounds for length %d", c.ordinal, width);
            const Column& col = tb.cols[c.ordinal];
            switch (col.kind) {  // switch (column.filterableType()) (:71-90)
                case COL_STR:
                    if (!c.is_str || c.is_bool)
                        return fail(ctx, COLQ_FAILURE, "The column is a string column but the criterion is not a string predicate.");
                    if (c.is_accept && !col.dict)
                        return fail(ctx, COLQ_FAILURE, "An opaque string predicate can only run over a dictionary-encoded column (colq_col_str_dict): it is evaluated per distinct value on the host; there is no CPU fallback for the row scan.");
                    if (c.is_accept && c.accept_n != col.dict->n)
                        return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "accept set has %lld entries but the column's dictionary has %lld", (long long)c.accept_n, (long long)col.dict->n);
                    break;
                case COL_I32:
                    if (c.is_str || c.is_bool)
                        return fail(ctx, COLQ_FAILURE, "The column is an integer column but the criterion is not an integer predicate.");
                    if (c.is_accept && !col.dict)
                        return fail(ctx, COLQ_FAILURE, "An opaque integer predicate can only run over a dictionary-encoded column (colq_col_i32_dict): it is evaluated per distinct value on the host; there is no CPU fallback for the row scan.");
                    if (c.is_accept && c.accept_n != col.dict->n)
                        return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "accept set has %lld entries but the column's dictionary has %lld", (long long)c.accept_n, (long long)col.dict->n);
                    break;
                case COL_BOOL:
                    // the reference stops here for every criterion (:82-84); Criteria has no boolean member, so an int
                    // or string criterion on a boolean column is all it can be asked.  colq_query_criteria_bool is the
                    // 8(f4) extension.
                    if (!c.is_bool) return fail(ctx, COLQ_FAILURE, "Boolean columns are not supported yet.");
                    break;
                case COL_ASSOC:
                    return fail(ctx, COLQ_FAILURE, "Association columns can't be matched on with a scalar criteria.");
                default:
                    return fail(ctx, COLQ_THROW_ILLEGAL_STATE, "column %d was never set", c.ordinal);
            }
            q->xnodes[xi].preds.push_back(&c);  // addColumnPredicate (:92)
        }
        for (auto [ordinal, child_q] : qn.children) {  // Map.copyOf iteration order is unspecified; results are order-free
            if (ordinal < 0 || ordinal >= width)       // columns().get(ordinal), unchecked (:100)
                return fail(ctx, COLQ_THROW_INDEX_OOB, "Index %d out of bounds for length %d", ordinal, width);
            const Column& col = tb.cols[ordinal];
            if (col.kind != COL_ASSOC)  // (:102-104)
                return fail(ctx, COLQ_FAILURE, "The column at ordinal %d is not an association column. It is a %s", ordinal,
                            java_class_name(col.kind));
            // createChildNode: Node(assoc.associatedEntity(), this, assoc.reverseAssociatedColumn()) (E/ExecutionContext.java:64-68)
            XNode cx;
            cx.table = col.peer_table;
            cx.parent = xi;
            cx.parent_ordinal = ordinal;
            q->xnodes.push_back(cx);
            int ci = (int)q->xnodes.size() - 1;
            q->xnodes[xi].children.emplace_back(ordinal, ci);
            to_visit.emplace_back(child_q, ci);
        }
    }
    return COLQ_OK;
}

// =====================================================================================================
// planner: post-order evaluation of the execution-node tree into a linear op list
// =====================================================================================================

struct Consume {
    bool push = false;          // false: the parent needs this node's bitmask materialised
    const Column* fwd = nullptr;  // forward column on THIS node's table pointing at the parent's rows
    u32* reach = nullptr;
    int64_t n_parent = 0;
};

struct NodeBits {
    u32* bits = nullptr;
    bool all_ones = false;
};

struct Planner {
    colq_query* q;
    colq_ctx* ctx;

    colq_status alloc_bitmap(int64_t n_rows, u32** out) {
        void* p;
        ST(pool_alloc(q, (size_t)bitmap_alloc_words(n_rows) * 4, &p));
        *out = (u32*)p;
        return COLQ_OK;
    }

    bool sharded(const Table& t) const { return ctx->n_ranks > 1 && t.placement == COLQ_SHARDED; }

    u32* oob_flag() const { return (u32*)q->idx_buf.ptr + RESULT_FLAGS_WORD; }

    // a to-one column about to be walked by a kernel: unvalidated (host-resident) ones report bad targets lazily
    u32* oob_for(const Column& fk_col) {
        if (fk_col.fk_validated) return nullptr;
        q->lazy_oob = true;
        return oob_flag();
    }

    // first-touch promotion of a host-resident column that the next launch reads in full: allocate the HBM copy the
    // kernel fills.  Out of device memory is not an error -- the column simply keeps being streamed over PCIe.
    // COLQ_OPT_PROMOTE=2 (default): a host-resident column that the next launch reads IN FULL is first brought to HBM
    // by the copy engine (one cudaMemcpyAsync on the query's stream, ahead of the kernels -- 55 GB/s against the
    // 49-50 GB/s a kernel reaches reading mapped host memory), and the launch then runs on the HBM copy at HBM speed.
    // Nothing else of the table moves.  Falls back to in-place streaming when the allocation fails.
    void upload_on_first_scan(const Column& col_c, size_t data_bytes, size_t data_copy, size_t offsets_bytes, size_t offsets_copy) {
        Column& col = const_cast<Column&>(col_c);
        if (!col.host_resident || q->opt_promote != 2 || data_copy == 0) return;
        DevBuf d, o;
        if (dev_alloc(ctx, d, data_bytes) != COLQ_OK || (offsets_bytes && dev_alloc(ctx, o, offsets_bytes) != COLQ_OK)) {
            cudaGetLastError();
            return;
        }
        cudaStream_t s = ctx->stream;
        cudaMemsetAsync((char*)d.ptr + (data_bytes - 64), 0, 64, s);
        cudaMemcpyAsync(d.ptr, col.data.ptr, data_copy, cudaMemcpyHostToDevice, s);
        if (offsets_bytes) {
            cudaMemsetAsync((char*)o.ptr + (offsets_bytes - 32), 0, 32, s);
            cudaMemcpyAsync(o.ptr, col.offsets.ptr, offsets_copy, cudaMemcpyHostToDevice, s);
        }
        q->timing.h2d_bytes += (int64_t)(data_copy + offsets_copy);
        q->promoted_bytes += (int64_t)(data_bytes + offsets_bytes);
        col.data = std::move(d);
        if (offsets_bytes) {
            col.offsets = std::move(o);
            col.bytes_capacity = (int64_t)(col.data.bytes & ~(size_t)15);
        }
        col.host_resident = false;
    }

    bool want_promotion(const Column& col_c, size_t data_bytes, size_t offsets_bytes) {
        Column& col = const_cast<Column&>(col_c);
        if (!col.host_resident || !q->opt_promote) return false;
        if (!col.promoted.ptr) {
            DevBuf d, o;
            if (dev_alloc(ctx, d, data_bytes) != COLQ_OK || (offsets_bytes && dev_alloc(ctx, o, offsets_bytes) != COLQ_OK)) {
                cudaGetLastError();
                return false;
            }
            // the kernels write whole 16-byte lines of real data; zero the padding behind them once
            cudaMemsetAsync((char*)d.ptr + (data_bytes - 64), 0, 64, ctx->stream);
            if (offsets_bytes) cudaMemsetAsync((char*)o.ptr + (offsets_bytes - 32), 0, 32, ctx->stream);
            col.promoted = std::move(d);
            col.promoted_offsets = std::move(o);
        }
        q->pending_promotions.push_back(&col);
        q->promoted_bytes += (int64_t)(data_bytes + offsets_bytes);
        return true;
    }

    bool lazy_eligible(int xi) const {
        const XNode& x = q->xnodes[xi];
        if (!q->opt_lazy || !x.preds.empty() || x.children.size() != 1) return false;
        const Table& t = ctx->tables[x.table];
        const Column& col = t.cols[x.children[0].first];
        if (!(col.forward && col.is_fk)) return false;
        const Table& ct = ctx->tables[col.peer_table];
        if (!sharded(t) && sharded(ct)) return false;
        if (col.global_targets && sharded(ct)) return false;  // the hop leaves the shard: its child is materialised and exchanged
        return true;
    }

    // a bitmap over the GLOBAL rows of a sharded table, at the same offset of the peer-mapped heap on every rank
    colq_status heap_bitmap(int64_t global_rows, u32** local, size_t* off) {
        auto& pb = ctx->peer;
        if (!pb.ok || pb.heap_half == 0)
            return fail(ctx, COLQ_FAILURE, "unsupported placement: an association hop between shards needs the peer-memory exchange (NVLink P2P mailboxes), which is not available on this communicator");
        const size_t bytes = (size_t)round_up(bitmap_words(std::max<int64_t>(global_rows, 1)) * 4 + 16, 256);
        if (!q->heap_used) {
            q->heap_used = true;
            q->heap_parity = (int)(pb.heap_step++ & 1);
        }
        if (q->heap_cursor + bytes > pb.heap_half)
            return fail(ctx, COLQ_ERR_CAPACITY, "the global bitmaps of this query's cross-shard hops need %zu bytes of peer heap but a half holds %zu: raise COLQ_PEER_HEAP_MB",
                        q->heap_cursor + bytes, pb.heap_half);
        *off = pb.heap_off + (size_t)q->heap_parity * pb.heap_half + q->heap_cursor;
        *local = (u32*)((char*)pb.local + *off);
        q->heap_cursor += bytes;
        return COLQ_OK;
    }

    PeerBitsParams peer_bits(const Table& owner, size_t heap_off) const {
        PeerBitsParams P{};
        P.heap_off = heap_off;
        P.word_base = owner.part[ctx->rank] / 32;
        P.n_words = bitmap_words(owner.n_rows);
        P.n_ranks = ctx->n_ranks; P.rank = ctx->rank; P.n_src = ctx->n_ranks;
        P.peers = ctx->peer.d_peers; P.status = ctx->peer.d_status; P.done = ctx->peer.d_done + MAX_RANKS + 9;
        return P;
    }

    colq_status eval(int xi, const Consume& consume, NodeBits* out) {
        XNode& x = q->xnodes[xi];
        const Table& T = ctx->tables[x.table];
        const int64_t n = T.n_rows;
        const size_t first_op = q->ops.size();

        std::vector<GatherD> gathers;
        struct CsrIn { const Column* col; NodeBits child; int64_t n_child; };
        std::vector<CsrIn> csrs;
        u32* cur = nullptr;  // null: every row still matches

        for (auto [ordinal, ci] : std::vector<std::pair<int, int>>(x.children)) {
            const Column& col = T.cols[ordinal];
            const Table& CT = ctx->tables[col.peer_table];
            if (col.forward) {
                const bool xs = col.global_targets && sharded(CT);  // the keys are GLOBAL rows of a sharded table
                if (!sharded(T) && sharded(CT) && !xs)
                    return fail(ctx, COLQ_FAILURE, "unsupported placement: a replicated table holds an association with shard-local targets into a sharded table (register it with colq_associate_*_global)");
                if (xs) {
                    // cross-shard PULL: materialise the child on its owners, all-gather its bits into a global bitmap on
                    // every rank, then test bit [global key] like any local one
                    if (CT.part.empty()) return fail(ctx, COLQ_THROW_ILLEGAL_STATE, "the sharded target table has no partition (colq_table_partition)");
                    NodeBits nb;
                    ST(eval(ci, Consume{}, &nb));
                    const int64_t gy = CT.global_rows();
                    u32* gbits = nullptr;
                    if (!nb.all_ones) {
                        size_t off = 0;
                        ST(heap_bitmap(gy, &gbits, &off));
                        Op ag{};
                        ag.kind = K_PEER_BITS_ALLGATHER; ag.node = xi; ag.name = "peer_bits_allgather";
                        ag.pbits = peer_bits(CT, off);
                        ag.pbits.src = nb.bits;
                        ag.acct_rows = CT.n_rows; ag.acct_bytes = ag.pbits.n_words * 4 * (ctx->n_ranks + 1);
                        q->ops.push_back(ag);
                    }
                    if (col.is_fk) {
                        GatherD g{};
                        g.fk[0] = (const int32_t*)col.data.ptr;
                        g.n[0] = gy;
                        g.depth = 1;
                        g.oob = oob_for(col);
                        g.bits = gbits;
                        gathers.push_back(g);
                    } else {
                        csrs.push_back({&col, NodeBits{gbits, gbits == nullptr}, gy});
                    }
                    continue;
                }
                if (col.is_fk) {
                    GatherD g{};
                    g.fk[0] = (const int32_t*)col.data.ptr;
                    g.n[0] = CT.n_rows;
                    g.depth = 1;
                    g.oob = oob_for(col);
                    int cj = ci;
                    while (g.depth < GATHER_MAX_DEPTH && lazy_eligible(cj)) {
                        XNode& c = q->xnodes[cj];
                        const Column& ccol = ctx->tables[c.table].cols[c.children[0].first];
                        g.fk[g.depth] = (const int32_t*)ccol.data.ptr;
                        g.n[g.depth] = ctx->tables[ccol.peer_table].n_rows;
                        if (u32* f = oob_for(ccol)) g.oob = f;
                        g.depth++;
                        c.fused = true;
                        cj = c.children[0].second;
                    }
                    NodeBits nb;
                    ST(eval(cj, Consume{}, &nb));
                    g.bits = nb.all_ones ? nullptr : nb.bits;
                    gathers.push_back(g);
                } else {
                    NodeBits nb;
                    ST(eval(ci, Consume{}, &nb));
                    csrs.push_back({&col, nb, CT.n_rows});
                }
            } else {
                // reverse side: the data is the child's forward column; the child pushes into `reach`
                const Column& f = CT.cols[col.peer_ordinal];
                const bool xs = f.global_targets && sharded(T);  // the child's keys are GLOBAL rows of this sharded table
                if (sharded(T) && !sharded(CT) && !xs)
                    return fail(ctx, COLQ_FAILURE, "unsupported placement: a replicated table holds an association with shard-local targets into a sharded table (register it with colq_associate_*_global)");
                u32* reach;
                ST(alloc_bitmap(n, &reach));
                if (xs) {
                    // cross-shard PUSH: the child sets bits in this rank's own GLOBAL-sized reach bitmap; every rank then
                    // ORs all ranks' copies of its own slice (OR-reduce-scatter by remote loads)
                    if (T.part.empty()) return fail(ctx, COLQ_THROW_ILLEGAL_STATE, "the sharded table has no partition (colq_table_partition)");
                    const int64_t gt = T.global_rows();
                    u32* reach_g = nullptr;
                    size_t off = 0;
                    ST(heap_bitmap(gt, &reach_g, &off));
                    Op z{};
                    z.kind = K_ZERO; z.node = xi; z.dst = reach_g; z.n_alloc_words = bitmap_words(std::max<int64_t>(gt, 1)) + 4; z.name = "memset_reach";
                    q->ops.push_back(z);
                    Consume cc;
                    cc.push = true; cc.fwd = &f; cc.reach = reach_g; cc.n_parent = gt;
                    NodeBits nb;
                    ST(eval(ci, cc, &nb));
                    Op rd{};
                    rd.kind = K_PEER_BITS_REDUCE; rd.node = xi; rd.name = "peer_bits_reduce";
                    rd.pbits = peer_bits(T, off);
                    rd.pbits.dst = reach;
                    rd.pbits.n_src = sharded(CT) ? ctx->n_ranks : 1;  // a replicated child pushed the same bits on every rank
                    rd.acct_rows = n; rd.acct_bytes = rd.pbits.n_words * 4 * (rd.pbits.n_src + 1);
                    q->ops.push_back(rd);
                    if (cur == nullptr) cur = reach;
                    else {
                        Op a{};
                        a.kind = K_AND; a.node = xi; a.dst = cur; a.src = reach; a.n_words = bitmap_words(n); a.name = "and_words";
                        a.acct_rows = n; a.acct_bytes = a.n_words * 12;
                        q->ops.push_back(a);
                    }
                    continue;
                }
                Op z{};
                z.kind = K_ZERO; z.node = xi; z.dst = reach; z.n_alloc_words = bitmap_alloc_words(n); z.name = "memset_reach";
                q->ops.push_back(z);
                Consume cc;
                cc.push = true; cc.fwd = &f; cc.reach = reach; cc.n_parent = n;
                NodeBits nb;
                ST(eval(ci, cc, &nb));
                if (sharded(CT) && !sharded(T)) {
                    // the only data-path collective: OR-allreduce of the replicated parent's mask
                    Op g{};
                    g.node = xi; g.dst = reach; g.n_words = bitmap_words(n);
                    const bool local_group = ctx->comm == nullptr;   // one process, no NCCL: peer memory is the only exchange
                    if (local_group && g.n_words > MASK_WORDS_MAX)
                        return fail(ctx, COLQ_FAILURE, "a replicated table of %lld rows is too large for the peer-memory mask exchange of a local communicator (limit %d rows)",
                                    (long long)n, MASK_WORDS_MAX * 32);
                    if (ctx->peer.ok && (q->opt_peer || local_group) && g.n_words <= MASK_WORDS_MAX) {
                        g.kind = K_PEER_MASK_PUBLISH; g.name = "peer_mask_publish";
                        PeerMaskParams& P = g.pmask;
                        P.reach = reach; P.n_words = (int)g.n_words; P.n_ranks = ctx->n_ranks; P.rank = ctx->rank;
                        P.peers = ctx->peer.d_peers; P.status = ctx->peer.d_status;
                        g.acct_bytes = g.n_words * 4 * ctx->n_ranks;
                        q->ops.push_back(g);
                        g.kind = K_PEER_MASK_COLLECT; g.name = "peer_mask_collect";
                        g.publish_op = (int)q->ops.size() - 1;
                    } else {
                        g.kind = K_ALLGATHER_OR; g.name = "allgather_or_mask";
                        void* gb;
                        ST(pool_alloc(q, (size_t)g.n_words * 4 * ctx->n_ranks, &gb));
                        g.gathered = (u32*)gb;
                        g.acct_bytes = g.n_words * 4 * ctx->n_ranks;
                    }
                    q->ops.push_back(g);
                }
                if (cur == nullptr) cur = reach;
                else {
                    Op a{};
                    a.kind = K_AND; a.node = xi; a.dst = cur; a.src = reach; a.n_words = bitmap_words(n); a.name = "and_words";
                    a.acct_rows = n; a.acct_bytes = a.n_words * 12;
                    q->ops.push_back(a);
                }
            }
        }

        XNode& xr = q->xnodes[xi];  // (children evals may not reallocate xnodes, but re-bind for clarity)
        u32* own = nullptr;         // this node's bitmask buffer, allocated on first use
        auto out_buf = [&](u32** p) -> colq_status {
            if (!own) ST(alloc_bitmap(n, &own));
            *p = own;
            return COLQ_OK;
        };

        const bool own_independent = (xi == 0 && cur == nullptr && csrs.empty());  // so far no child touched `cur`
        const size_t own_first = q->ops.size();
        // ---- string criteria: one TMA-staged scan each
        auto make_scan_str = [&](const Column& col, int64_t rows, const Crit* c, const u32* in_bits, u32* out_bits) -> Op {
            Op o{};
            o.kind = K_SCAN_STR; o.node = xi; o.name = "scan_str";
            ScanStrParams& P = o.str;
            P.n = rows;
            if (rows > 0)
                upload_on_first_scan(col, (size_t)round_up(col.n_bytes, 16) + 64, (size_t)col.n_bytes, (size_t)round_up((rows + 1) * 4, 16) + 32,
                                     (size_t)(rows + 1) * 4);
            P.offsets = (const u32*)col.offsets.ptr;
            P.bytes = (const uint8_t*)col.data.ptr;
            P.bytes_capacity = col.bytes_capacity;
            P.needle = (const uint8_t*)c->needle_dev.ptr;
            P.needle_len = (int)c->needle.size();
            P.op = c->op;
            // ring geometry: a slot holds the column's largest tile payload (measured at ingest) plus the two 16-byte
            // alignment margins, so every tile is staged unless it exceeds the 40 KB limit (then it is read from global
            // memory); 3 slots per CTA leave room for 4-5 resident CTAs per SM, which is what saturates HBM (r01 sweep
            // in profiles/r01_scan_str_ring_sweep.txt)
            static const int stages_env = getenv("COLQ_STR_STAGES") ? atoi(getenv("COLQ_STR_STAGES")) : 3;
            int64_t cap = std::min<int64_t>(std::max<int64_t>(col.max_tile_bytes + 48, 2048), 40960);
            P.cap = (int)round_up(cap, 16);
            P.stages = std::max(2, std::min(stages_env, (int)ST_MAX_STAGES));
            P.n_tiles = (rows + ST_ROWS - 1) / ST_ROWS;
            if (col.host_resident) q->timing.h2d_bytes += (rows + 1) * 4 + col.n_bytes;  // streamed over PCIe by this launch
            if (P.n_tiles > 0 && want_promotion(col, (size_t)round_up(col.n_bytes, 16) + 64, (size_t)round_up((rows + 1) * 4, 16) + 32)) {
                P.promote_bytes = (uint8_t*)col.promoted.ptr;
                P.promote_offsets = (u32*)col.promoted_offsets.ptr;
            }
            P.in_bits = in_bits;
            P.out_bits = out_bits;
            P.tile_counter = ctx->d_tile_counters;
            o.smem = (size_t)P.stages * st_stage_bytes(P.cap) + st_needle_region(P.needle_len) + 2 * ST_MAX_STAGES * 8 +
                     ST_MAX_STAGES * sizeof(StrTileMeta) + PUSH_SMEM_WORDS * 4;
            o.acct_rows = rows;
            o.acct_bytes = (rows + 1) * 4 + col.n_bytes + bitmap_words(rows) * 4;
            return o;
        };
        for (const Crit* c : xr.preds) {
            const Column& col = T.cols[c->ordinal];
            if (c->is_bool) {  // one byte per row against the predicate's truth table
                u32* ob;
                ST(out_buf(&ob));
                Op o{};
                o.kind = K_SCAN_BOOL; o.node = xi; o.name = "scan_bool";
                o.boolp.n = n;
                o.boolp.n_alloc_words = bitmap_alloc_words(n);
                o.boolp.values = (const uint8_t*)col.data.ptr;
                o.boolp.accept_false = c->accept_false;
                o.boolp.accept_true = c->accept_true;
                o.boolp.in_bits = cur;
                o.boolp.out_bits = ob;
                o.acct_rows = n;
                o.acct_bytes = n + bitmap_words(n) * 4;
                q->ops.push_back(o);
                cur = ob;
                continue;
            }
            if (!c->is_str && !col.dict) continue;  // plain int columns: fused row scans below
            u32* ob;
            ST(out_buf(&ob));
            if (!col.dict) {
                q->ops.push_back(make_scan_str(col, n, c, cur, ob));
                cur = ob;
                continue;
            }
            // dictionary-encoded column: evaluate the predicate once per DISTINCT value, then test bit `code` per row
            const Column& D = *col.dict;
            const u32* accept = (const u32*)c->accept_dev.ptr;  // opaque predicate: the host evaluated it per entry
            if (!c->is_accept) {
                u32* dbits;
                ST(alloc_bitmap(D.n, &dbits));
                if (D.n == 0 || (!c->is_str && c->lo > c->hi)) {  // nothing can match; the mask must still be defined
                    Op z{};
                    z.kind = K_ZERO; z.node = xi; z.dst = dbits; z.n_alloc_words = bitmap_alloc_words(D.n); z.name = "memset_dict_mask";
                    q->ops.push_back(z);
                } else if (c->is_str) {
                    Op ds = make_scan_str(D, D.n, c, nullptr, dbits);
                    ds.name = "scan_str_dictionary";
                    q->ops.push_back(ds);
                } else {  // the closed interval over the DISTINCT values
                    Op ds{};
                    ds.kind = K_SCAN_ROWS; ds.node = xi; ds.name = "scan_rows"; ds.np = 1;
                    ds.rows.n = D.n;
                    ds.rows.out_bits = dbits;
                    ds.rows.pred[0].col = (const int32_t*)D.data.ptr;
                    ds.rows.pred[0].lo = c->lo;
                    ds.rows.pred[0].span = (u32)((int64_t)c->hi - (int64_t)c->lo);
                    ds.dict_scan = true;
                    ds.acct_rows = D.n; ds.acct_bytes = D.n * 4 + bitmap_words(D.n) * 4;
                    q->ops.push_back(ds);
                }
                accept = dbits;
            }
            Op o{};
            o.kind = K_SCAN_CODES; o.node = xi; o.name = "scan_codes";
            ScanCodesParams& P = o.codes;
            P.n = n;
            if (D.n > 0 && n > 0) upload_on_first_scan(col, (size_t)round_up(n * 4 + 16, 16) + 64, (size_t)n * 4, 0, 0);
            P.codes = (const int32_t*)col.data.ptr;
            P.accept = accept;
            P.n_dict = (u32)D.n;
            const int64_t mask_words = (bitmap_words(D.n) + 3) & ~(int64_t)3;
            P.mask_words = mask_words <= SC_SMEM_MASK_WORDS ? (int)mask_words : 0;
            P.in_bits = cur;
            P.out_bits = ob;
            if (D.n == 0) o.never = true;
            else {
                if (col.host_resident) q->timing.h2d_bytes += n * 4;
                if (n > 0 && want_promotion(col, (size_t)round_up(n * 4 + 16, 16) + 64, 0)) P.promote = (int32_t*)col.promoted.ptr;
            }
            o.acct_rows = n;
            o.acct_bytes = n * 4 + bitmap_words(n) * 4;
            q->ops.push_back(o);
            cur = ob;
        }

        // ---- int criteria + forward to-one hops: fused row scans, at most 2 predicates and 2 chains per launch
        std::vector<const Crit*> ints;
        for (const Crit* c : xr.preds)
            if (!c->is_str && !c->is_bool && !T.cols[c->ordinal].dict) ints.push_back(c);
        size_t pi = 0, gi = 0;
        // The root's criteria-free to-one chains need not stall the streaming scan: when a predicate (or an earlier
        // mask) already thins the rows out, hand the chains to the fused compaction, which walks them for the
        // surviving bits only, spread over the whole grid (COLQ_OPT_DEFER_CHAINS).
        if (xi == 0 && !consume.push && q->opt_defer && q->opt_lazy && q->opt_fused_compact && !gathers.empty() &&
            gathers.size() <= (size_t)CF_MAX_GATHER && (!ints.empty() || cur != nullptr)) {
            q->deferred = gathers;
            gi = gathers.size();
        }
        // an empty closed interval anywhere in the AND: no row of this node matches, every launch degenerates to clearing
        // its mask -- and must not stream, upload or promote any column (a promotion would never be filled)
        bool node_never = false;
        for (const Crit* c : ints) node_never = node_never || c->lo > c->hi;
        while (pi < ints.size() || gi < gathers.size()) {
            Op o{};
            o.kind = K_SCAN_ROWS; o.node = xi; o.name = "scan_rows";
            ScanRowsParams& P = o.rows;
            P.n = n;
            P.in_bits = cur;
            int64_t bytes = 0;
            while (pi < ints.size() && o.np < SR_MAX_PRED) {
                const Crit* c = ints[pi++];
                const Column& col = T.cols[c->ordinal];
                IntPredD& d = P.pred[o.np++];
                if (!node_never && n > 0) upload_on_first_scan(col, (size_t)round_up(n * 4 + 16, 16) + 64, (size_t)n * 4, 0, 0);
                d.col = (const int32_t*)col.data.ptr;
                if (node_never) {  // empty closed interval: no value satisfies it
                    o.never = true;
                    d.lo = 0; d.span = 0;
                } else {
                    d.lo = c->lo;
                    d.span = (u32)((int64_t)c->hi - (int64_t)c->lo);
                    if (col.host_resident) q->timing.h2d_bytes += n * 4;
                    if (n > 0 && want_promotion(col, (size_t)round_up(n * 4 + 16, 16) + 64, 0)) d.promote = (int32_t*)col.promoted.ptr;
                }
                bytes += n * 4;
            }
            while (gi < gathers.size() && o.ng < SR_MAX_GATHER) {
                P.gather[o.ng++] = gathers[gi++];
            }
            o.eager = (o.np == 0 && cur == nullptr);
            if (o.eager) bytes += (int64_t)o.ng * n * 4;
            u32* ob;
            ST(out_buf(&ob));
            P.out_bits = ob;
            o.acct_rows = n;
            o.acct_bytes = bytes + bitmap_words(n) * 4;
            q->ops.push_back(o);
            cur = ob;
        }

        if (own_independent && q->deferred.size() == gathers.size() && q->ops.size() > own_first) {
            q->own_begin = (int)own_first;
            q->own_end = (int)q->ops.size();
        }

        // ---- forward to-many hops
        for (const CsrIn& ci : csrs) {
            Op o{};
            o.kind = K_CSR_PULL; o.node = xi; o.name = "csr_pull";
            CsrPullParams& P = o.csr;
            P.n = n;
            P.offsets = (const int64_t*)ci.col->offsets.ptr;
            P.targets = (const int32_t*)ci.col->targets.ptr;
            P.child_bits = ci.child.all_ones ? nullptr : ci.child.bits;
            P.n_child = ci.n_child;
            P.nnz = ci.col->nnz;
            P.in_bits = cur;
            u32* ob;
            ST(out_buf(&ob));
            P.out_bits = ob;
            o.acct_rows = n;
            o.acct_bytes = (n + 1) * 8 + ci.col->nnz * 4 + bitmap_words(n) * 4;
            q->ops.push_back(o);
            cur = ob;
        }

        // ---- hand the result to the consumer
        if (consume.push) {
            Op* last = nullptr;
            if (q->ops.size() > first_op) {
                Op& l = q->ops.back();
                if (l.node == xi && (l.kind == K_SCAN_ROWS || l.kind == K_SCAN_CODES || l.kind == K_SCAN_STR || l.kind == K_CSR_PULL)) last = &l;
            }
            if (last && last->never) {
                // no row of this node matches: nothing to push, and nobody else reads the mask
                last->rows.out_bits = nullptr;
                last->codes.out_bits = nullptr;
                xr.bits = nullptr; xr.all_ones = false; xr.fused = true;
            } else if (last && consume.fwd->is_fk) {
                PushD pd{(const int32_t*)consume.fwd->data.ptr, consume.reach, consume.n_parent, oob_for(*consume.fwd)};
                if (last->kind == K_SCAN_ROWS) { last->rows.push = pd; last->rows.out_bits = nullptr; }
                else if (last->kind == K_SCAN_CODES) { last->codes.push = pd; last->codes.out_bits = nullptr; }
                else if (last->kind == K_SCAN_STR) { last->str.push = pd; last->str.out_bits = nullptr; }
                else { last->csr.push = pd; last->csr.out_bits = nullptr; }
                // out_bits dropped: nobody else reads this node's mask (debug cardinality reports -1)
                xr.bits = nullptr;
                xr.all_ones = false;
                xr.fused = true;
            } else {
                Op o{};
                o.kind = K_PUSH_BITS; o.node = xi; o.name = "push_bits";
                PushBitsParams& P = o.pushb;
                P.n_child = n;
                P.child_bits = cur;
                if (consume.fwd->is_fk) P.fk = (const int32_t*)consume.fwd->data.ptr;
                else {
                    P.offsets = (const int64_t*)consume.fwd->offsets.ptr;
                    P.targets = (const int32_t*)consume.fwd->targets.ptr;
                }
                P.reach = consume.reach;
                P.n_parent = consume.n_parent;
                P.oob = consume.fwd->is_fk ? oob_for(*consume.fwd) : nullptr;
                o.acct_rows = n;
                o.acct_bytes = bitmap_words(n) * 4 + (consume.fwd->is_fk ? n * 4 : consume.fwd->nnz * 4);
                q->ops.push_back(o);
                xr.bits = cur;
                xr.all_ones = (cur == nullptr);
            }
            if (out) *out = NodeBits{};
            return COLQ_OK;
        }
        xr.bits = cur;
        xr.all_ones = (cur == nullptr);
        if (out) { out->bits = cur; out->all_ones = (cur == nullptr); }
        return COLQ_OK;
    }
};

// =====================================================================================================
// launch
// =====================================================================================================

template <int NP, int NG>
void launch_scan_rows_e(const Op& o, cudaStream_t s) {
    int grid = (int)((o.rows.n + SR_BLOCK_ROWS - 1) / SR_BLOCK_ROWS);
    if (grid == 0) return;
    if (o.eager) scan_rows_kernel<NP, NG, true><<<grid, SR_THREADS, 0, s>>>(o.rows);
    else scan_rows_kernel<NP, NG, false><<<grid, SR_THREADS, 0, s>>>(o.rows);
}

void launch_scan_rows(const Op& o, cudaStream_t s) {
    switch (o.np * 3 + o.ng) {
        case 0: launch_scan_rows_e<0, 0>(o, s); break;
        case 1: launch_scan_rows_e<0, 1>(o, s); break;
        case 2: launch_scan_rows_e<0, 2>(o, s); break;
        case 3: launch_scan_rows_e<1, 0>(o, s); break;
        case 4: launch_scan_rows_e<1, 1>(o, s); break;
        case 5: launch_scan_rows_e<1, 2>(o, s); break;
        case 6: launch_scan_rows_e<2, 0>(o, s); break;
        case 7: launch_scan_rows_e<2, 1>(o, s); break;
        case 8: launch_scan_rows_e<2, 2>(o, s); break;
    }
}

void stage_name(const Op& o, char* out, size_t cap) {
    if (o.kind == K_SCAN_CODES) snprintf(out, cap, "scan_codes%s%s", o.codes.push.fk ? "+push" : "", o.tail_publish ? "+publish" : "");
    else if (o.kind == K_SCAN_ROWS && o.dict_scan) snprintf(out, cap, "scan_rows_dictionary");
    else if (o.kind == K_SCAN_ROWS && o.list) snprintf(out, cap, "scan_rows<%d,%d,list>", o.np, o.ng);
    else if (o.kind == K_SCAN_ROWS) snprintf(out, cap, "scan_rows<%d,%d,%s>%s", o.np, o.ng, o.eager ? "eager" : "lazy", o.rows.push.fk ? "+push" : "");
    else if (o.kind == K_SCAN_STR) snprintf(out, cap, "%s<op%d>%s%s", o.name, o.str.op, o.str.push.fk ? "+push" : "", o.tail_publish ? "+publish" : "");
    else if (o.kind == K_ROOT_FINISH)
        snprintf(out, cap, "root_finish<%d>%s%s", o.ng, o.rfused.pre.n > 0 ? (o.rfused.pre.pm.n_words > 0 ? "+collect+csr" : "+csr") : "",
                 o.rfused.pg.n_ranks > 0 ? "+gather" : "");
    else if (o.kind == K_ROOT_FUSED)
        snprintf(out, cap, "root_fused<%d,%d>%s%s%s", o.np, o.ng, o.rfused.pre.n > 0 ? (o.rfused.pre.pm.n_words > 0 ? "+collect+csr" : "+csr") : "",
                 o.rfused.pg.n_ranks > 0 ? "+gather" : "", "");
    else snprintf(out, cap, "%s", o.name);
}

inline const void* compact_fused_fn(int ng, bool gather = false) {
    switch (ng * 2 + (gather ? 1 : 0)) {
        case 0: return (const void*)compact_fused_kernel<0, false>;
        case 1: return (const void*)compact_fused_kernel<0, true>;
        case 2: return (const void*)compact_fused_kernel<1, false>;
        case 3: return (const void*)compact_fused_kernel<1, true>;
        case 4: return (const void*)compact_fused_kernel<2, false>;
        default: return (const void*)compact_fused_kernel<2, true>;
    }
}

inline int grid_for(int64_t work_items, int threads, int sm_count, int per_sm) {
    int64_t g = (work_items + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, (int64_t)sm_count * per_sm));
}

colq_status launch_op(colq_query* q, Op& o, cudaStream_t s, bool count_only = false) {
    colq_ctx* ctx = q->ctx;
    (void)count_only;
    switch (o.kind) {
        case K_SCAN_ROWS:
            if (o.never) {  // an empty int interval: no row matches
                if (o.rows.out_bits)
                    CU(ctx, cudaMemsetAsync(o.rows.out_bits, 0, (size_t)bitmap_alloc_words(o.rows.n) * 4, s));
            } else {
                // A/B variant (COLQ_SCAN_ROWS_TMA=1): the plain single-predicate scan staged through shared memory with TMA bulk
                // copies instead of per-lane LDG.128 (measured slower or equal on B200: profiles/r02_scan_rows_tma_ab.txt)
                if (o.list) {
                    // the root's scan in front of root_finish_kernel; a programmatic dependent of the string scan when it reads
                    // nothing an earlier launch wrote
                    static const bool pdl = !(getenv("COLQ_PDL") && getenv("COLQ_PDL")[0] == '0');
                    cudaLaunchConfig_t cfg{};
                    cfg.gridDim = dim3((unsigned)((o.rows.n + SR_BLOCK_ROWS - 1) / SR_BLOCK_ROWS)); cfg.blockDim = dim3(SR_THREADS); cfg.stream = s;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at[0].val.programmaticStreamSerializationAllowed = (pdl && o.rows.in_bits == nullptr) ? 1 : 0;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    if (cfg.gridDim.x > 0) {
                        if (o.np == 1) CU(ctx, cudaLaunchKernelEx(&cfg, scan_rows_kernel<1, 0, false, true>, o.rows));
                        else CU(ctx, cudaLaunchKernelEx(&cfg, scan_rows_kernel<2, 0, false, true>, o.rows));
                        q->timing.kernel_launches++;
                    }
                    break;
                }
                static const bool tma_env = getenv("COLQ_SCAN_ROWS_TMA") && getenv("COLQ_SCAN_ROWS_TMA")[0] == '1';
                const bool plain = o.np == 1 && o.ng == 0 && !o.eager && o.rows.in_bits == nullptr && o.rows.push.fk == nullptr &&
                                   o.rows.pred[0].promote == nullptr && o.rows.out_bits != nullptr && o.rows.n > 0;
                if (tma_env && plain) {
                    static bool attr_set = false;
                    const size_t smem = (size_t)SRT_STAGES * SRT_TILE_BYTES;
                    if (!attr_set) {
                        CU(ctx, cudaFuncSetAttribute(scan_rows_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        attr_set = true;
                    }
                    int occ = 0;
                    CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, scan_rows_tma_kernel, SRT_THREADS, smem));
                    const int64_t n_tiles = (o.rows.n + SR_BLOCK_ROWS - 1) / SR_BLOCK_ROWS;
                    ScanRowsTmaParams T{o.rows.n, o.rows.pred[0].col, o.rows.pred[0].lo, o.rows.pred[0].span, o.rows.out_bits, ctx->d_tile_counters};
                    scan_rows_tma_kernel<<<(int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * std::max(1, occ)), SRT_THREADS, smem, s>>>(T);
                } else {
                    launch_scan_rows(o, s);
                }
                q->timing.kernel_launches++;
            }
            break;
        case K_SCAN_CODES: {
            if (o.never) {  // an empty dictionary: no row matches
                if (o.codes.out_bits) CU(ctx, cudaMemsetAsync(o.codes.out_bits, 0, (size_t)bitmap_alloc_words(o.codes.n) * 4, s));
                break;
            }
            const int grid = (int)((o.codes.n + SC_BLOCK_ROWS - 1) / SC_BLOCK_ROWS);
            if (grid == 0) break;
            if (o.tail_publish) o.pmask.epoch = o.codes.pub.epoch = ++ctx->peer.mask_epoch;
            const size_t smem = (size_t)(PUSH_SMEM_WORDS + o.codes.mask_words) * 4;
            if (o.codes.mask_words > 0) scan_codes_kernel<true><<<grid, SR_THREADS, smem, s>>>(o.codes);
            else scan_codes_kernel<false><<<grid, SR_THREADS, smem, s>>>(o.codes);
            q->timing.kernel_launches++;
            break;
        }
        case K_SCAN_BOOL: {
            const int grid = (int)((o.boolp.n_alloc_words * 32 + SB_BLOCK_ROWS - 1) / SB_BLOCK_ROWS);
            scan_bool_kernel<<<grid, SB_THREADS, 0, s>>>(o.boolp);
            q->timing.kernel_launches++;
            break;
        }
        case K_SCAN_STR: {
            if (o.str.n_tiles == 0) break;
            // fixed family: EQ / NE / STARTS_WITH / ENDS_WITH with a needle of 1..16 bytes; generic family otherwise
            const int nl = o.str.needle_len, op = o.str.op;
            const bool fixed = (op == OP_EQ || op == OP_NE || op == OP_STARTS_WITH || op == OP_ENDS_WITH) && nl >= 1 && nl <= 16;
            const int mode = fixed ? -((nl + 3) / 4) - (op == OP_EQ ? 0 : 4) : op;
            const bool promote = o.str.promote_bytes != nullptr;
            void (*kern)(const ScanStrParams) = nullptr;
#define COLQ_STR_KERNEL(M) (promote ? scan_str_kernel<M, true> : scan_str_kernel<M, false>)
            switch (mode) {
                case -1: kern = COLQ_STR_KERNEL(-1); break;
                case -2: kern = COLQ_STR_KERNEL(-2); break;
                case -3: kern = COLQ_STR_KERNEL(-3); break;
                case -4: kern = COLQ_STR_KERNEL(-4); break;
                case -5: kern = COLQ_STR_KERNEL(-5); break;
                case -6: kern = COLQ_STR_KERNEL(-6); break;
                case -7: kern = COLQ_STR_KERNEL(-7); break;
                case -8: kern = COLQ_STR_KERNEL(-8); break;
                case OP_EQ: kern = COLQ_STR_KERNEL(OP_EQ); break;
                case OP_CONTAINS: kern = COLQ_STR_KERNEL(OP_CONTAINS); break;
                case OP_CMP_GT: kern = COLQ_STR_KERNEL(OP_CMP_GT); break;
                case OP_CMP_LT: kern = COLQ_STR_KERNEL(OP_CMP_LT); break;
                case OP_CMP_GE: kern = COLQ_STR_KERNEL(OP_CMP_GE); break;
                case OP_CMP_LE: kern = COLQ_STR_KERNEL(OP_CMP_LE); break;
                case OP_NE: kern = COLQ_STR_KERNEL(OP_NE); break;
                case OP_STARTS_WITH: kern = COLQ_STR_KERNEL(OP_STARTS_WITH); break;
                default: kern = COLQ_STR_KERNEL(OP_ENDS_WITH); break;
            }
#undef COLQ_STR_KERNEL
            if (o.grid == 0) {
                const std::pair<int, size_t> key(promote ? mode + 100 : mode, o.smem);
                auto it = ctx->str_occupancy.find(key);
                if (it == ctx->str_occupancy.end()) {
                    CU(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                    int occ = 0;
                    CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, ST_THREADS, o.smem));
                    if (occ < 1) return fail(ctx, COLQ_ERR_DEVICE, "scan_str_kernel does not fit on an SM (smem %zu)", o.smem);
                    it = ctx->str_occupancy.emplace(key, occ).first;
                }
                // (COLQ_STR_CTAS, experiment knob: resident CTAs per SM to use -- tiles are claimed dynamically, so any grid works)
                static const int ctas_env = getenv("COLQ_STR_CTAS") ? atoi(getenv("COLQ_STR_CTAS")) : 0;
                const int per_sm = ctas_env > 0 ? std::min(ctas_env, it->second) : it->second;
                o.grid = (int)std::min<int64_t>(o.str.n_tiles, (int64_t)ctx->sm_count * per_sm);
            }
            if (o.tail_publish) o.pmask.epoch = o.str.pub.epoch = ++ctx->peer.mask_epoch;
            if (o.str.early) {
                // pipelined behind the previous execution's root kernel (COLQ_OPT_PIPELINE): programmatic dependent launch
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(o.grid); cfg.blockDim = dim3(ST_THREADS); cfg.dynamicSmemBytes = o.smem; cfg.stream = s;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                CU(ctx, cudaLaunchKernelEx(&cfg, kern, o.str));
            } else {
                kern<<<o.grid, ST_THREADS, o.smem, s>>>(o.str);
            }
            q->timing.kernel_launches++;
            break;
        }
        case K_CSR_PULL: {
            int grid = (int)((o.csr.n + 255) / 256);
            if (o.csr.pm.n_words > 0) {
                o.csr.pm.epoch = q->ops[o.publish_op].pmask.epoch;
                grid = 1;  // the exchange must be collected even when the parent table is empty
            }
            if (grid == 0) break;
            csr_pull_kernel<<<grid, 256, 0, s>>>(o.csr);
            q->timing.kernel_launches++;
            break;
        }
        case K_PUSH_BITS:
            if (o.pushb.n_child == 0) break;
            push_bits_kernel<<<grid_for(o.pushb.n_child, 256, ctx->sm_count, 8), 256, 0, s>>>(o.pushb);
            q->timing.kernel_launches++;
            break;
        case K_AND:
            and_words_kernel<<<grid_for(o.n_words, 256, ctx->sm_count, 8), 256, 0, s>>>(o.dst, o.src, o.n_words);
            q->timing.kernel_launches++;
            break;
        case K_FILL:
            fill_ones_kernel<<<grid_for(o.n_alloc_words, 256, ctx->sm_count, 8), 256, 0, s>>>(o.dst, o.n_rows, o.n_alloc_words);
            q->timing.kernel_launches++;
            break;
        case K_ZERO:
            CU(ctx, cudaMemsetAsync(o.dst, 0, (size_t)o.n_alloc_words * 4, s));
            break;
        case K_ALLGATHER_OR:
            NC(ctx, ctx->nccl.AllGather(o.dst, o.gathered, (size_t)o.n_words, kNcclUint32, ctx->comm, (void*)s));
            q->timing.collectives++;
            or_ranks_kernel<<<grid_for(o.n_words, 256, ctx->sm_count, 8), 256, 0, s>>>(o.dst, o.gathered, o.n_words, ctx->n_ranks);
            q->timing.kernel_launches++;
            break;
        case K_POPC:
            popc_blocks_kernel<<<(int)o.n_blocks, CP_THREADS, 0, s>>>(o.src, o.n_words, o.block_counts);
            q->timing.kernel_launches++;
            break;
        case K_SCAN_COUNTS:
            scan_counts_kernel<<<1, 1024, 0, s>>>(o.block_counts, o.n_blocks, o.block_offsets, o.total);
            q->timing.kernel_launches++;
            break;
        case K_COMPACT:
            compact_kernel<<<(int)o.n_blocks, CP_THREADS, 0, s>>>(o.src, o.n_words, o.block_offsets, o.out_idx, o.capacity, o.row_base);
            q->timing.kernel_launches++;
            break;
        case K_COMPACT_LOOKBACK: {
            if (++q->lookback_epoch >= (1u << 30)) {  // epoch wrap: start over on cleared states
                CU(ctx, cudaMemsetAsync(q->lookback_buf.ptr, 0, q->lookback_buf.bytes, s));
                q->lookback_epoch = 1;
            }
            o.clook.epoch = q->lookback_epoch;
            switch (o.ng) {
                case 0: compact_lookback_kernel<0><<<o.grid, CP_THREADS, 0, s>>>(o.clook); break;
                case 1: compact_lookback_kernel<1><<<o.grid, CP_THREADS, 0, s>>>(o.clook); break;
                default: compact_lookback_kernel<2><<<o.grid, CP_THREADS, 0, s>>>(o.clook); break;
            }
            q->timing.kernel_launches++;
            break;
        }
        case K_ROOT_FINISH: {
            RootFusedParams& P = o.rfused;
            if (P.pre.pm.n_words > 0) P.pre.pm.epoch = q->ops[o.publish_op].pmask.epoch;
            if (P.pg.n_ranks > 0) {
                P.pg.epoch = ++ctx->peer.gather_epoch;
                q->lazy_pg = P.pg;
            }
            if (++q->rf_epoch == 0xffffffffu) {
                CU(ctx, cudaMemsetAsync(q->rf_state_buf.ptr, 0, q->rf_state_buf.bytes, s));
                q->rf_epoch = 1;
            }
            P.epoch = q->rf_epoch;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(o.grid); cfg.blockDim = dim3(RF_THREADS); cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CU(ctx, cudaLaunchKernelEx(&cfg, root_finish_kernel, P));
            q->timing.kernel_launches++;
            break;
        }
        case K_ROOT_FUSED: {
            RootFusedParams& P = o.rfused;
            if (P.pre.pm.n_words > 0) P.pre.pm.epoch = q->ops[o.publish_op].pmask.epoch;
            if (P.pg.n_ranks > 0) {
                P.pg.epoch = ++ctx->peer.gather_epoch;
                q->lazy_pg = P.pg;
            }
            if (++q->rf_epoch == 0xffffffffu) {  // epoch wrap: start over on cleared states
                CU(ctx, cudaMemsetAsync(q->rf_state_buf.ptr, 0, q->rf_state_buf.bytes, s));
                q->rf_epoch = 1;
            }
            P.epoch = q->rf_epoch;
            // programmatic dependent launch: the CTAs may move onto SMs that the kernel in front (the string scan) is
            // leaving; the kernel orders itself against that kernel with griddepcontrol.wait (pdl_wait)
            static const bool pdl = !(getenv("COLQ_PDL") && getenv("COLQ_PDL")[0] == '0');
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(o.grid); cfg.blockDim = dim3(RF_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
            cfg.attrs = at; cfg.numAttrs = 1;
            if (o.np == 1) CU(ctx, cudaLaunchKernelEx(&cfg, root_fused_kernel<1>, P));
            else CU(ctx, cudaLaunchKernelEx(&cfg, root_fused_kernel<2>, P));
            q->timing.kernel_launches++;
            break;
        }
        case K_COMPACT_FUSED: {
            if (o.gather) {
                o.cfused.pg.epoch = ++ctx->peer.gather_epoch;
                q->lazy_pg = o.cfused.pg;
            }
            void* args[] = {(void*)&o.cfused};
            CU(ctx, cudaLaunchCooperativeKernel(compact_fused_fn(o.ng, o.gather), dim3(o.grid), dim3(CP_THREADS), args, 0, s));
            q->timing.kernel_launches++;
            break;
        }
        case K_PEER_BITS_ALLGATHER:
            o.pbits.epoch = ++ctx->peer.mask_epoch;
            peer_bits_allgather_kernel<<<grid_for(std::max<int64_t>(o.pbits.n_words, 1), 256, ctx->sm_count, 2), 256, 0, s>>>(o.pbits);
            q->timing.kernel_launches++;
            break;
        case K_PEER_BITS_REDUCE:
            if (o.pbits.n_src > 1) o.pbits.epoch = ++ctx->peer.mask_epoch;
            peer_bits_reduce_kernel<<<grid_for(std::max<int64_t>(o.pbits.n_words, 1), 256, ctx->sm_count, 2), 256, 0, s>>>(o.pbits);
            q->timing.kernel_launches++;
            break;
        case K_PEER_MASK_PUBLISH:
            o.pmask.epoch = ++ctx->peer.mask_epoch;
            peer_mask_publish_kernel<<<1, 256, 0, s>>>(o.pmask);
            q->timing.kernel_launches++;
            break;
        case K_PEER_MASK_COLLECT:
            o.pmask.epoch = q->ops[o.publish_op].pmask.epoch;
            peer_mask_collect_kernel<<<1, 256, 0, s>>>(o.pmask);
            q->timing.kernel_launches++;
            break;
        case K_PEER_GATHER:
            o.pgather.epoch = ++ctx->peer.gather_epoch;
            peer_gather_send_kernel<<<ctx->n_ranks * o.pgather.blocks_per_peer, 256, 0, s>>>(o.pgather);
            peer_gather_recv_kernel<<<grid_for(o.capacity * ctx->n_ranks / 4, 256, ctx->sm_count, 2), 256, 0, s>>>(o.pgather);
            q->timing.kernel_launches += 2;
            break;
        case K_GATHER:
            NC(ctx, ctx->nccl.AllGather(q->idx_buf.ptr, q->gather_buf.ptr, (size_t)(o.capacity + GATHER_HEADER_WORDS), kNcclInt32,
                                        ctx->comm, (void*)s));
            q->timing.collectives++;
            unpack_gather_kernel<<<grid_for(o.capacity * ctx->n_ranks, 256, ctx->sm_count, 4), 256, 0, s>>>(
                (const int32_t*)q->gather_buf.ptr, ctx->n_ranks, o.capacity, (int32_t*)q->gout_buf.ptr, (u64*)q->ginfo_buf.ptr);
            q->timing.kernel_launches++;
            break;
    }
    CU(ctx, cudaGetLastError());
    return COLQ_OK;
}

colq_status ensure_idx_capacity(colq_query* q, int64_t want) {
    const size_t need = (size_t)(want + GATHER_HEADER_WORDS) * 4;
    if (q->idx_buf.bytes < need) {
        ST(dev_alloc(q->ctx, q->idx_buf, need));
        CU(q->ctx, cudaMemsetAsync(q->idx_buf.ptr, 0, GATHER_HEADER_WORDS * 4, q->ctx->stream));
    }
    q->d_total = (u64*)q->idx_buf.ptr;
    q->d_idx = (int32_t*)q->idx_buf.ptr + GATHER_HEADER_WORDS;
    // exactly what was asked for, never the slack of a recycled (up to 25 % larger) block: the capacity is an NCCL send
    // count and the stride of the gathered blocks, so it must be the same number on every rank
    q->idx_capacity = want;
    return COLQ_OK;
}

// verify + plan + enqueue. No host synchronisation.
colq_status run_pipeline(colq_query* q) {
    colq_ctx* ctx = q->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    ST(verify(q));
    q->pool.reset();
    q->ops.clear();
    q->timing = colq_timing{};
    q->local_count = -1;
    q->deferred.clear();
    q->own_begin = q->own_end = -1;
    q->pending_promotions.clear();
    q->lazy_oob = false;
    q->heap_cursor = 0;
    q->heap_used = false;
    q->promoted_bytes = 0;
    // the result block first: the planner hands its flags word to kernels that range-check lazily
    const Table& RT = ctx->tables[q->root_table];
    const int64_t n = RT.n_rows;
    q->gathered = ctx->n_ranks > 1 && RT.placement == COLQ_SHARDED;
    const bool local_group = ctx->n_ranks > 1 && ctx->comm == nullptr;
    const bool peer_gather = q->gathered && ctx->peer.ok && (q->opt_peer || local_group);
    if (q->want_idx_capacity <= 0) q->want_idx_capacity = (q->gathered && !peer_gather) ? (1 << 16) : (1 << 20);
    ST(ensure_idx_capacity(q, q->want_idx_capacity));
    Planner pl{q, ctx};
    NodeBits root;
    ST(pl.eval(0, Consume{}, &root));

    // ---- root fusion (COLQ_OPT_ROOT_FUSED): when the root node ends in a plain predicate scan (its to-one chains, if
    //      any, deferred), that scan, the chains, the compaction and the final gather become ONE persistent launch
    //      (root_fused_kernel); a tiny to-many hop that feeds a chain is folded in as well (below)
    bool fuse_root = false;
    if (q->opt_root_fused && q->opt_fused_compact == 1 && !q->ops.empty() && !root.all_ones) {
        const Op& l = q->ops.back();
        fuse_root = l.kind == K_SCAN_ROWS && l.node == 0 && !l.never && !l.dict_scan && !l.eager && l.ng == 0 && l.np >= 1 &&
                    l.rows.push.fk == nullptr && l.rows.out_bits != nullptr && l.rows.out_bits == root.bits;
        for (int p = 0; fuse_root && p < l.np; ++p) fuse_root = l.rows.pred[p].promote == nullptr;
    }

    // ---- peepholes over the op list (multi-GPU exchanges)
    // (1) overlap: the root's own predicate scans depend on no child, so they run between the first mask PUBLISH and
    //     its COLLECT -- the NVLink round trip and the wait for the slowest rank hide behind a bandwidth-bound scan
    //     (a fused root gets the same overlap inside its kernel: the COLLECT sits between its phases A and B)
    if (q->own_begin >= 0 && !fuse_root) {
        int pub = -1;
        for (int i = 0; i < q->own_begin; ++i)
            if (q->ops[i].kind == K_PEER_MASK_PUBLISH) { pub = i; break; }
        if (pub >= 0 && pub + 1 < q->own_begin) {
            const int shift = q->own_end - q->own_begin;
            std::rotate(q->ops.begin() + pub + 1, q->ops.begin() + q->own_begin, q->ops.begin() + q->own_end);
            for (Op& o : q->ops)  // ops between pub and own_begin moved right by `shift`; publish itself did not move
                if (o.publish_op > pub && o.publish_op < q->own_begin) o.publish_op += shift;
        }
    }
    // (2) a COLLECT directly followed by a single-block csr_pull over the collected mask becomes that kernel's prologue
    for (size_t i = 0; i + 1 < q->ops.size(); ++i) {
        Op& c = q->ops[i];
        Op& n = q->ops[i + 1];
        if (c.kind == K_PEER_MASK_COLLECT && n.kind == K_CSR_PULL && n.csr.n <= 256 && n.csr.child_bits == c.pmask.reach) {
            n.csr.pm = c.pmask;
            n.publish_op = c.publish_op;
            n.name = "peer_mask_collect+csr_pull";
            n.acct_bytes += c.acct_bytes;
            q->ops.erase(q->ops.begin() + i);
            for (Op& o : q->ops)
                if (o.publish_op > (int)i) o.publish_op -= 1;
        }
    }
    // (3) a PUBLISH directly behind the scan whose push epilogue produced that mask is done by the scan's last CTA
    for (size_t i = 0; q->opt_tail_publish && i + 1 < q->ops.size(); ++i) {
        Op& sc = q->ops[i];
        Op& pb = q->ops[i + 1];
        if (pb.kind != K_PEER_MASK_PUBLISH) continue;
        PeerMaskParams* slot = nullptr;
        if (sc.kind == K_SCAN_STR && sc.str.push.fk && sc.str.push.reach == pb.pmask.reach && sc.str.n_tiles > 0) {
            slot = &sc.str.pub; sc.str.pub_done = ctx->peer.d_done + MAX_RANKS + 8;
        } else if (sc.kind == K_SCAN_CODES && !sc.never && sc.codes.push.fk && sc.codes.push.reach == pb.pmask.reach && sc.codes.n > 0) {
            slot = &sc.codes.pub; sc.codes.pub_done = ctx->peer.d_done + MAX_RANKS + 8;
        }
        if (!slot) continue;
        *slot = pb.pmask;
        sc.tail_publish = true;
        sc.acct_bytes += pb.acct_bytes;
        q->ops.erase(q->ops.begin() + i + 1);
        for (Op& o : q->ops) {
            if (o.publish_op == (int)i + 1) o.publish_op = (int)i;
            else if (o.publish_op > (int)i + 1) o.publish_op -= 1;
        }
    }

    if (root.all_ones) {  // matchingBits.set(0, size) (E/ExecutionContext.java:83-87)
        u32* b;
        ST(pl.alloc_bitmap(n, &b));
        Op f{};
        f.kind = K_FILL; f.node = 0; f.dst = b; f.n_rows = n; f.n_alloc_words = bitmap_alloc_words(n); f.name = "fill_ones";
        f.acct_rows = n; f.acct_bytes = bitmap_words(n) * 4;
        q->ops.push_back(f);
        root.bits = b;
        q->xnodes[0].bits = b;
    }
    q->root_bits = root.bits;

    // ---- compaction of the root mask (M/InMemoryTable.java:121-131)
    const int64_t n_words = bitmap_words(n);
    const int64_t n_blocks = std::max<int64_t>(1, (n_words + CP_WORDS_PER_BLOCK - 1) / CP_WORDS_PER_BLOCK);
    void *bc, *bo;
    ST(pool_alloc(q, (size_t)n_blocks * 4, &bc));
    ST(pool_alloc(q, (size_t)n_blocks * 8, &bo));
    q->gather_is_peer = peer_gather;
    q->gather_block_cap = q->idx_capacity;
    q->gather_lazy = false;
    // peer-memory final gather written by the kernel that produces the indices (gather_store / gather_tail)
    auto fill_fused_gather = [&](PeerGatherParams& G) -> colq_status {
        if (!q->ginfo_buf.ptr) ST(dev_alloc(ctx, q->ginfo_buf, 64));
        const int64_t cap = std::min<int64_t>(q->idx_capacity, ctx->peer.slot_cap);
        if (q->gout_buf.bytes < (size_t)cap * 4 * ctx->n_ranks) ST(dev_alloc(ctx, q->gout_buf, (size_t)cap * 4 * ctx->n_ranks));
        G.count = q->d_total; G.idx = q->d_idx; G.idx_capacity = q->idx_capacity; G.slot_cap = cap;
        G.slot_bytes = ctx->peer.slot_bytes; G.n_ranks = ctx->n_ranks; G.rank = ctx->rank; G.peers = ctx->peer.d_peers;
        G.done = ctx->peer.d_done; G.blocks_per_peer = 0; G.status = ctx->peer.d_status;
        G.out = (int32_t*)q->gout_buf.ptr; G.info = (u64*)q->ginfo_buf.ptr;
        q->gather_is_peer = true;
        q->gather_block_cap = cap;
        q->gather_lazy = true;
        // The last block normally waits for every peer's flag before the launch ends: that keeps the ranks within one
        // execution of each other (slot parity).  A plan that already synchronises the ranks once per execution -- a mask
        // or bitmap exchange -- does not need a second all-rank wait per step: the flags are then awaited only when the
        // host fetches the result (peer_gather_recv_kernel).
        bool synced = false;
        for (const Op& o : q->ops)
            synced = synced || o.kind == K_PEER_MASK_PUBLISH || o.kind == K_PEER_MASK_COLLECT || o.kind == K_PEER_BITS_ALLGATHER ||
                     (o.kind == K_PEER_BITS_REDUCE && o.pbits.n_src > 1) || o.tail_publish || (o.kind == K_CSR_PULL && o.csr.pm.n_words > 0);
        G.wait = (synced && q->opt_lazy_gather_wait) ? 0 : 1;
        return COLQ_OK;
    };
    bool coop_gather = peer_gather && q->opt_fused_gather && q->opt_fused_compact == 1;  // the gather rides on the index writer
    // COLQ_OPT_ROOT_FUSED == 2: the bandwidth half stays the ordinary non-persistent scan_rows (LIST instantiation: it also
    // lists every chunk's survivors), the latency half is root_finish_kernel; tables too large for its per-CTA prefix
    // array use the single persistent kernel
    bool split_root = false;
    if (fuse_root && q->opt_root_fused == 2) {
        if (ctx->root_finish_grid == 0) {
            int occ = 0;
            CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)root_finish_kernel, RF_THREADS, 0));
            ctx->root_finish_grid = ctx->sm_count * std::max(1, occ);
        }
        const int64_t n_chunks = (n + SR_WARP_ROWS - 1) / SR_WARP_ROWS;
        split_root = n_chunks <= (int64_t)ctx->root_finish_grid * RF_MAX_UPC;
    }
    if (fuse_root && split_root) {
        Op& sc = q->ops.back();
        const int64_t n_chunks = (n + SR_WARP_ROWS - 1) / SR_WARP_ROWS;
        void *lists, *ucount;
        ST(pool_alloc(q, (size_t)std::max<int64_t>(n_chunks, 1) * (SR_WARP_ROWS / 32) * 4, &lists));
        ST(pool_alloc(q, (size_t)std::max<int64_t>(n_chunks, 1) * 4 + 16, &ucount));
        sc.list = true;
        sc.rows.lists = (u32*)lists; sc.rows.ucount = (u32*)ucount; sc.rows.list_cap = SR_WARP_ROWS / 32;
        sc.acct_bytes += n_chunks * 4;
        Op f{};
        f.kind = K_ROOT_FINISH; f.node = 0; f.np = sc.np; f.ng = (int)q->deferred.size();
        f.acct_rows = n; f.acct_bytes = n_chunks * 4;
        RootFusedParams& P = f.rfused;
        P.n = n;
        P.bits = root.bits;
        P.n_chunks = n_chunks;
        P.lists = (u32*)lists; P.ucount = (u32*)ucount; P.list_cap = SR_WARP_ROWS / 32;
        f.grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_chunks + 31) / 32, ctx->root_finish_grid));
        P.units_per_cta = std::max<int64_t>(1, (n_chunks + f.grid - 1) / f.grid);
        if (q->rf_state_ctas < f.grid) {
            ST(dev_alloc(ctx, q->rf_state_buf, 64 + (size_t)f.grid * 8));
            CU(ctx, cudaMemsetAsync(q->rf_state_buf.ptr, 0, q->rf_state_buf.bytes, ctx->stream));
            q->rf_state_ctas = f.grid;
            q->rf_epoch = 0;
        }
        P.counters = (u32*)q->rf_state_buf.ptr;
        P.cta_state = (u64*)((char*)q->rf_state_buf.ptr + 64);
        P.ng = f.ng;
        for (int g = 0; g < f.ng; ++g) P.gather[g] = q->deferred[g];
        P.total = q->d_total; P.out_idx = q->d_idx; P.capacity = q->idx_capacity; P.row_base = RT.row_base;
        for (int g = 0; g < f.ng && P.pre.n == 0; ++g) {
            if (!P.gather[g].bits) continue;
            for (size_t k = 0; k < q->ops.size(); ++k) {
                const Op& c = q->ops[k];
                if (c.kind != K_CSR_PULL || c.csr.out_bits != P.gather[g].bits || c.csr.push.fk != nullptr) continue;
                if (c.csr.n <= 0 || c.csr.n > RF_PRE_ROWS || c.csr.nnz > RF_PRE_EDGES || c.csr.n_child > PUSH_SMEM_BITS) continue;
                P.pre = c.csr;
                f.publish_op = c.publish_op;
                f.acct_bytes += c.acct_bytes;
                q->ops.erase(q->ops.begin() + k);
                for (Op& o : q->ops)
                    if (o.publish_op > (int)k) o.publish_op -= 1;
                if (f.publish_op > (int)k) f.publish_op -= 1;
                break;
            }
        }
        if (P.pre.n > 0)
            for (int g = 0; g < f.ng; ++g)
                if (P.gather[g].bits == P.pre.out_bits) P.pre_mask |= 1u << g;
        if (coop_gather) ST(fill_fused_gather(P.pg));
#ifdef COLQ_RF_DEBUG
        {
            void* dbg;
            ST(pool_alloc(q, (size_t)f.grid * 64, &dbg));
            P.dbg = (u64*)dbg;
            q->rf_dbg = P.dbg;
            q->rf_dbg_ctas = f.grid;
        }
#endif
        f.name = "root_finish";
        q->ops.push_back(f);
    } else if (fuse_root) {
        // ---- the root's scan + chains + (tiny to-many hop) + compaction + gather as one persistent launch
        Op sc = q->ops.back();
        q->ops.pop_back();
        Op f{};
        f.kind = K_ROOT_FUSED; f.node = 0; f.np = sc.np; f.ng = (int)q->deferred.size();
        f.acct_rows = n; f.acct_bytes = sc.acct_bytes;
        RootFusedParams& P = f.rfused;
        P.n = n;
        for (int p = 0; p < sc.np; ++p) P.pred[p] = sc.rows.pred[p];
        P.in_bits = sc.rows.in_bits;
        P.bits = root.bits;
        P.n_chunks = (n + SR_WARP_ROWS - 1) / SR_WARP_ROWS;
        if (ctx->root_fused_grid[sc.np] == 0) {
            int occ = 0;
            const void* fn = sc.np == 1 ? (const void*)root_fused_kernel<1> : (const void*)root_fused_kernel<2>;
            CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, RF_THREADS, 0));
            ctx->root_fused_grid[sc.np] = ctx->sm_count * std::max(1, occ);
        }
        static const int rf_grid_env = getenv("COLQ_RF_GRID") ? atoi(getenv("COLQ_RF_GRID")) : 0;  // experiment knob: CTAs per SM
        const int64_t max_grid = rf_grid_env > 0 ? (int64_t)ctx->sm_count * rf_grid_env : ctx->root_fused_grid[sc.np];
        f.grid = (int)std::max<int64_t>(1, std::min<int64_t>((P.n_chunks + RF_WARPS - 1) / RF_WARPS, max_grid));
        P.chunks_per_warp = std::max<int64_t>(1, (P.n_chunks + (int64_t)f.grid * RF_WARPS - 1) / ((int64_t)f.grid * RF_WARPS));
        // candidate lists: room for ~3 % of a warp's rows, at least 128; denser CTAs take the dense path
        P.list_cap = (int)std::max<int64_t>(128, P.chunks_per_warp * (SR_WARP_ROWS / 32));
        void* lists;
        ST(pool_alloc(q, (size_t)f.grid * RF_WARPS * P.list_cap * 4 * (1 + f.ng), &lists));
        P.lists = (u32*)lists;
        if (q->rf_state_ctas < f.grid) {
            ST(dev_alloc(ctx, q->rf_state_buf, 64 + (size_t)f.grid * 8));
            CU(ctx, cudaMemsetAsync(q->rf_state_buf.ptr, 0, q->rf_state_buf.bytes, ctx->stream));
            q->rf_state_ctas = f.grid;
            q->rf_epoch = 0;
        }
        P.counters = (u32*)q->rf_state_buf.ptr;
        P.cta_state = (u64*)((char*)q->rf_state_buf.ptr + 64);
        P.ng = f.ng;
        for (int g = 0; g < f.ng; ++g) P.gather[g] = q->deferred[g];
        P.total = q->d_total; P.out_idx = q->d_idx; P.capacity = q->idx_capacity; P.row_base = RT.row_base;
        // fold a tiny to-many hop whose output only feeds a deferred chain: every CTA redoes it in shared memory
        for (int g = 0; g < f.ng && P.pre.n == 0; ++g) {
            if (!P.gather[g].bits) continue;
            for (size_t k = 0; k < q->ops.size(); ++k) {
                const Op& c = q->ops[k];
                if (c.kind != K_CSR_PULL || c.csr.out_bits != P.gather[g].bits || c.csr.push.fk != nullptr) continue;
                if (c.csr.n <= 0 || c.csr.n > RF_PRE_ROWS || c.csr.nnz > RF_PRE_EDGES || c.csr.n_child > PUSH_SMEM_BITS) continue;
                P.pre = c.csr;
                f.publish_op = c.publish_op;
                f.acct_bytes += c.acct_bytes;
                q->ops.erase(q->ops.begin() + k);
                for (Op& o : q->ops)
                    if (o.publish_op > (int)k) o.publish_op -= 1;
                if (f.publish_op > (int)k) f.publish_op -= 1;
                break;
            }
        }
        if (P.pre.n > 0)
            for (int g = 0; g < f.ng; ++g)
                if (P.gather[g].bits == P.pre.out_bits) P.pre_mask |= 1u << g;
        if (coop_gather) ST(fill_fused_gather(P.pg));
#ifdef COLQ_RF_DEBUG
        {
            void* dbg;
            ST(pool_alloc(q, (size_t)f.grid * 64, &dbg));
            P.dbg = (u64*)dbg;
            q->rf_dbg = P.dbg;
            q->rf_dbg_ctas = f.grid;
        }
#endif
        f.name = "root_fused";
        q->ops.push_back(f);
    } else if (q->opt_fused_compact >= 2 && !coop_gather) {
        // single pass with decoupled look-back: one CTA per tile, ordinary launch
        const int ng = (int)q->deferred.size();
        const int64_t n_tiles = std::max<int64_t>(1, (n_words + CF_WORDS_PER_TILE - 1) / CF_WORDS_PER_TILE);
        if (q->lookback_tiles < n_tiles) {
            ST(dev_alloc(ctx, q->lookback_buf, 64 + (size_t)n_tiles * 8));
            CU(ctx, cudaMemsetAsync(q->lookback_buf.ptr, 0, q->lookback_buf.bytes, ctx->stream));
            q->lookback_tiles = n_tiles;
            q->lookback_epoch = 0;
        }
        Op f{};
        f.kind = K_COMPACT_LOOKBACK; f.node = 0; f.acct_rows = n; f.acct_bytes = n_words * 4;
        f.name = ng ? "compact_lookback+chains" : "compact_lookback";
        f.ng = ng;
        CompactLookbackParams& P = f.clook;
        P.bits = root.bits; P.n_words = n_words; P.n_tiles = n_tiles;
        P.counters = (u32*)q->lookback_buf.ptr; P.tile_state = (u64*)((char*)q->lookback_buf.ptr + 64);
        P.total = q->d_total; P.out_idx = q->d_idx; P.capacity = q->idx_capacity;
        P.row_base = RT.row_base; P.n_rows = n;
        for (int g = 0; g < ng; ++g) P.gather[g] = q->deferred[g];
        f.grid = (int)n_tiles;
        q->ops.push_back(f);
    } else if (q->opt_fused_compact) {
        // one cooperative launch: per-tile popcount, grid barrier, ordered write (+ the peer-memory final gather:
        // COLQ_OPT_FUSED_GATHER, default on)
        const int ng = (int)q->deferred.size();
        const bool fuse_gather = coop_gather;
        const int variant = ng * 2 + (fuse_gather ? 1 : 0);
        if (ctx->compact_grid[variant] == 0) {
            int occ = 0;
            CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, compact_fused_fn(ng, fuse_gather), CP_THREADS, 0));
            ctx->compact_grid[variant] = ctx->sm_count * std::max(1, std::min(occ, 8));
        }
        if (!q->barrier_buf.ptr) {
            ST(dev_alloc(ctx, q->barrier_buf, 64));
            CU(ctx, cudaMemsetAsync(q->barrier_buf.ptr, 0, 64, ctx->stream));
        }
        Op f{};
        f.kind = K_COMPACT_FUSED; f.node = 0; f.name = "compact_fused"; f.acct_rows = n; f.acct_bytes = n_words * 8;
        CompactFusedParams& P = f.cfused;
        const int64_t n_tiles = std::max<int64_t>(1, (n_words + CF_WORDS_PER_TILE - 1) / CF_WORDS_PER_TILE);
        P.bits = root.bits; P.n_words = n_words; P.n_tiles = n_tiles; P.tile_counts = (u32*)bc;
        P.barrier = (u32*)q->barrier_buf.ptr; P.total = q->d_total; P.out_idx = q->d_idx; P.capacity = q->idx_capacity;
        P.row_base = RT.row_base;  // global row index = shard-local index + the shard's base
        P.n_rows = n;
        f.ng = ng;
        for (int g = 0; g < ng; ++g) P.gather[g] = q->deferred[g];
        f.name = ng ? (fuse_gather ? "compact_fused+chains+gather" : "compact_fused+chains") : (fuse_gather ? "compact_fused+gather" : "compact_fused");
        f.grid = (int)std::min<int64_t>(n_tiles, ctx->compact_grid[variant]);
        if (fuse_gather) {
            ST(fill_fused_gather(P.pg));
            f.gather = true;
            f.capacity = q->gather_block_cap;
        }
        q->ops.push_back(f);
    } else {
        coop_gather = false;
        Op p{};
        p.kind = K_POPC; p.node = 0; p.src = root.bits; p.n_words = n_words; p.n_blocks = n_blocks; p.block_counts = (u32*)bc;
        p.name = "popc_blocks"; p.acct_rows = n; p.acct_bytes = n_words * 4;
        q->ops.push_back(p);
        Op sc{};
        sc.kind = K_SCAN_COUNTS; sc.node = 0; sc.block_counts = (u32*)bc; sc.n_blocks = n_blocks; sc.block_offsets = (u64*)bo;
        sc.total = q->d_total; sc.name = "scan_counts"; sc.acct_bytes = n_blocks * 12;
        q->ops.push_back(sc);
        Op c{};
        c.kind = K_COMPACT; c.node = 0; c.src = root.bits; c.n_words = n_words; c.n_blocks = n_blocks; c.block_offsets = (u64*)bo;
        c.out_idx = q->d_idx; c.capacity = q->idx_capacity;
        c.row_base = RT.row_base;
        c.name = "compact"; c.acct_rows = n; c.acct_bytes = n_words * 4;
        q->ops.push_back(c);
    }
    if (q->gathered && !coop_gather) {
        // final gather of matched indices as launches of its own, entirely on the device
        if (!q->ginfo_buf.ptr) ST(dev_alloc(ctx, q->ginfo_buf, 64));
        if (peer_gather) {
            // own kernels over NVLink peer memory: every rank stores its indices into every peer's mailbox slot
            const int64_t cap = std::min<int64_t>(q->idx_capacity, ctx->peer.slot_cap);
            if (q->gout_buf.bytes < (size_t)cap * 4 * ctx->n_ranks) ST(dev_alloc(ctx, q->gout_buf, (size_t)cap * 4 * ctx->n_ranks));
            q->gather_block_cap = cap;
            Op g{};
            g.kind = K_PEER_GATHER; g.node = 0; g.name = "peer_gather_indices"; g.capacity = cap;
            PeerGatherParams& P = g.pgather;
            P.count = q->d_total; P.idx = q->d_idx; P.idx_capacity = q->idx_capacity; P.slot_cap = cap;
            P.slot_bytes = ctx->peer.slot_bytes; P.n_ranks = ctx->n_ranks; P.rank = ctx->rank; P.peers = ctx->peer.d_peers;
            P.done = ctx->peer.d_done; P.blocks_per_peer = std::max(1, std::min(32, 2 * ctx->sm_count / ctx->n_ranks)); P.status = ctx->peer.d_status;
            P.out = (int32_t*)q->gout_buf.ptr; P.info = (u64*)q->ginfo_buf.ptr;
            q->ops.push_back(g);
        } else {
            // NCCL: one all-gather of fixed-size result blocks, then a kernel that concatenates their valid prefixes
            const int64_t cap = q->idx_capacity;
            const size_t block_bytes = (size_t)(cap + GATHER_HEADER_WORDS) * 4;
            if (q->gather_buf.bytes < block_bytes * ctx->n_ranks) ST(dev_alloc(ctx, q->gather_buf, block_bytes * ctx->n_ranks));
            if (q->gout_buf.bytes < (size_t)cap * 4 * ctx->n_ranks) ST(dev_alloc(ctx, q->gout_buf, (size_t)cap * 4 * ctx->n_ranks));
            Op g{};
            g.kind = K_GATHER; g.node = 0; g.capacity = cap; g.name = "allgather_indices";
            g.acct_bytes = (int64_t)block_bytes * ctx->n_ranks;
            q->ops.push_back(g);
        }
    }

    // ---- pipelining of back-to-back executions (COLQ_OPT_PIPELINE): exactly [memset R][scan_str pushing into R]
    //      [root_fused whose folded hop reads R (or the exchange of R)] with R small.  The root kernel's last CTA re-zeroes
    //      R, so the memset is dropped from the second execution on; and when the previous thing on the stream was this
    //      query's root kernel, the string scan is launched as its programmatic dependent (ScanStrParams::early).
    static const bool pdl_env = !(getenv("COLQ_PDL") && getenv("COLQ_PDL")[0] == '0');
    bool pipelined = false, early = false;
    if (q->opt_pipeline && q->opt_profile != 1 && q->ops.size() == 3 && q->ops[0].kind == K_ZERO && q->ops[1].kind == K_SCAN_STR &&
        q->ops[2].kind == K_ROOT_FUSED) {
        const Op& z = q->ops[0];
        Op& sc = q->ops[1];
        Op& rf = q->ops[2];
        u32* R = z.dst;
        const bool exchanged = rf.rfused.pre.pm.n_words > 0;
        const bool reads_r = rf.rfused.pre.n > 0 && (exchanged ? (rf.rfused.pre.pm.reach == R && sc.tail_publish && sc.str.pub.reach == R)
                                                                : rf.rfused.pre.child_bits == R);
        if (R != nullptr && reads_r && z.n_alloc_words <= 4096 && sc.str.push.fk != nullptr && sc.str.push.reach == R &&
            sc.str.push.n_parent <= PUSH_SMEM_BITS && sc.str.in_bits == nullptr && sc.str.out_bits == nullptr && sc.str.n_tiles > 0 &&
            rf.rfused.in_bits != R && rf.rfused.pre.in_bits != R) {
            pipelined = true;
            rf.rfused.clean = R;
            rf.rfused.clean_words = (int)z.n_alloc_words;
            if (exchanged) rf.rfused.pre.pm.reach = nullptr;  // the reduced mask is not written back into R
            for (XNode& x : q->xnodes)
                if (x.bits == R) { x.bits = nullptr; x.fused = true; }  // the mask does not outlive the execution (cardinality -1)
            const bool was_clean = q->clean_ready == R && z.n_alloc_words <= q->clean_words;
            // the scan may start early only if it has no side effect in global memory before its wait: no promotion, no
            // lazy range-check flag, and nothing between it and the root kernel on the stream (events, memsets)
            early = was_clean && pdl_env && ctx->chain_query == q && q->opt_profile == 0 && !q->lazy_oob && sc.str.push.oob == nullptr &&
                    sc.str.promote_bytes == nullptr && sc.str.promote_offsets == nullptr;
            sc.str.early = early ? 1u : 0u;
            if (was_clean) {
                q->ops.erase(q->ops.begin());
                for (Op& o : q->ops)
                    if (o.publish_op > 0) o.publish_op -= 1;
            }
        }
    }
    q->clean_ready = nullptr;   // set again below, once everything is enqueued
    ctx->chain_query = nullptr;

    // ---- enqueue
    cudaStream_t s = ctx->stream;
    if (!q->ev_start) {
        CU(ctx, cudaEventCreate(&q->ev_start));
        CU(ctx, cudaEventCreate(&q->ev_stop));
    }
    const bool prof = q->opt_profile == 1;
    int hot = -1;
    if (q->opt_profile == 2) {
        int64_t best = -1;
        for (size_t i = 0; i < q->ops.size(); ++i)
            if (q->ops[i].acct_bytes > best && q->ops[i].kind != K_ZERO) { best = q->ops[i].acct_bytes; hot = (int)i; }
        if (q->hot_used == q->hot_ring.size()) {
            if (q->hot_ring.size() >= 4096) hot = -1;  // ring full: stop sampling until colq_profile_hot drains it
            else {
                cudaEvent_t a, b;
                CU(ctx, cudaEventCreate(&a));
                CU(ctx, cudaEventCreate(&b));
                q->hot_ring.emplace_back(a, b);
            }
        }
    }
    if (prof) {
        while (q->stage_ev.size() < q->ops.size() + 1) {
            cudaEvent_t e;
            CU(ctx, cudaEventCreate(&e));
            q->stage_ev.push_back(e);
        }
    }
    // (experiment knob COLQ_PIPELINE_EVENTS=1: record the per-execution events even between pipelined executions)
    static const bool ev_env = getenv("COLQ_PIPELINE_EVENTS") && getenv("COLQ_PIPELINE_EVENTS")[0] == '1';
    q->timed = !early || ev_env;
    if (q->timed) CU(ctx, cudaEventRecord(q->ev_start, s));
    if (q->lazy_oob) CU(ctx, cudaMemsetAsync((u32*)q->idx_buf.ptr + RESULT_FLAGS_WORD, 0, 4, s));
    if (prof) CU(ctx, cudaEventRecord(q->stage_ev[0], s));
    for (size_t i = 0; i < q->ops.size(); ++i) {
        if ((int)i == hot) CU(ctx, cudaEventRecord(q->hot_ring[q->hot_used].first, s));
        ST(launch_op(q, q->ops[i], s));
        if (prof) CU(ctx, cudaEventRecord(q->stage_ev[i + 1], s));
        if ((int)i == hot) {
            CU(ctx, cudaEventRecord(q->hot_ring[q->hot_used++].second, s));
            const Op& o = q->ops[i];
            colq_stage& st = q->hot_stage;
            stage_name(o, st.name, sizeof st.name);
            st.rows = o.acct_rows;
            st.bytes = o.acct_bytes;
        }
    }
    // (an event between two executions would keep the next one from starting early: when this one may be followed by a
    //  pipelined execution, the stop event is recorded only if this execution itself was timed from its start)
    if (q->timed && (ev_env || !(pipelined && pdl_env && q->opt_profile == 0))) CU(ctx, cudaEventRecord(q->ev_stop, s));
    else q->timed = false;
    if (pipelined) {
        q->clean_ready = q->ops.back().rfused.clean;
        q->clean_words = q->ops.back().rfused.clean_words;
        if (q->opt_profile == 0) ctx->chain_query = q;
    }
    // first-touch promotion: the scans enqueued above fill the HBM copies; everything enqueued later on this stream
    // (the next query's plan included) reads those instead of the pinned host memory
    for (Column* c : q->pending_promotions) {
        if (!c->host_resident || !c->promoted.ptr) continue;
        c->data = std::move(c->promoted);
        if (c->kind == COL_STR) {
            c->offsets = std::move(c->promoted_offsets);
            c->bytes_capacity = (int64_t)(c->data.bytes & ~(size_t)15);
        }
        c->host_resident = false;
    }
    q->pending_promotions.clear();
    q->executed = true;
    return COLQ_OK;
}

// internal: a rank of a LOCAL communicator needs a larger result block; its peers' kernels wait for it across GPUs, so the
// re-run has to be enqueued on every rank before anybody fetches again (colq_fetch_group does that)
constexpr colq_status RERUN_GROUP = (colq_status)100;

// count D2H, (multi-GPU) final gather to rank 0, result copies. Synchronises the stream.
colq_status fetch_results_impl(colq_query* q, uint64_t* out_bitmask, int64_t bitmask_cap, int32_t* out_idx, int64_t idx_cap,
                               int64_t* out_count, colq_timing* out_timing);

colq_status fetch_results(colq_query* q, uint64_t* out_bitmask, int64_t bitmask_cap, int32_t* out_idx, int64_t idx_cap,
                          int64_t* out_count, colq_timing* out_timing) {
    const colq_status st = fetch_results_impl(q, out_bitmask, bitmask_cap, out_idx, idx_cap, out_count, out_timing);
    if (st != COLQ_OK) {  // a failed execution may have left the push target dirty: the next one starts from a memset
        q->clean_ready = nullptr;
        q->ctx->chain_query = nullptr;
    }
    return st;
}

colq_status fetch_results_impl(colq_query* q, uint64_t* out_bitmask, int64_t bitmask_cap, int32_t* out_idx, int64_t idx_cap,
                               int64_t* out_count, colq_timing* out_timing) {
    colq_ctx* ctx = q->ctx;
    if (!q->executed) return fail(ctx, COLQ_THROW_ILLEGAL_STATE, "colq_fetch before colq_execute_async");
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const Table& RT = ctx->tables[q->root_table];
    u64 header[2] = {0, 0};  // [count | flags, pad]
    u64 ginfo[4] = {0, 0, 0, 0};
    const bool gather = q->gathered;
    // small results (the 29k-row / 51-row queries of BASELINE configs 0 and 2 are launch-latency bound): the header and
    // the first FETCH_SPEC indices are contiguous in the result block, so one copy and one synchronisation fetch both
    constexpr int64_t FETCH_SPEC = 2048;
    const int64_t n_spec = (!gather && out_idx) ? std::min<int64_t>(FETCH_SPEC, q->idx_capacity) : 0;
    if (n_spec > 0 && !ctx->h_stage) CU(ctx, cudaHostAlloc(&ctx->h_stage, 16 + FETCH_SPEC * 4, cudaHostAllocDefault));
    if (n_spec > 0) {
        CU(ctx, cudaMemcpyAsync(ctx->h_stage, q->d_total, 16 + (size_t)n_spec * 4, cudaMemcpyDeviceToHost, s));
        CU(ctx, cudaStreamSynchronize(s));
        memcpy(header, ctx->h_stage, 16);
    } else {
        if (gather && q->gather_lazy) {
            // the ranks' indices sit in this rank's mailbox slots: wait for every rank's flag (if the launch itself did
            // not), summarise the counts and concatenate the valid prefixes in rank order
            peer_gather_recv_kernel<<<grid_for(q->gather_block_cap * ctx->n_ranks / 4 + 1, 256, ctx->sm_count, 2), 256, 0, s>>>(q->lazy_pg);
            CU(ctx, cudaGetLastError());
            q->timing.kernel_launches++;
        }
        CU(ctx, cudaMemcpyAsync(header, q->d_total, 16, cudaMemcpyDeviceToHost, s));
        if (gather) CU(ctx, cudaMemcpyAsync(ginfo, q->ginfo_buf.ptr, 32, cudaMemcpyDeviceToHost, s));
        CU(ctx, cudaStreamSynchronize(s));
    }
    q->timing.d2h_bytes += gather ? 48 : 16;
    const u64 local = header[0];
    if (q->lazy_oob && (header[1] & 1u))  // M/InMemoryTable.java:70-71 would have thrown at associateTo
        return fail(ctx, COLQ_THROW_NULL, "association target outside the associated table (found while walking a host-resident to-one column)");
    float ms = -1.f;
    if (q->timed) CU(ctx, cudaEventElapsedTime(&ms, q->ev_start, q->ev_stop));
    q->timing.gpu_ms = ms;

    if (ctx->peer.ok) {
        u32 status = 0;
        CU(ctx, cudaMemcpyAsync(&status, ctx->peer.d_status, 4, cudaMemcpyDeviceToHost, s));
        CU(ctx, cudaStreamSynchronize(s));
        if (status != 0) {
            cudaMemsetAsync(ctx->peer.d_status, 0, 4, s);  // report once; later fetches on this context start clean
            return fail(ctx, COLQ_ERR_DEVICE, "peer-memory exchange timed out waiting for another rank");
        }
    }
    int64_t count = (int64_t)local;
    const int32_t* src_idx = q->d_idx;
    if (gather) {
        const bool peer_gather = q->gather_is_peer;
        const int64_t block_cap = q->gather_block_cap;
        if ((int64_t)ginfo[1] > block_cap) {
            // some rank found more rows than a result block holds: every rank sees the same gathered counts, so all
            // of them grow the block (or leave the fixed-size mailbox path) and run the query again
            q->want_idx_capacity = (int64_t)ginfo[1] + (int64_t)ginfo[1] / 4 + 1024;
            if (ctx->comm == nullptr && q->want_idx_capacity <= ctx->peer.slot_cap) return RERUN_GROUP;  // all ranks re-run together (colq_fetch_group)
            if (peer_gather && q->want_idx_capacity > ctx->peer.slot_cap) {
                if (ctx->comm == nullptr)
                    return fail(ctx, COLQ_ERR_CAPACITY, "a rank matched %llu rows but a mailbox slot holds %lld; a local communicator has no NCCL path: raise COLQ_PEER_SLOT_CAP",
                                (unsigned long long)ginfo[1], (long long)ctx->peer.slot_cap);
                q->opt_peer = 0;
            }
            ST(run_pipeline(q));
            return fetch_results(q, out_bitmask, bitmask_cap, out_idx, idx_cap, out_count, out_timing);
        }
        count = (int64_t)ginfo[2];
        src_idx = (const int32_t*)q->gout_buf.ptr;

    } else if ((int64_t)local > q->idx_capacity) {
        // index buffer was too small: grow it and run again (happens once per query, the capacity sticks)
        q->want_idx_capacity = (int64_t)local + (int64_t)local / 8 + 1024;
        if (ctx->n_ranks > 1 && ctx->comm == nullptr) return RERUN_GROUP;
        ST(run_pipeline(q));
        return fetch_results(q, out_bitmask, bitmask_cap, out_idx, idx_cap, out_count, out_timing);
    }

    q->local_count = (int64_t)local;
    if (out_count) *out_count = count;
    colq_status rc = COLQ_OK;
    if (out_bitmask) {
        int64_t words64 = (RT.n_rows + 63) / 64;
        if (bitmask_cap < words64) rc = fail(ctx, COLQ_ERR_CAPACITY, "bitmask capacity %lld < %lld words", (long long)bitmask_cap, (long long)words64);
        else {
            CU(ctx, cudaMemcpyAsync(out_bitmask, q->root_bits, (size_t)words64 * 8, cudaMemcpyDeviceToHost, s));
            q->timing.d2h_bytes += words64 * 8;
        }
    }
    bool pending = out_bitmask != nullptr && rc == COLQ_OK;
    if (out_idx) {
        if (idx_cap < count) rc = fail(ctx, COLQ_ERR_CAPACITY, "index capacity %lld < %lld matches", (long long)idx_cap, (long long)count);
        else if (count > 0 && count <= n_spec) {
            memcpy(out_idx, (const char*)ctx->h_stage + 16, (size_t)count * 4);  // already here
            q->timing.d2h_bytes += count * 4;
        } else if (count > 0) {
            CU(ctx, cudaMemcpyAsync(out_idx, src_idx, (size_t)count * 4, cudaMemcpyDeviceToHost, s));
            q->timing.d2h_bytes += count * 4;
            pending = true;
        }
    }
    if (pending) CU(ctx, cudaStreamSynchronize(s));

    // per-stage profile
    q->stages.clear();
    for (size_t i = 0; i < q->ops.size(); ++i) {
        const Op& o = q->ops[i];
        if (o.kind == K_ZERO) continue;
        colq_stage st{};
        stage_name(o, st.name, sizeof st.name);
        st.rows = o.acct_rows;
        st.bytes = o.acct_bytes;
        st.ms = -1.0;
        if (q->opt_profile && q->stage_ev.size() > i + 1) {
            float t = 0;
            if (cudaEventElapsedTime(&t, q->stage_ev[i], q->stage_ev[i + 1]) == cudaSuccess) st.ms = t;
        }
        q->stages.push_back(st);
    }
    if (out_timing) *out_timing = q->timing;
    return rc;
}

colq_status upload(colq_ctx* ctx, DevBuf& b, const void* host, size_t bytes, size_t alloc_bytes) {
    ST(dev_alloc(ctx, b, alloc_bytes));
    if (alloc_bytes > bytes) CU(ctx, cudaMemsetAsync((char*)b.ptr + bytes, 0, alloc_bytes - bytes, ctx->stream));
    if (bytes) CU(ctx, cudaMemcpyAsync(b.ptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return COLQ_OK;
}

colq_status check_fk_range(colq_ctx* ctx, const int32_t* d_fk, int64_t n, int64_t n_target) {
    if (n == 0) return COLQ_OK;
    DevBuf mm;
    ST(dev_alloc(ctx, mm, 8));
    int32_t init[2] = {INT32_MAX, INT32_MIN};
    CU(ctx, cudaMemcpyAsync(mm.ptr, init, 8, cudaMemcpyHostToDevice, ctx->stream));
    fk_minmax_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(d_fk, n, (int32_t*)mm.ptr, (int32_t*)mm.ptr + 1);
    CU(ctx, cudaGetLastError());
    int32_t got[2];
    CU(ctx, cudaMemcpyAsync(got, mm.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (got[0] < -1 || got[1] >= n_target)  // M/InMemoryTable.java:70-71: yIndexToXAssociations.get(yIndex) is null -> NPE
        return fail(ctx, COLQ_THROW_NULL, "association target outside the associated table (min %d, max %d, size %lld)", got[0],
                    got[1], (long long)n_target);
    return COLQ_OK;
}

// degree / order / range statistics of a CSR that already sits in device memory (one pass, no host loop)
colq_status assoc_stats(colq_ctx* ctx, const int64_t* d_offsets, const int32_t* d_targets, int64_t n, int64_t nnz, AssocStats* out) {
    DevBuf st;
    ST(dev_alloc(ctx, st, sizeof(AssocStats)));
    AssocStats init{0, 0, INT32_MAX, INT32_MIN};
    CU(ctx, cudaMemcpyAsync(st.ptr, &init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    assoc_stats_kernel<<<grid_for(std::max<int64_t>(std::max(n, nnz), 1), 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(d_offsets, d_targets, n, nnz, (AssocStats*)st.ptr);
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out, st.ptr, sizeof init, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return COLQ_OK;
}

colq_status check_csr(colq_ctx* ctx, const AssocStats& a, int64_t nnz, int64_t n_target) {
    if (a.bad_offsets) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "CSR offsets must be non-decreasing");
    if (nnz > 0 && (a.min_target < 0 || a.max_target >= n_target))  // M/InMemoryTable.java:70-71
        return fail(ctx, COLQ_THROW_NULL, "association target outside the associated table (min %d, max %d, size %lld)", a.min_target, a.max_target,
                    (long long)n_target);
    return COLQ_OK;
}

colq_status link_assoc(colq_ctx* ctx, colq_table x, int xo, colq_table y, int yo, bool is_fk, Column** fwd_out) {
    Table* X = get_table(ctx, x);
    Table* Y = get_table(ctx, y);
    if (!X || !Y) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown table handle");
    if (x == y && xo == yo) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "forward and reverse column need distinct ordinals");
    Column *f, *r;
    ST(slot_for(ctx, x, xo, X->n_rows, &f));
    f->kind = COL_ASSOC;  // claim before asking for the second slot (x may equal y)
    colq_status st = slot_for(ctx, y, yo, Y->n_rows, &r);
    f = &ctx->tables[x].cols[xo];  // slot_for may have grown the vector
    if (st != COLQ_OK) { f->kind = COL_UNSET; return st; }
    f->n = X->n_rows; f->forward = true; f->is_fk = is_fk; f->peer_table = y; f->peer_ordinal = yo;
    r->kind = COL_ASSOC; r->n = Y->n_rows; r->forward = false; r->is_fk = false; r->peer_table = x; r->peer_ordinal = xo;
    *fwd_out = f;
    return COLQ_OK;
}

void unlink_assoc(colq_ctx* ctx, colq_table x, int xo, colq_table y, int yo) {
    ctx->tables[x].cols[xo] = Column();
    ctx->tables[y].cols[yo] = Column();
}

// size of one half of the peer heap (global bitmaps of cross-shard hops): COLQ_PEER_HEAP_MB, default 128 MB = 1 G rows of bitmap
size_t peer_heap_half_bytes() {
    const char* e = getenv("COLQ_PEER_HEAP_MB");
    const double mb = e ? atof(e) : 128.0;
    return (size_t)round_up((int64_t)(std::max(mb, 1.0) * 1048576.0), 256);
}

// CUDA-IPC mailboxes for the peer-memory exchange kernels.  Collective over the freshly created NCCL communicator
// (used here only to ship the 64-byte IPC handles and to agree on success).  Any failure leaves peer.ok == false and
// the data path falls back to NCCL all-gathers.
colq_status setup_peerbox(colq_ctx* ctx) {
    auto& pb = ctx->peer;
    const char* env = getenv("COLQ_PEER");
    if (ctx->n_ranks < 2 || (env && env[0] == '0')) return COLQ_OK;
    cudaStream_t s = ctx->stream;
    const char* cap_env = getenv("COLQ_PEER_SLOT_CAP");
    pb.slot_cap = cap_env ? std::max<int64_t>(1024, atoll(cap_env)) : ((int64_t)1 << 20);
    pb.slot_bytes = (size_t)GATHER_SLOT_HEADER + (size_t)pb.slot_cap * 8;
    pb.heap_off = (size_t)round_up((int64_t)(PEER_GATHER_AREA_OFFSET + (size_t)2 * ctx->n_ranks * pb.slot_bytes), 256);
    pb.heap_half = peer_heap_half_bytes();
    pb.bytes = pb.heap_off + 2 * pb.heap_half;
    struct Msg { cudaIpcMemHandle_t handle; int ok; int pad[15]; };
    static_assert(sizeof(Msg) == 128, "message layout");
    Msg mine{};
    mine.ok = 1;
    if (cudaMalloc(&pb.local, pb.bytes) != cudaSuccess) { mine.ok = 0; pb.local = nullptr; }
    if (mine.ok && cudaMemset(pb.local, 0, pb.bytes) != cudaSuccess) mine.ok = 0;
    if (mine.ok && cudaIpcGetMemHandle(&mine.handle, pb.local) != cudaSuccess) mine.ok = 0;
    cudaGetLastError();
    DevBuf send, recv;
    ST(dev_alloc(ctx, send, sizeof(Msg)));
    ST(dev_alloc(ctx, recv, sizeof(Msg) * ctx->n_ranks));
    std::vector<Msg> all(ctx->n_ranks);
    auto exchange = [&]() -> colq_status {
        CU(ctx, cudaMemcpyAsync(send.ptr, &mine, sizeof(Msg), cudaMemcpyHostToDevice, s));
        NC(ctx, ctx->nccl.AllGather(send.ptr, recv.ptr, sizeof(Msg), kNcclUint8, ctx->comm, (void*)s));
        CU(ctx, cudaMemcpyAsync(all.data(), recv.ptr, sizeof(Msg) * ctx->n_ranks, cudaMemcpyDeviceToHost, s));
        CU(ctx, cudaStreamSynchronize(s));
        return COLQ_OK;
    };
    ST(exchange());
    bool all_ok = true;
    for (const Msg& m : all) all_ok = all_ok && m.ok;
    if (all_ok) {
        for (int r = 0; r < ctx->n_ranks; ++r) {
            if (r == ctx->rank) { pb.peer_ptr[r] = pb.local; continue; }
            if (cudaIpcOpenMemHandle(&pb.peer_ptr[r], all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                pb.peer_ptr[r] = nullptr;
                mine.ok = 0;
                cudaGetLastError();
            }
        }
    } else {
        mine.ok = 0;
    }
    ST(exchange());  // second round: did every rank map every mailbox?
    all_ok = true;
    for (const Msg& m : all) all_ok = all_ok && m.ok;
    if (!all_ok) {
        for (int r = 0; r < ctx->n_ranks; ++r)
            if (r != ctx->rank && pb.peer_ptr[r]) cudaIpcCloseMemHandle(pb.peer_ptr[r]);
        if (pb.local) cudaFree(pb.local);
        pb = colq_ctx::PeerBox();
        cudaGetLastError();
        return COLQ_OK;
    }
    void* p;
    CU(ctx, cudaMalloc(&p, sizeof(void*) * MAX_RANKS));
    pb.d_peers = (uint8_t**)p;
    CU(ctx, cudaMemcpy(pb.d_peers, pb.peer_ptr, sizeof(void*) * MAX_RANKS, cudaMemcpyHostToDevice));
    CU(ctx, cudaMalloc(&p, sizeof(u32) * (MAX_RANKS + 16)));
    CU(ctx, cudaMemset(p, 0, sizeof(u32) * (MAX_RANKS + 16)));
    pb.d_done = (u32*)p;
    pb.d_status = pb.d_done + MAX_RANKS;
    pb.ok = true;
    return COLQ_OK;
}

void destroy_peerbox(colq_ctx* ctx) {
    auto& pb = ctx->peer;
    for (int r = 0; r < ctx->n_ranks; ++r)
        if (pb.ipc && r != ctx->rank && pb.peer_ptr[r]) cudaIpcCloseMemHandle(pb.peer_ptr[r]);
    if (pb.local) cudaFree(pb.local);
    if (pb.d_peers) cudaFree(pb.d_peers);
    if (pb.d_done) cudaFree(pb.d_done);
    pb = colq_ctx::PeerBox();
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================

extern "C" {

int colq_abi_version(void) { return COLQ_ABI_VERSION; }

#ifndef COLQ_BUILD_ID
#define COLQ_BUILD_ID "unknown"
#endif
const char* colq_build_id(void) { return COLQ_BUILD_ID; }

colq_status colq_create(int device, colq_ctx** out_ctx) {
    if (!out_ctx) return COLQ_THROW_NULL;
    *out_ctx = nullptr;
    std::unique_ptr<colq_ctx> ctx(new colq_ctx());
    ctx->device = device;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) {
        // no CPU fallback: the module is useless without its GPU (north_star)
        fprintf(stderr, "colq_create: no usable CUDA device %d (%s)\n", device, e != cudaSuccess ? cudaGetErrorString(e) : "device count");
        return COLQ_ERR_DEVICE;
    }
    if (cudaSetDevice(device) != cudaSuccess) return COLQ_ERR_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return COLQ_ERR_DEVICE;
    if (prop.major < 10) {
        fprintf(stderr, "colq_create: device %d is sm_%d%d; libcolq is built for sm_100a only\n", device, prop.major, prop.minor);
        return COLQ_ERR_DEVICE;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) return COLQ_ERR_DEVICE;
    ctx->stream = ctx->own_stream;
    {
        void* p = nullptr;
        if (cudaMalloc(&p, 64) != cudaSuccess || cudaMemset(p, 0, 64) != cudaSuccess) return COLQ_ERR_DEVICE;
        ctx->d_tile_counters = (u32*)p;
    }
    {
        DeviceCache& cache = device_cache();
        std::lock_guard<std::mutex> g(cache.mu);
        cache.live_contexts++;
    }
    *out_ctx = ctx.release();
    return COLQ_OK;
}

colq_status colq_destroy(colq_ctx* ctx) {
    if (!ctx) return COLQ_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    while (!ctx->queries.empty()) colq_query_destroy(ctx->queries.back());  // a context owns its queries
    destroy_peerbox(ctx);
    if (ctx->comm && ctx->nccl.CommDestroy) ctx->nccl.CommDestroy(ctx->comm);
    ctx->tables.clear();
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->d_tile_counters) cudaFree(ctx->d_tile_counters);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    DeviceCache& cache = device_cache();
    bool last;
    {
        std::lock_guard<std::mutex> g(cache.mu);
        last = --cache.live_contexts == 0;
    }
    if (last) cache.trim();  // the last context of this device returns every parked block to the driver
    return COLQ_OK;
}

colq_status colq_trim(colq_ctx* ctx) {
    if (!ctx) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    device_cache().trim();
    return COLQ_OK;
}

const char* colq_last_error(const colq_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

colq_status colq_set_stream(colq_ctx* ctx, void* cuda_stream) {
    if (!ctx) return COLQ_THROW_NULL;
    // buffers are recycled in stream order (DeviceCache): drain the old stream before work moves to another one
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    ctx->chain_query = nullptr;
    return COLQ_OK;
}

colq_status colq_get_stream(colq_ctx* ctx, void** out) {
    if (!ctx || !out) return COLQ_THROW_NULL;
    *out = (void*)ctx->stream;
    return COLQ_OK;
}

colq_status colq_synchronize(colq_ctx* ctx) {
    if (!ctx) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return COLQ_OK;
}

colq_status colq_comm_unique_id(colq_ctx* ctx, uint8_t out_id[128]) {
    if (!ctx || !out_id) return COLQ_THROW_NULL;
    std::string err;
    if (!ctx->nccl.load(err)) return fail(ctx, COLQ_ERR_DEVICE, "%s", err.c_str());
    ncclUniqueId id;
    NC(ctx, ctx->nccl.GetUniqueId(&id));
    memcpy(out_id, id.internal, 128);
    return COLQ_OK;
}

colq_status colq_comm_init(colq_ctx* ctx, const uint8_t id_bytes[128], int n_ranks, int rank) {
    if (!ctx || !id_bytes) return COLQ_THROW_NULL;
    if (n_ranks < 1 || n_ranks > MAX_RANKS || rank < 0 || rank >= n_ranks) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "bad rank %d of %d", rank, n_ranks);
    if (ctx->comm) return fail(ctx, COLQ_THROW_ILLEGAL_STATE, "communicator already initialised");
    std::string err;
    if (!ctx->nccl.load(err)) return fail(ctx, COLQ_ERR_DEVICE, "%s", err.c_str());
    CU(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(id.internal, id_bytes, 128);
    NC(ctx, ctx->nccl.CommInitRank(&ctx->comm, n_ranks, id, rank));
    ctx->n_ranks = n_ranks;
    ctx->rank = rank;
    return setup_peerbox(ctx);
}

// One host process (the reference is ONE DataSystemSerialIndices object in one JVM, E/DataSystemSerialIndices.java:14-22)
// driving one context per GPU: the mailboxes are plain cudaMalloc allocations made reachable with
// cudaDeviceEnablePeerAccess, the exchange kernels are the same as in the one-process-per-GPU mode.
colq_status colq_comm_init_local(colq_ctx** ctxs, int n_ranks) {
    if (!ctxs) return COLQ_THROW_NULL;
    if (n_ranks < 1 || n_ranks > MAX_RANKS) return COLQ_THROW_ILLEGAL_ARG;
    for (int i = 0; i < n_ranks; ++i) {
        if (!ctxs[i]) return COLQ_THROW_NULL;
        if (ctxs[i]->comm || ctxs[i]->n_ranks != 1) return fail(ctxs[i], COLQ_THROW_ILLEGAL_STATE, "communicator already initialised");
        for (int j = 0; j < i; ++j)
            if (ctxs[j]->device == ctxs[i]->device) return fail(ctxs[i], COLQ_THROW_ILLEGAL_ARG, "contexts %d and %d are on the same GPU %d", j, i, ctxs[i]->device);
    }
    if (n_ranks == 1) return COLQ_OK;
    colq_ctx* c0 = ctxs[0];
    const char* cap_env = getenv("COLQ_PEER_SLOT_CAP");
    const int64_t slot_cap = cap_env ? std::max<int64_t>(1024, atoll(cap_env)) : ((int64_t)1 << 20);
    const size_t slot_bytes = (size_t)GATHER_SLOT_HEADER + (size_t)slot_cap * 8;
    const size_t heap_off = (size_t)round_up((int64_t)(PEER_GATHER_AREA_OFFSET + (size_t)2 * n_ranks * slot_bytes), 256);
    const size_t heap_half = peer_heap_half_bytes();
    const size_t bytes = heap_off + 2 * heap_half;
    for (int i = 0; i < n_ranks; ++i) {
        colq_ctx* c = ctxs[i];
        CU(c, cudaSetDevice(c->device));
        for (int j = 0; j < n_ranks; ++j) {
            if (j == i) continue;
            int can = 0;
            CU(c, cudaDeviceCanAccessPeer(&can, c->device, ctxs[j]->device));
            if (!can) return fail(c0, COLQ_ERR_DEVICE, "GPU %d cannot access GPU %d's memory (no NVLink / P2P path)", c->device, ctxs[j]->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(c0, COLQ_ERR_DEVICE, "cudaDeviceEnablePeerAccess(%d -> %d): %s", c->device, ctxs[j]->device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        auto& pb = c->peer;
        pb.slot_cap = slot_cap; pb.slot_bytes = slot_bytes; pb.bytes = bytes; pb.ipc = false;
        pb.heap_off = heap_off; pb.heap_half = heap_half;
        CU(c, cudaMalloc(&pb.local, bytes));
        CU(c, cudaMemset(pb.local, 0, bytes));
    }
    for (int i = 0; i < n_ranks; ++i) {
        colq_ctx* c = ctxs[i];
        CU(c, cudaSetDevice(c->device));
        auto& pb = c->peer;
        for (int r = 0; r < n_ranks; ++r) pb.peer_ptr[r] = ctxs[r]->peer.local;
        void* p;
        CU(c, cudaMalloc(&p, sizeof(void*) * MAX_RANKS));
        pb.d_peers = (uint8_t**)p;
        CU(c, cudaMemcpy(pb.d_peers, pb.peer_ptr, sizeof(void*) * MAX_RANKS, cudaMemcpyHostToDevice));
        CU(c, cudaMalloc(&p, sizeof(u32) * (MAX_RANKS + 16)));
        CU(c, cudaMemset(p, 0, sizeof(u32) * (MAX_RANKS + 16)));
        pb.d_done = (u32*)p;
        pb.d_status = pb.d_done + MAX_RANKS;
        pb.ok = true;
        c->n_ranks = n_ranks;
        c->rank = i;
        CU(c, cudaDeviceSynchronize());
    }
    return COLQ_OK;
}

// colq_execute_async on every context of a group (a local communicator's kernels wait for one another across GPUs, so
// every rank's work must be enqueued before any rank's result is fetched)
colq_status colq_execute_group(colq_ctx** ctxs, colq_query** queries, int n) {
    if (!ctxs || !queries) return COLQ_THROW_NULL;
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i] || !queries[i]) return COLQ_THROW_NULL;
        if (queries[i]->ctx != ctxs[i]) return fail(ctxs[i], COLQ_THROW_ILLEGAL_ARG, "query %d belongs to another context", i);
    }
    colq_status first = COLQ_OK;
    for (int i = 0; i < n; ++i) {
        // a Failure is the same on every rank (same schema, same query); a rank that failed to enqueue must not leave its
        // peers spinning, so stop at the first one
        colq_status st = run_pipeline(queries[i]);
        if (st != COLQ_OK) { first = st; break; }
    }
    return first;
}

// Waits for every rank's execution and reports the match counts; if a rank's result block turned out too small, the
// query is re-run on ALL ranks with a larger block first (the ranks decide alike: they all see the gathered counts).
colq_status colq_fetch_group(colq_ctx** ctxs, colq_query** queries, int n, int64_t* out_counts) {
    if (!ctxs || !queries) return COLQ_THROW_NULL;
    for (int i = 0; i < n; ++i)
        if (!ctxs[i] || !queries[i] || queries[i]->ctx != ctxs[i]) return COLQ_THROW_ILLEGAL_ARG;
    for (int attempt = 0; attempt < 4; ++attempt) {
        bool rerun = false;
        int64_t want = 0;
        for (int i = 0; i < n; ++i) {
            int64_t count = 0;
            colq_status st = fetch_results(queries[i], nullptr, 0, nullptr, 0, &count, nullptr);
            if (st == RERUN_GROUP) rerun = true;
            else if (st != COLQ_OK) return st;
            else if (out_counts) out_counts[i] = count;
            want = std::max(want, queries[i]->want_idx_capacity);
        }
        if (!rerun) return COLQ_OK;
        for (int i = 0; i < n; ++i) queries[i]->want_idx_capacity = want;
        ST(colq_execute_group(ctxs, queries, n));
    }
    return fail(ctxs[0], COLQ_ERR_CAPACITY, "result block still too small after re-running the group");
}

colq_status colq_comm_info(const colq_ctx* ctx, int* out_n_ranks, int* out_rank) {
    if (!ctx) return COLQ_THROW_NULL;
    if (out_n_ranks) *out_n_ranks = ctx->n_ranks;
    if (out_rank) *out_rank = ctx->rank;
    return COLQ_OK;
}

colq_status colq_table_create(colq_ctx* ctx, int64_t n_rows, colq_placement placement, int64_t global_row_base, colq_table* out) {
    if (!ctx || !out) return COLQ_THROW_NULL;
    if (n_rows < 0 || n_rows > INT32_MAX) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "row count %lld outside [0, 2^31)", (long long)n_rows);
    if (placement != COLQ_REPLICATED && placement != COLQ_SHARDED) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "bad placement %d", (int)placement);
    // result rows are int32 (a Java int / BitSet index): the GLOBAL index of a sharded table's last row must fit too
    if (global_row_base < 0 || global_row_base + n_rows > (int64_t)INT32_MAX)
        return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "global rows [%lld, %lld) exceed the int32 row-index range of the result", (long long)global_row_base,
                    (long long)(global_row_base + n_rows));
    Table t;
    t.n_rows = n_rows; t.placement = placement; t.row_base = global_row_base;
    ctx->tables.push_back(std::move(t));
    *out = (colq_table)ctx->tables.size() - 1;
    return COLQ_OK;
}

colq_status colq_register(colq_ctx* ctx, const char* name, colq_table table) {
    if (!ctx || !name) return COLQ_THROW_NULL;
    if (!get_table(ctx, table)) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown table handle %d", table);
    ctx->registry[name] = table;  // HashMap.put (E/DataSystemSerialIndices.java:28)
    return COLQ_OK;
}

colq_status colq_col_i32(colq_ctx* ctx, colq_table table, int ordinal, const int32_t* values, int64_t n) {
    if (!ctx || (!values && n > 0)) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    Column* c;
    ST(slot_for(ctx, table, ordinal, n, &c));
    ST(upload(ctx, c->data, values, (size_t)n * 4, (size_t)round_up(n * 4 + 16, 16)));
    c->kind = COL_I32; c->n = n;
    return COLQ_OK;
}

colq_status colq_col_i32_device(colq_ctx* ctx, colq_table table, int ordinal, const void* values_device, int64_t n) {
    if (!ctx || (!values_device && n > 0)) return COLQ_THROW_NULL;
    if ((uintptr_t)values_device & 15) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "device column must be 16-byte aligned");
    Column* c;
    ST(slot_for(ctx, table, ordinal, n, &c));
    c->data.ptr = const_cast<void*>(values_device); c->data.bytes = (size_t)n * 4; c->data.owned = false;
    c->kind = COL_I32; c->n = n;
    return COLQ_OK;
}

static colq_status finish_str(colq_ctx* ctx, Column* c, int64_t n, int64_t n_bytes) {
    c->kind = COL_STR; c->n = n; c->n_bytes = n_bytes;
    c->max_tile_bytes = 0;
    if (n > 0) {
        DevBuf mx;
        ST(dev_alloc(ctx, mx, 4));
        CU(ctx, cudaMemsetAsync(mx.ptr, 0, 4, ctx->stream));
        const int64_t n_tiles = (n + ST_ROWS - 1) / ST_ROWS;
        tile_payload_max_kernel<<<grid_for(n_tiles, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>((const u32*)c->offsets.ptr, n, (u32*)mx.ptr);
        CU(ctx, cudaGetLastError());
        u32 got = 0;
        CU(ctx, cudaMemcpyAsync(&got, mx.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        c->max_tile_bytes = got;
    }
    return COLQ_OK;
}

// upload an (offsets, bytes) pair into `c` with the padding the TMA path needs
static colq_status fill_str(colq_ctx* ctx, Column* c, const uint32_t* offsets, const uint8_t* bytes, int64_t n, int64_t n_bytes) {
    ST(upload(ctx, c->offsets, offsets, (size_t)(n + 1) * 4, (size_t)round_up((n + 1) * 4, 16) + 16));
    size_t cap = (size_t)round_up(n_bytes, 16) + ST_SLACK;
    ST(upload(ctx, c->data, bytes, (size_t)n_bytes, cap));
    c->bytes_capacity = (int64_t)cap;
    return finish_str(ctx, c, n, n_bytes);
}

colq_status colq_col_str(colq_ctx* ctx, colq_table table, int ordinal, const uint32_t* offsets, const uint8_t* bytes, int64_t n,
                         int64_t n_bytes) {
    if (!ctx || !offsets || (!bytes && n_bytes > 0)) return COLQ_THROW_NULL;
    if (n_bytes < 0 || n_bytes > (int64_t)0xfffffff0ll) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "string payload of %lld bytes exceeds the uint32 offset range", (long long)n_bytes);
    if (offsets[0] != 0 || (int64_t)offsets[n] != n_bytes) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "offsets must start at 0 and end at n_bytes");
    CU(ctx, cudaSetDevice(ctx->device));
    Column* c;
    ST(slot_for(ctx, table, ordinal, n, &c));
    return fill_str(ctx, c, offsets, bytes, n, n_bytes);
}

// ---- dictionary-encoded string columns ---------------------------------------------------------------

static colq_status pinned_alias(colq_ctx* ctx, const void* host, const char* what, const void** out);

static colq_status check_dict_args(colq_ctx* ctx, const uint32_t* dict_offsets, const uint8_t* dict_bytes, int64_t n_dict, int64_t n_dict_bytes) {
    if (!dict_offsets || (!dict_bytes && n_dict_bytes > 0)) return COLQ_THROW_NULL;
    if (n_dict < 0 || n_dict > INT32_MAX) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "dictionary size %lld outside [0, 2^31)", (long long)n_dict);
    if (n_dict_bytes < 0 || n_dict_bytes > (int64_t)0xfffffff0ll) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "dictionary payload exceeds the uint32 offset range");
    if (dict_offsets[0] != 0 || (int64_t)dict_offsets[n_dict] != n_dict_bytes) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "dictionary offsets must start at 0 and end at n_dict_bytes");
    return COLQ_OK;
}

// codes are in HBM (or pinned host memory) at c->data; attach the dictionary and validate the code range
static colq_status finish_dict(colq_ctx* ctx, Column* c, int64_t n, const uint32_t* dict_offsets, const uint8_t* dict_bytes, int64_t n_dict,
                               int64_t n_dict_bytes, bool check_codes) {
    std::unique_ptr<Column> d(new Column());
    ST(fill_str(ctx, d.get(), dict_offsets, dict_bytes, n_dict, n_dict_bytes));
    if (check_codes && n > 0) {
        DevBuf mm;
        ST(dev_alloc(ctx, mm, 8));
        int32_t init[2] = {INT32_MAX, INT32_MIN};
        CU(ctx, cudaMemcpyAsync(mm.ptr, init, 8, cudaMemcpyHostToDevice, ctx->stream));
        fk_minmax_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>((const int32_t*)c->data.ptr, n, (int32_t*)mm.ptr, (int32_t*)mm.ptr + 1);
        CU(ctx, cudaGetLastError());
        int32_t got[2];
        CU(ctx, cudaMemcpyAsync(got, mm.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (got[0] < 0 || got[1] >= n_dict)
            return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "dictionary code outside [0, %lld) (min %d, max %d)", (long long)n_dict, got[0], got[1]);
    }
    c->kind = COL_STR; c->n = n; c->n_bytes = 0;
    c->dict = std::move(d);
    return COLQ_OK;
}

colq_status colq_col_str_dict(colq_ctx* ctx, colq_table table, int ordinal, const int32_t* codes, int64_t n, const uint32_t* dict_offsets,
                              const uint8_t* dict_bytes, int64_t n_dict, int64_t n_dict_bytes) {
    if (!ctx || (!codes && n > 0)) return COLQ_THROW_NULL;
    ST(check_dict_args(ctx, dict_offsets, dict_bytes, n_dict, n_dict_bytes));
    CU(ctx, cudaSetDevice(ctx->device));
    Column* c;
    ST(slot_for(ctx, table, ordinal, n, &c));
    ST(upload(ctx, c->data, codes, (size_t)n * 4, (size_t)round_up(n * 4 + 16, 16)));
    colq_status st = finish_dict(ctx, c, n, dict_offsets, dict_bytes, n_dict, n_dict_bytes, true);
    if (st != COLQ_OK) *c = Column();
    return st;
}

colq_status colq_col_str_dict_device(colq_ctx* ctx, colq_table table, int ordinal, const void* codes_device, int64_t n,
                                     const uint32_t* dict_offsets, const uint8_t* dict_bytes, int64_t n_dict, int64_t n_dict_bytes) {
    if (!ctx || (!codes_device && n > 0)) return COLQ_THROW_NULL;
    if ((uintptr_t)codes_device & 15) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "device column must be 16-byte aligned");
    ST(check_dict_args(ctx, dict_offsets, dict_bytes, n_dict, n_dict_bytes));
    CU(ctx, cudaSetDevice(ctx->device));
    Column* c;
    ST(slot_for(ctx, table, ordinal, n, &c));
    c->data.ptr = const_cast<void*>(codes_device); c->data.bytes = (size_t)n * 4; c->data.owned = false;
    colq_status st = finish_dict(ctx, c, n, dict_offsets, dict_bytes, n_dict, n_dict_bytes, true);
    if (st != COLQ_OK) *c = Column();
    return st;
}

// IntegerColumn stored dictionary-encoded: codes + the DISTINCT int32 values
static colq_status finish_dict_i32(colq_ctx* ctx, Column* c, int64_t n, const int32_t* dict_values, int64_t n_dict, bool check_codes) {
    if (n_dict < 0 || n_dict > INT32_MAX) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "dictionary size %lld outside [0, 2^31)", (long long)n_dict);
    std::unique_ptr<Column> d(new Column());
    ST(upload(ctx, d->data, dict_values, (size_t)n_dict * 4, (size_t)round_up(n_dict * 4 + 16, 16)));
    d->kind = COL_I32; d->n = n_dict;
    if (check_codes && n > 0) {
        DevBuf mm;
        ST(dev_alloc(ctx, mm, 8));
        int32_t init[2] = {INT32_MAX, INT32_MIN};
        CU(ctx, cudaMemcpyAsync(mm.ptr, init, 8, cudaMemcpyHostToDevice, ctx->stream));
        fk_minmax_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>((const int32_t*)c->data.ptr, n, (int32_t*)mm.ptr, (int32_t*)mm.ptr + 1);
        CU(ctx, cudaGetLastError());
        int32_t got[2];
        CU(ctx, cudaMemcpyAsync(got, mm.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (got[0] < 0 || got[1] >= n_dict)
            return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "dictionary code outside [0, %lld) (min %d, max %d)", (long long)n_dict, got[0], got[1]);
    }
    c->kind = COL_I32; c->n = n;
    c->dict = std::move(d);
    return COLQ_OK;
}

colq_status colq_col_i32_dict(colq_ctx* ctx, colq_table table, int ordinal, const int32_t* codes, int64_t n, const int32_t* dict_values,
                              int64_t n_dict) {
    if (!ctx || (!codes && n > 0) || (!dict_values && n_dict > 0)) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    Column* c;
    ST(slot_for(ctx, table, ordinal, n, &c));
    ST(upload(ctx, c->data, codes, (size_t)n * 4, (size_t)round_up(n * 4 + 16, 16)));
    colq_status st = finish_dict_i32(ctx, c, n, dict_values, n_dict, true);
    if (st != COLQ_OK) *c = Column();
    return st;
}

colq_status colq_col_i32_dict_host(colq_ctx* ctx, colq_table table, int ordinal, const int32_t* codes_pinned, int64_t capacity_bytes,
                                   int64_t n, const int32_t* dict_values, int64_t n_dict) {
    if (!ctx || !codes_pinned || (!dict_values && n_dict > 0)) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    if (capacity_bytes < round_up(n * 4, 16)) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "host column buffer must be padded to a multiple of 16 bytes");
    const void* alias;
    ST(pinned_alias(ctx, codes_pinned, "the dictionary-code buffer", &alias));
    Column* c;
    ST(slot_for(ctx, table, ordinal, n, &c));
    c->data.ptr = const_cast<void*>(alias); c->data.bytes = (size_t)capacity_bytes; c->data.owned = false;
    colq_status st = finish_dict_i32(ctx, c, n, dict_values, n_dict, false);
    if (st != COLQ_OK) *c = Column();
    else c->host_resident = true;
    return st;
}

colq_status colq_col_str_dict_host(colq_ctx* ctx, colq_table table, int ordinal, const int32_t* codes_pinned, int64_t capacity_bytes,
                                   int64_t n, const uint32_t* dict_offsets, const uint8_t* dict_bytes, int64_t n_dict, int64_t n_dict_bytes) {
    if (!ctx || !codes_pinned) return COLQ_THROW_NULL;
    ST(check_dict_args(ctx, dict_offsets, dict_bytes, n_dict, n_dict_bytes));
    CU(ctx, cudaSetDevice(ctx->device));
    if (capacity_bytes < round_up(n * 4, 16)) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "host column buffer must be padded to a multiple of 16 bytes");
    const void* alias;
    ST(pinned_alias(ctx, codes_pinned, "the dictionary-code buffer", &alias));
    Column* c;
    ST(slot_for(ctx, table, ordinal, n, &c));
    c->data.ptr = const_cast<void*>(alias); c->data.bytes = (size_t)capacity_bytes; c->data.owned = false;
    // codes stay in host memory and are not read at registration: the row scan treats a code outside the dictionary
    // as "no match" (it fails the [0, n_dict) range test before the mask lookup)
    colq_status st = finish_dict(ctx, c, n, dict_offsets, dict_bytes, n_dict, n_dict_bytes, false);
    if (st != COLQ_OK) *c = Column();
    else c->host_resident = true;
    return st;
}

colq_status colq_col_str_device(colq_ctx* ctx, colq_table table, int ordinal, const void* offsets_device, int64_t offsets_capacity,
                                const void* bytes_device, int64_t bytes_capacity, int64_t n, int64_t n_bytes) {
    if (!ctx || !offsets_device || (!bytes_device && n_bytes > 0)) return COLQ_THROW_NULL;
    if (((uintptr_t)offsets_device & 15) || ((uintptr_t)bytes_device & 15)) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "device buffers must be 16-byte aligned");
    if (n_bytes < 0 || n_bytes > (int64_t)0xfffffff0ll) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "string payload of %lld bytes exceeds the uint32 offset range", (long long)n_bytes);
    CU(ctx, cudaSetDevice(ctx->device));
    Column* c;
    ST(slot_for(ctx, table, ordinal, n, &c));
    const int64_t need_off = round_up((n + 1) * 4, 16), need_bytes = round_up(n_bytes, 16) + ST_SLACK;
    if (offsets_capacity >= need_off) {
        c->offsets.ptr = const_cast<void*>(offsets_device); c->offsets.bytes = (size_t)offsets_capacity; c->offsets.owned = false;
    } else {  // too tight for whole-line TMA reads: keep a padded private copy
        ST(dev_alloc(ctx, c->offsets, (size_t)need_off + 16));
        CU(ctx, cudaMemsetAsync(c->offsets.ptr, 0, c->offsets.bytes, ctx->stream));
        CU(ctx, cudaMemcpyAsync(c->offsets.ptr, offsets_device, (size_t)(n + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (bytes_capacity >= need_bytes) {
        c->data.ptr = const_cast<void*>(bytes_device); c->data.bytes = (size_t)bytes_capacity; c->data.owned = false;
        c->bytes_capacity = bytes_capacity & ~(int64_t)15;
    } else {
        ST(dev_alloc(ctx, c->data, (size_t)need_bytes));
        CU(ctx, cudaMemsetAsync(c->data.ptr, 0, c->data.bytes, ctx->stream));
        if (n_bytes) CU(ctx, cudaMemcpyAsync(c->data.ptr, bytes_device, (size_t)n_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        c->bytes_capacity = need_bytes;
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return finish_str(ctx, c, n, n_bytes);
}

// ---- pinned host memory and host-resident columns -----------------------------------------------------

colq_status colq_host_alloc(colq_ctx* ctx, int64_t bytes, void** out_ptr) {
    if (!ctx || !out_ptr) return COLQ_THROW_NULL;
    *out_ptr = nullptr;
    if (bytes < 0) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "negative size");
    CU(ctx, cudaSetDevice(ctx->device));
    void* p = nullptr;
    CU(ctx, cudaHostAlloc(&p, (size_t)std::max<int64_t>(bytes, 16), cudaHostAllocMapped | cudaHostAllocPortable));
    *out_ptr = p;
    return COLQ_OK;
}

colq_status colq_host_free(colq_ctx* ctx, void* ptr) {
    if (!ctx) return COLQ_THROW_NULL;
    if (!ptr) return COLQ_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaFreeHost(ptr));
    return COLQ_OK;
}

colq_status colq_host_register(colq_ctx* ctx, void* ptr, int64_t bytes) {
    if (!ctx || !ptr) return COLQ_THROW_NULL;
    if (bytes <= 0) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "non-positive size");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    return COLQ_OK;
}

colq_status colq_host_unregister(colq_ctx* ctx, void* ptr) {
    if (!ctx || !ptr) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaHostUnregister(ptr));
    return COLQ_OK;
}

// the device-side alias of a pinned host buffer (identical to the host address under UVA for cudaHostAlloc memory)
static colq_status pinned_alias(colq_ctx* ctx, const void* host, const char* what, const void** out) {
    cudaPointerAttributes at{};
    cudaError_t e = cudaPointerGetAttributes(&at, host);
    if (e != cudaSuccess || at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) {
        cudaGetLastError();
        return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "%s is not pinned host memory (allocate it with colq_host_alloc or pin it with colq_host_register)", what);
    }
    if ((uintptr_t)at.devicePointer & 15) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "%s must be 16-byte aligned", what);
    *out = at.devicePointer;
    return COLQ_OK;
}

colq_status colq_col_i32_host(colq_ctx* ctx, colq_table table, int ordinal, const int32_t* values_pinned, int64_t capacity_bytes, int64_t n) {
    if (!ctx || !values_pinned) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    if (capacity_bytes < round_up(n * 4, 16)) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "host column buffer must be padded to a multiple of 16 bytes (need %lld, have %lld)", (long long)round_up(n * 4, 16), (long long)capacity_bytes);
    const void* alias;
    ST(pinned_alias(ctx, values_pinned, "the int column buffer", &alias));
    Column* c;
    ST(slot_for(ctx, table, ordinal, n, &c));
    c->data.ptr = const_cast<void*>(alias); c->data.bytes = (size_t)capacity_bytes; c->data.owned = false;
    c->kind = COL_I32; c->n = n; c->host_resident = true;
    return COLQ_OK;
}

colq_status colq_col_str_host(colq_ctx* ctx, colq_table table, int ordinal, const uint32_t* offsets_pinned, int64_t offsets_capacity,
                              const uint8_t* bytes_pinned, int64_t bytes_capacity, int64_t n, int64_t n_bytes) {
    if (!ctx || !offsets_pinned || !bytes_pinned) return COLQ_THROW_NULL;
    if (n_bytes < 0 || n_bytes > (int64_t)0xfffffff0ll) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "string payload of %lld bytes exceeds the uint32 offset range", (long long)n_bytes);
    if (offsets_pinned[0] != 0 || (int64_t)offsets_pinned[n] != n_bytes) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "offsets must start at 0 and end at n_bytes");
    const int64_t need_off = round_up((n + 1) * 4, 16), need_bytes = round_up(n_bytes, 16) + ST_SLACK;
    if (offsets_capacity < need_off || bytes_capacity < need_bytes)
        return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "host string buffers are too tight for whole-line reads: offsets need %lld bytes (have %lld), bytes need %lld (have %lld)",
                    (long long)need_off, (long long)offsets_capacity, (long long)need_bytes, (long long)bytes_capacity);
    CU(ctx, cudaSetDevice(ctx->device));
    const void *oa, *ba;
    ST(pinned_alias(ctx, offsets_pinned, "the string offsets buffer", &oa));
    ST(pinned_alias(ctx, bytes_pinned, "the string bytes buffer", &ba));
    Column* c;
    ST(slot_for(ctx, table, ordinal, n, &c));
    c->offsets.ptr = const_cast<void*>(oa); c->offsets.bytes = (size_t)offsets_capacity; c->offsets.owned = false;
    c->data.ptr = const_cast<void*>(ba); c->data.bytes = (size_t)bytes_capacity; c->data.owned = false;
    c->bytes_capacity = bytes_capacity & ~(int64_t)15;
    c->host_resident = true;
    // the ring slot size needs the largest tile payload: one offset per 1024 rows, read over PCIe by a tiny kernel
    return finish_str(ctx, c, n, n_bytes);
}

colq_status colq_associate_fk_host(colq_ctx* ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal, const int32_t* fk_pinned,
                                   int64_t capacity_bytes, int64_t n) {
    if (!ctx || !fk_pinned) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    if (capacity_bytes < round_up(n * 4, 16)) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "host column buffer must be padded to a multiple of 16 bytes");
    const void* alias;
    ST(pinned_alias(ctx, fk_pinned, "the association buffer", &alias));
    Column* f;
    ST(link_assoc(ctx, x, x_ordinal, y, y_ordinal, true, &f));
    if (n != f->n) { unlink_assoc(ctx, x, x_ordinal, y, y_ordinal); return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "association height %lld != table rows", (long long)n); }
    f->data.ptr = const_cast<void*>(alias); f->data.bytes = (size_t)capacity_bytes; f->data.owned = false;
    f->host_resident = true;
    f->fk_validated = false;  // checked on the rows a query walks (kernels flag a bad target; colq_execute reports it)
    return COLQ_OK;
}

colq_status colq_col_bool(colq_ctx* ctx, colq_table table, int ordinal, const uint8_t* values, int64_t n) {
    if (!ctx || (!values && n > 0)) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    Column* c;
    ST(slot_for(ctx, table, ordinal, n, &c));
    ST(upload(ctx, c->data, values, (size_t)n, (size_t)round_up(n + 16, 16)));
    c->kind = COL_BOOL; c->n = n;
    return COLQ_OK;
}

colq_status colq_associate_fk(colq_ctx* ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal, const int32_t* fk, int64_t n) {
    if (!ctx || (!fk && n > 0)) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    Column* f;
    ST(link_assoc(ctx, x, x_ordinal, y, y_ordinal, true, &f));
    if (n != f->n) { unlink_assoc(ctx, x, x_ordinal, y, y_ordinal); return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "association height %lld != table rows", (long long)n); }
    colq_status st = upload(ctx, f->data, fk, (size_t)n * 4, (size_t)round_up(n * 4 + 16, 16));
    if (st == COLQ_OK) st = check_fk_range(ctx, (const int32_t*)f->data.ptr, n, ctx->tables[y].n_rows);
    if (st != COLQ_OK) unlink_assoc(ctx, x, x_ordinal, y, y_ordinal);
    return st;
}

colq_status colq_associate_fk_device(colq_ctx* ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal, const void* fk_device,
                                     int64_t n) {
    if (!ctx || (!fk_device && n > 0)) return COLQ_THROW_NULL;
    if ((uintptr_t)fk_device & 15) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "device column must be 16-byte aligned");
    CU(ctx, cudaSetDevice(ctx->device));
    Column* f;
    ST(link_assoc(ctx, x, x_ordinal, y, y_ordinal, true, &f));
    if (n != f->n) { unlink_assoc(ctx, x, x_ordinal, y, y_ordinal); return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "association height %lld != table rows", (long long)n); }
    f->data.ptr = const_cast<void*>(fk_device); f->data.bytes = (size_t)n * 4; f->data.owned = false;
    colq_status st = check_fk_range(ctx, (const int32_t*)fk_device, n, ctx->tables[y].n_rows);
    if (st != COLQ_OK) unlink_assoc(ctx, x, x_ordinal, y, y_ordinal);
    return st;
}

colq_status colq_associate_csr(colq_ctx* ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal, const int64_t* offsets,
                               const int32_t* targets, int64_t n, int64_t nnz) {
    if (!ctx || !offsets || (!targets && nnz > 0)) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    Table* Y = get_table(ctx, y);
    if (!Y) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown table handle %d", y);
    if (offsets[0] != 0 || offsets[n] != nnz) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "CSR offsets must start at 0 and end at nnz");
    const int64_t n_target = Y->n_rows;
    Column* f;
    ST(link_assoc(ctx, x, x_ordinal, y, y_ordinal, false, &f));
    if (n != f->n) { unlink_assoc(ctx, x, x_ordinal, y, y_ordinal); return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "association height %lld != table rows", (long long)n); }
    f->nnz = nnz;
    colq_status st = upload(ctx, f->offsets, offsets, (size_t)(n + 1) * 8, (size_t)(n + 1) * 8 + 16);
    if (st == COLQ_OK) st = upload(ctx, f->targets, targets, (size_t)nnz * 4, (size_t)nnz * 4 + 16);
    AssocStats a{};
    if (st == COLQ_OK) st = assoc_stats(ctx, (const int64_t*)f->offsets.ptr, (const int32_t*)f->targets.ptr, n, nnz, &a);  // validated on the device
    if (st == COLQ_OK) st = check_csr(ctx, a, nnz, n_target);
    if (st != COLQ_OK) unlink_assoc(ctx, x, x_ordinal, y, y_ordinal);
    return st;
}

// x.associateTo(y, Association[]) with the representation chosen on the device
static colq_status associate_auto(colq_ctx* ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal, DevBuf&& d_off, DevBuf&& d_tgt,
                                  int64_t n, int64_t nnz, int* out_is_fk) {
    Table* Y = get_table(ctx, y);
    if (!Y) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown table handle %d", y);
    const int64_t n_target = Y->n_rows;
    AssocStats a{};
    ST(assoc_stats(ctx, (const int64_t*)d_off.ptr, (const int32_t*)d_tgt.ptr, n, nnz, &a));
    ST(check_csr(ctx, a, nnz, n_target));
    const bool to_one = a.max_degree <= 1;   // every row is Association.None or Association.One (DS/Association.java:27-43)
    Column* f;
    ST(link_assoc(ctx, x, x_ordinal, y, y_ordinal, to_one, &f));
    if (n != f->n) { unlink_assoc(ctx, x, x_ordinal, y, y_ordinal); return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "association height %lld != table rows", (long long)n); }
    if (to_one) {
        colq_status st = dev_alloc(ctx, f->data, (size_t)round_up(n * 4 + 16, 16));
        if (st != COLQ_OK) { unlink_assoc(ctx, x, x_ordinal, y, y_ordinal); return st; }
        cudaMemsetAsync((char*)f->data.ptr + n * 4, 0, f->data.bytes - (size_t)n * 4 < 64 ? f->data.bytes - (size_t)n * 4 : 64, ctx->stream);
        if (n > 0) csr_to_fk_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>((const int64_t*)d_off.ptr, (const int32_t*)d_tgt.ptr, n, (int32_t*)f->data.ptr);
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaStreamSynchronize(ctx->stream));  // the CSR scratch goes back to the buffer cache when this returns
    } else {
        f->nnz = nnz;
        f->offsets = std::move(d_off);
        f->targets = std::move(d_tgt);
    }
    if (out_is_fk) *out_is_fk = to_one ? 1 : 0;
    return COLQ_OK;
}

colq_status colq_associate(colq_ctx* ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal, const int64_t* offsets, const int32_t* targets,
                           int64_t n, int64_t nnz, int* out_is_fk) {
    if (!ctx || !offsets || (!targets && nnz > 0)) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    if (offsets[0] != 0 || offsets[n] != nnz) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "CSR offsets must start at 0 and end at nnz");
    DevBuf d_off, d_tgt;
    ST(upload(ctx, d_off, offsets, (size_t)(n + 1) * 8, (size_t)(n + 1) * 8 + 16));
    ST(upload(ctx, d_tgt, targets, (size_t)nnz * 4, (size_t)nnz * 4 + 16));
    return associate_auto(ctx, x, x_ordinal, y, y_ordinal, std::move(d_off), std::move(d_tgt), n, nnz, out_is_fk);
}

colq_status colq_associate_device(colq_ctx* ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal, const void* offsets_device,
                                  const void* targets_device, int64_t n, int64_t nnz, int* out_is_fk) {
    if (!ctx || !offsets_device || (!targets_device && nnz > 0)) return COLQ_THROW_NULL;
    if (((uintptr_t)offsets_device & 7) || ((uintptr_t)targets_device & 3)) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "device CSR buffers must be naturally aligned");
    CU(ctx, cudaSetDevice(ctx->device));
    int64_t ends[2] = {0, 0};
    CU(ctx, cudaMemcpyAsync(&ends[0], offsets_device, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(&ends[1], (const int64_t*)offsets_device + n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (ends[0] != 0 || ends[1] != nnz) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "CSR offsets must start at 0 and end at nnz");
    // adopted buffers (the caller keeps them alive when the column stays a CSR)
    DevBuf d_off, d_tgt;
    d_off.ptr = const_cast<void*>(offsets_device); d_off.bytes = (size_t)(n + 1) * 8; d_off.owned = false;
    d_tgt.ptr = const_cast<void*>(targets_device); d_tgt.bytes = (size_t)nnz * 4; d_tgt.owned = false;
    return associate_auto(ctx, x, x_ordinal, y, y_ordinal, std::move(d_off), std::move(d_tgt), n, nnz, out_is_fk);
}

// ---- dictionary encoding on the device ------------------------------------------------------------------------------

colq_status colq_col_str_encode(colq_ctx* ctx, colq_table table, int ordinal, int64_t* out_n_dict) {
    if (!ctx) return COLQ_THROW_NULL;
    Table* t = get_table(ctx, table);
    if (!t) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown table handle %d", table);
    if (ordinal < 0 || (size_t)ordinal >= t->cols.size()) return fail(ctx, COLQ_THROW_INDEX_OOB, "Index %d out of bounds for length %d", ordinal, (int)t->cols.size());
    Column& c = t->cols[ordinal];
    if (c.kind != COL_STR) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "column %d is not a string column", ordinal);
    if (c.dict) { if (out_n_dict) *out_n_dict = c.dict->n; return COLQ_OK; }
    CU(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int64_t n = c.n;
    std::unique_ptr<Column> d(new Column());
    DevBuf codes;
    ST(dev_alloc(ctx, codes, (size_t)round_up(n * 4 + 16, 16)));
    CU(ctx, cudaMemsetAsync((char*)codes.ptr + n * 4, 0, std::min<size_t>(codes.bytes - (size_t)n * 4, 64), s));
    int64_t n_dict = 0;
    if (n > 0) {
        // scratch hash table, at most half full.  It starts SMALL -- 1 M slots, 16 MB, L2-resident: a column worth
        // dictionary-encoding has few distinct values, and probes into a table sized for the row count were TLB and L2
        // misses -- and grows 16x whenever the distinct values exceed half of it, up to 64 M slots (1 GB, 32 M distinct
        // values); beyond that dictionary encoding is pointless and the call fails
        const int64_t max_slots = (int64_t)1 << 26;
        int64_t n_slots = 1024;
        while (n_slots < 2 * n && n_slots < ((int64_t)1 << 20)) n_slots <<= 1;
        DevBuf slots, first_bits, first_rows, counts, offs, status;
        ST(dev_alloc(ctx, slots, (size_t)n_slots * sizeof(DictSlot)));
        ST(dev_alloc(ctx, first_bits, (size_t)bitmap_alloc_words(n) * 4));
        ST(dev_alloc(ctx, status, 16));
        const int64_t n_words = bitmap_words(n);
        const int64_t n_blocks = std::max<int64_t>(1, (n_words + CP_WORDS_PER_BLOCK - 1) / CP_WORDS_PER_BLOCK);
        ST(dev_alloc(ctx, counts, (size_t)n_blocks * 4));
        ST(dev_alloc(ctx, offs, (size_t)n_blocks * 8 + 8));
        const u32* off = (const u32*)c.offsets.ptr;
        const uint8_t* bytes = (const uint8_t*)c.data.ptr;
        u32 st_host[2] = {0, 0};
        for (int attempt = 0;; ++attempt) {
            const u64 seed = 0x5DEECE66Dull * (u64)(attempt + 1);
            CU(ctx, cudaMemsetAsync(status.ptr, 0, 16, s));
            dict_init_kernel<<<grid_for(n_slots, 256, ctx->sm_count, 8), 256, 0, s>>>((DictSlot*)slots.ptr, n_slots);
            dict_insert_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, s>>>(off, bytes, n, (DictSlot*)slots.ptr, (u32)(n_slots - 1), seed, (int32_t*)codes.ptr,
                                                                                  (u32*)status.ptr);
            CU(ctx, cudaMemsetAsync(first_bits.ptr, 0, first_bits.bytes, s));
            dict_verify_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, s>>>(off, bytes, n, (const DictSlot*)slots.ptr, (const int32_t*)codes.ptr,
                                                                                  (u32*)first_bits.ptr, (u32*)status.ptr);
            CU(ctx, cudaGetLastError());
            CU(ctx, cudaMemcpyAsync(st_host, status.ptr, 8, cudaMemcpyDeviceToHost, s));
            CU(ctx, cudaStreamSynchronize(s));
            if (st_host[0]) {  // more distinct values than half the table: a larger table, same seed
                if (n_slots >= max_slots || n_slots >= 4 * n)
                    return fail(ctx, COLQ_ERR_CAPACITY, "too many distinct values for dictionary encoding (more than %lld)", (long long)(n_slots / 2));
                n_slots = std::min<int64_t>(n_slots * 16, max_slots);
                ST(dev_alloc(ctx, slots, (size_t)n_slots * sizeof(DictSlot)));
                --attempt;
                continue;
            }
            if (!st_host[1]) break;
            if (attempt == 3) return fail(ctx, COLQ_ERR_DEVICE, "dictionary encoding: 64-bit hash collisions with four different seeds");
        }
        // representatives in ascending row order = first-appearance order (the three-launch compaction)
        popc_blocks_kernel<<<(int)n_blocks, CP_THREADS, 0, s>>>((const u32*)first_bits.ptr, n_words, (u32*)counts.ptr);
        scan_counts_kernel<<<1, 1024, 0, s>>>((const u32*)counts.ptr, n_blocks, (u64*)offs.ptr, (u64*)offs.ptr + n_blocks);
        CU(ctx, cudaGetLastError());
        u64 total = 0;
        CU(ctx, cudaMemcpyAsync(&total, (u64*)offs.ptr + n_blocks, 8, cudaMemcpyDeviceToHost, s));
        CU(ctx, cudaStreamSynchronize(s));
        n_dict = (int64_t)total;
        ST(dev_alloc(ctx, first_rows, (size_t)n_dict * 4 + 16));
        compact_kernel<<<(int)n_blocks, CP_THREADS, 0, s>>>((const u32*)first_bits.ptr, n_words, (const u64*)offs.ptr, (int32_t*)first_rows.ptr, n_dict, 0);
        dict_assign_kernel<<<grid_for(n_dict, 256, ctx->sm_count, 8), 256, 0, s>>>((const int32_t*)first_rows.ptr, n_dict, (const int32_t*)codes.ptr, (DictSlot*)slots.ptr);
        // the distinct values themselves: lengths -> offsets -> bytes at the representative rows
        DevBuf off64;
        ST(dev_alloc(ctx, off64, (size_t)(n_dict + 1) * 8));
        GatherVarParams<u32> G{off, nullptr, (const int32_t*)first_rows.ptr, 0, n_dict, (u64*)off64.ptr};
        gather_var_lens_kernel<u32><<<grid_for(n_dict, 256, ctx->sm_count, 8), 256, 0, s>>>(G);
        scan_u64_inplace_kernel<<<1, 1024, 0, s>>>(G.out_off, n_dict);
        CU(ctx, cudaGetLastError());
        u64 dict_bytes = 0;
        CU(ctx, cudaMemcpyAsync(&dict_bytes, G.out_off + n_dict, 8, cudaMemcpyDeviceToHost, s));
        CU(ctx, cudaStreamSynchronize(s));
        ST(dev_alloc(ctx, d->offsets, (size_t)round_up((n_dict + 1) * 4, 16) + 16));
        CU(ctx, cudaMemsetAsync(d->offsets.ptr, 0, d->offsets.bytes, s));
        narrow_offsets_kernel<<<grid_for(n_dict + 1, 256, ctx->sm_count, 8), 256, 0, s>>>(G.out_off, (u32*)d->offsets.ptr, n_dict + 1);
        const size_t cap = (size_t)round_up((int64_t)dict_bytes, 16) + ST_SLACK;
        ST(dev_alloc(ctx, d->data, cap));
        CU(ctx, cudaMemsetAsync(d->data.ptr, 0, d->data.bytes, s));
        if (dict_bytes > 0)
            gather_var_copy_kernel<u32, uint8_t><<<grid_for((n_dict + 31) / 32 * 32, 256, ctx->sm_count, 8), 256, 0, s>>>(G, bytes, (uint8_t*)d->data.ptr);
        dict_codes_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, s>>>((int32_t*)codes.ptr, n, (const DictSlot*)slots.ptr);
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaStreamSynchronize(s));
        d->bytes_capacity = (int64_t)cap;
        ST(finish_str(ctx, d.get(), n_dict, (int64_t)dict_bytes));
    } else {
        ST(dev_alloc(ctx, d->offsets, 32));
        CU(ctx, cudaMemsetAsync(d->offsets.ptr, 0, 32, s));
        ST(dev_alloc(ctx, d->data, 16 + ST_SLACK));
        d->bytes_capacity = 16 + ST_SLACK;
        ST(finish_str(ctx, d.get(), 0, 0));
    }
    // the column becomes a dictionary-encoded one in place: codes + distinct values; the plain buffers are dropped
    c.data = std::move(codes);
    c.offsets = DevBuf();
    c.promoted = DevBuf();
    c.promoted_offsets = DevBuf();
    c.host_resident = false;
    c.n_bytes = 0;
    c.dict = std::move(d);
    if (out_n_dict) *out_n_dict = n_dict;
    return COLQ_OK;
}

colq_status colq_col_dict_str(colq_ctx* ctx, colq_table table, int ordinal, uint32_t* out_offsets, int64_t offsets_capacity, uint8_t* out_bytes,
                              int64_t bytes_capacity, int64_t* out_n_dict, int64_t* out_n_bytes) {
    if (!ctx) return COLQ_THROW_NULL;
    Table* t = get_table(ctx, table);
    if (!t) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown table handle %d", table);
    if (ordinal < 0 || (size_t)ordinal >= t->cols.size()) return fail(ctx, COLQ_THROW_INDEX_OOB, "Index %d out of bounds for length %d", ordinal, (int)t->cols.size());
    const Column& c = t->cols[ordinal];
    if (c.kind != COL_STR || !c.dict) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "column %d is not a dictionary-encoded string column", ordinal);
    const Column& d = *c.dict;
    if (out_n_dict) *out_n_dict = d.n;
    if (out_n_bytes) *out_n_bytes = d.n_bytes;
    if (!out_offsets || offsets_capacity < d.n + 1 || (d.n_bytes > 0 && (!out_bytes || bytes_capacity < d.n_bytes)))
        return fail(ctx, COLQ_ERR_CAPACITY, "dictionary needs %lld offsets and %lld bytes", (long long)(d.n + 1), (long long)d.n_bytes);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(out_offsets, d.offsets.ptr, (size_t)(d.n + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (d.n_bytes > 0) CU(ctx, cudaMemcpyAsync(out_bytes, d.data.ptr, (size_t)d.n_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return COLQ_OK;
}

colq_status colq_table_partition(colq_ctx* ctx, colq_table table, const int64_t* bounds, int n_ranks) {
    if (!ctx || !bounds) return COLQ_THROW_NULL;
    Table* t = get_table(ctx, table);
    if (!t) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown table handle %d", table);
    if (t->placement != COLQ_SHARDED) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "only a COLQ_SHARDED table has a partition");
    if (n_ranks != ctx->n_ranks) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "partition has %d ranks but the communicator has %d", n_ranks, ctx->n_ranks);
    if (bounds[0] != 0) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "partition bounds must start at 0");
    for (int r = 0; r < n_ranks; ++r) {
        if (bounds[r + 1] < bounds[r]) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "partition bounds must be non-decreasing");
        if (r > 0 && bounds[r] % 64 != 0 && bounds[r] != bounds[n_ranks])  // (a bound at the very end only closes empty shards)
            return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "partition bound %lld of rank %d is not a multiple of 64 rows (shards must be whole BitSet words)", (long long)bounds[r], r);
    }
    if (bounds[n_ranks] > (int64_t)INT32_MAX) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "global row count exceeds the int32 row-index range");
    if (bounds[ctx->rank] != t->row_base || bounds[ctx->rank + 1] - bounds[ctx->rank] != t->n_rows)
        return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "this rank's shard is rows [%lld, %lld) but the partition says [%lld, %lld)", (long long)t->row_base,
                    (long long)(t->row_base + t->n_rows), (long long)bounds[ctx->rank], (long long)bounds[ctx->rank + 1]);
    t->part.assign(bounds, bounds + n_ranks + 1);
    return COLQ_OK;
}

// the row count association targets are checked against: all ranks' rows when the keys are global
static int64_t target_rows(const Table& y, bool global) { return global ? y.global_rows() : y.n_rows; }

colq_status colq_associate_fk_global(colq_ctx* ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal, const int32_t* fk, int64_t n) {
    if (!ctx || (!fk && n > 0)) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    Table* Y = get_table(ctx, y);
    if (!Y) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown table handle %d", y);
    if (Y->placement == COLQ_SHARDED && ctx->n_ranks > 1 && Y->part.empty())
        return fail(ctx, COLQ_THROW_ILLEGAL_STATE, "declare the target table's partition first (colq_table_partition)");
    Column* f;
    ST(link_assoc(ctx, x, x_ordinal, y, y_ordinal, true, &f));
    if (n != f->n) { unlink_assoc(ctx, x, x_ordinal, y, y_ordinal); return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "association height %lld != table rows", (long long)n); }
    colq_status st = upload(ctx, f->data, fk, (size_t)n * 4, (size_t)round_up(n * 4 + 16, 16));
    if (st == COLQ_OK) st = check_fk_range(ctx, (const int32_t*)f->data.ptr, n, target_rows(ctx->tables[y], true));
    if (st != COLQ_OK) { unlink_assoc(ctx, x, x_ordinal, y, y_ordinal); return st; }
    ctx->tables[x].cols[x_ordinal].global_targets = true;
    return COLQ_OK;
}

colq_status colq_associate_csr_global(colq_ctx* ctx, colq_table x, int x_ordinal, colq_table y, int y_ordinal, const int64_t* offsets,
                                      const int32_t* targets, int64_t n, int64_t nnz) {
    if (!ctx || !offsets || (!targets && nnz > 0)) return COLQ_THROW_NULL;
    CU(ctx, cudaSetDevice(ctx->device));
    Table* Y = get_table(ctx, y);
    if (!Y) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown table handle %d", y);
    if (Y->placement == COLQ_SHARDED && ctx->n_ranks > 1 && Y->part.empty())
        return fail(ctx, COLQ_THROW_ILLEGAL_STATE, "declare the target table's partition first (colq_table_partition)");
    if (offsets[0] != 0 || offsets[n] != nnz) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "CSR offsets must start at 0 and end at nnz");
    const int64_t gy = Y->global_rows();
    Column* f;
    ST(link_assoc(ctx, x, x_ordinal, y, y_ordinal, false, &f));
    if (n != f->n) { unlink_assoc(ctx, x, x_ordinal, y, y_ordinal); return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "association height %lld != table rows", (long long)n); }
    f->nnz = nnz;
    colq_status st = upload(ctx, f->offsets, offsets, (size_t)(n + 1) * 8, (size_t)(n + 1) * 8 + 16);
    if (st == COLQ_OK) st = upload(ctx, f->targets, targets, (size_t)nnz * 4, (size_t)nnz * 4 + 16);
    // offsets order and target range against the GLOBAL row count, in one pass on the device like colq_associate_csr
    // (M/InMemoryTable.java:70-71)
    AssocStats a{};
    if (st == COLQ_OK) st = assoc_stats(ctx, (const int64_t*)f->offsets.ptr, (const int32_t*)f->targets.ptr, n, nnz, &a);
    if (st == COLQ_OK) st = check_csr(ctx, a, nnz, gy);
    if (st != COLQ_OK) { unlink_assoc(ctx, x, x_ordinal, y, y_ordinal); return st; }
    ctx->tables[x].cols[x_ordinal].global_targets = true;
    return COLQ_OK;
}

colq_status colq_table_destroy(colq_ctx* ctx, colq_table table) {
    if (!ctx) return COLQ_THROW_NULL;
    Table* t = get_table(ctx, table);
    if (!t) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown table handle %d", table);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto it = ctx->registry.begin(); it != ctx->registry.end();)
        it = (it->second == table) ? ctx->registry.erase(it) : std::next(it);
    // association columns of OTHER tables that point here become unset (their peer is gone)
    for (Table& o : ctx->tables)
        for (Column& c : o.cols)
            if (c.kind == COL_ASSOC && c.peer_table == table && &o != t) c = Column();
    t->cols.clear();
    t->part.clear();
    t->n_rows = 0;
    return COLQ_OK;
}

colq_status colq_table_size(const colq_ctx* ctx, colq_table table, int64_t* out_rows) {
    if (!ctx || !out_rows) return COLQ_THROW_NULL;
    if (table < 0 || (size_t)table >= ctx->tables.size()) return COLQ_THROW_ILLEGAL_ARG;
    *out_rows = ctx->tables[table].n_rows;
    return COLQ_OK;
}

colq_status colq_table_width(const colq_ctx* ctx, colq_table table, int* out_columns) {
    if (!ctx || !out_columns) return COLQ_THROW_NULL;
    if (table < 0 || (size_t)table >= ctx->tables.size()) return COLQ_THROW_ILLEGAL_ARG;
    *out_columns = (int)ctx->tables[table].cols.size();
    return COLQ_OK;
}

colq_status colq_query_create(colq_ctx* ctx, const char* table_name, colq_query** out_query) {
    if (!ctx || !table_name || !out_query) return COLQ_THROW_NULL;
    colq_query* q = new colq_query();
    if (const char* e = getenv("COLQ_COMPACT")) q->opt_fused_compact = atoi(e);  // experiment knob: default compaction kernel
    if (const char* e = getenv("COLQ_ROOT_FUSED")) q->opt_root_fused = atoi(e);   // experiment knob: default root plan (0, 1, 2)
    if (const char* e = getenv("COLQ_PIPELINE")) q->opt_pipeline = atoi(e);       // experiment knob: default COLQ_OPT_PIPELINE
    q->ctx = ctx;
    q->table_name = table_name;
    q->nodes.emplace_back();  // rootNode (DS/Query.java:22-25)
    ctx->queries.push_back(q);
    *out_query = q;
    return COLQ_OK;
}

colq_status colq_query_destroy(colq_query* q) {
    if (!q) return COLQ_OK;
    auto& live = q->ctx->queries;
    live.erase(std::remove(live.begin(), live.end(), q), live.end());
    if (q->ctx->chain_query == q) q->ctx->chain_query = nullptr;
    cudaSetDevice(q->ctx->device);
    cudaStreamSynchronize(q->ctx->stream);
    if (q->ev_start) cudaEventDestroy(q->ev_start);
    if (q->ev_stop) cudaEventDestroy(q->ev_stop);
    for (cudaEvent_t e : q->stage_ev) cudaEventDestroy(e);
    for (auto& pr : q->hot_ring) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    delete q;
    return COLQ_OK;
}

colq_status colq_query_child(colq_query* q, int parent_node, int ordinal, int* out_node) {
    if (!q || !out_node) return COLQ_THROW_NULL;
    if (parent_node < 0 || (size_t)parent_node >= q->nodes.size()) return fail(q->ctx, COLQ_THROW_ILLEGAL_ARG, "unknown query node %d", parent_node);
    for (auto& ch : q->nodes[parent_node].children)
        if (ch.first == ordinal)  // DS/Query.java:33-35
            return fail(q->ctx, COLQ_THROW_ILLEGAL_ARG, "A child already exists at ordinal %d", ordinal);
    q->nodes.emplace_back();
    int id = (int)q->nodes.size() - 1;
    q->nodes[parent_node].children.emplace_back(ordinal, id);
    *out_node = id;
    return COLQ_OK;
}

colq_status colq_query_criteria_i32_range(colq_query* q, int node, int ordinal, int32_t lo, int32_t hi) {
    if (!q) return COLQ_THROW_NULL;
    if (node < 0 || (size_t)node >= q->nodes.size()) return fail(q->ctx, COLQ_THROW_ILLEGAL_ARG, "unknown query node %d", node);
    Crit c;
    c.ordinal = ordinal; c.is_str = false; c.lo = lo; c.hi = hi;
    q->nodes[node].crit.push_back(std::move(c));
    return COLQ_OK;
}

colq_status colq_query_criteria_str(colq_query* q, int node, int ordinal, colq_str_op op, const uint8_t* needle, int32_t len) {
    if (!q || (!needle && len > 0)) return COLQ_THROW_NULL;
    colq_ctx* ctx = q->ctx;
    if (node < 0 || (size_t)node >= q->nodes.size()) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown query node %d", node);
    if ((int)op < 0 || (int)op > (int)COLQ_STR_ENDS_WITH) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown string operator %d", (int)op);
    if (len < 0 || len > ST_MAX_NEEDLE) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "needle length %d outside [0, %d]", len, ST_MAX_NEEDLE);
    CU(ctx, cudaSetDevice(ctx->device));
    Crit c;
    c.ordinal = ordinal; c.is_str = true; c.op = (int)op;
    c.needle.assign(needle, needle + len);
    ST(upload(ctx, c.needle_dev, needle, (size_t)len, (size_t)round_up(len + 16, 16)));
    q->nodes[node].crit.push_back(std::move(c));
    return COLQ_OK;
}

colq_status colq_query_criteria_str_accept(colq_query* q, int node, int ordinal, const uint64_t* accept_words, int64_t n_dict) {
    if (!q || (!accept_words && n_dict > 0)) return COLQ_THROW_NULL;
    colq_ctx* ctx = q->ctx;
    if (node < 0 || (size_t)node >= q->nodes.size()) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "unknown query node %d", node);
    if (n_dict < 0) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "negative dictionary size");
    CU(ctx, cudaSetDevice(ctx->device));
    Crit c;
    c.ordinal = ordinal; c.is_str = true; c.is_accept = true; c.accept_n = n_dict;
    const size_t words = (size_t)((n_dict + 63) / 64);
    ST(upload(ctx, c.accept_dev, accept_words, words * 8, (size_t)bitmap_alloc_words(n_dict) * 4));
    q->nodes[node].crit.push_back(std::move(c));
    return COLQ_OK;
}

colq_status colq_query_criteria_i32_accept(colq_query* q, int node, int ordinal, const uint64_t* accept_words, int64_t n_dict) {
    colq_status st = colq_query_criteria_str_accept(q, node, ordinal, accept_words, n_dict);
    if (st == COLQ_OK) q->nodes[node].crit.back().is_str = false;
    return st;
}

colq_status colq_query_criteria_bool(colq_query* q, int node, int ordinal, int accept_false, int accept_true) {
    if (!q) return COLQ_THROW_NULL;
    if (node < 0 || (size_t)node >= q->nodes.size()) return fail(q->ctx, COLQ_THROW_ILLEGAL_ARG, "unknown query node %d", node);
    Crit c;
    c.ordinal = ordinal; c.is_bool = true; c.accept_false = accept_false != 0; c.accept_true = accept_true != 0;
    q->nodes[node].crit.push_back(std::move(c));
    return COLQ_OK;
}

colq_status colq_query_set_option(colq_query* q, colq_option option, int value) {
    if (!q) return COLQ_THROW_NULL;
    switch (option) {
        case COLQ_OPT_LAZY_FK: q->opt_lazy = value; break;
        case COLQ_OPT_PROFILE: q->opt_profile = value; break;
        case COLQ_OPT_PEER_EXCHANGE: q->opt_peer = value; break;
        case COLQ_OPT_FUSED_COMPACT: q->opt_fused_compact = value; break;
        case COLQ_OPT_DEFER_CHAINS: q->opt_defer = value; break;
        case COLQ_OPT_PROMOTE: q->opt_promote = value; break;
        case COLQ_OPT_FUSED_GATHER: q->opt_fused_gather = value; break;
        case COLQ_OPT_TAIL_PUBLISH: q->opt_tail_publish = value; break;
        case COLQ_OPT_ROOT_FUSED: q->opt_root_fused = value; break;
        case COLQ_OPT_LAZY_GATHER_WAIT: q->opt_lazy_gather_wait = value; break;
        case COLQ_OPT_PIPELINE: q->opt_pipeline = value; break;
        default: return fail(q->ctx, COLQ_THROW_ILLEGAL_ARG, "unknown option %d", (int)option);
    }
    return COLQ_OK;
}

colq_status colq_execute_async(colq_ctx* ctx, colq_query* q) {
    if (!ctx || !q) return COLQ_THROW_NULL;
    if (q->ctx != ctx) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "query belongs to another context");
    return run_pipeline(q);
}

colq_status colq_fetch(colq_ctx* ctx, colq_query* q, uint64_t* out_bitmask, int64_t bitmask_capacity_words, int32_t* out_indices,
                       int64_t indices_capacity, int64_t* out_count, colq_timing* out_timing) {
    if (!ctx || !q) return COLQ_THROW_NULL;
    if (q->ctx != ctx) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "query belongs to another context");
    colq_status st = fetch_results(q, out_bitmask, bitmask_capacity_words, out_indices, indices_capacity, out_count, out_timing);
    if (st == RERUN_GROUP)
        return fail(ctx, COLQ_ERR_CAPACITY, "the result block of this rank is too small and the query must be re-run on every rank of the local communicator: use colq_fetch_group");
    return st;
}

colq_status colq_execute(colq_ctx* ctx, colq_query* q, uint64_t* out_bitmask, int64_t bitmask_capacity_words, int32_t* out_indices,
                         int64_t indices_capacity, int64_t* out_count, colq_timing* out_timing) {
    if (!ctx || !q) return COLQ_THROW_NULL;
    if (q->ctx != ctx) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "query belongs to another context");
    if (indices_capacity > q->want_idx_capacity && out_indices && !(ctx->n_ranks > 1))
        q->want_idx_capacity = std::min<int64_t>(indices_capacity, (int64_t)1 << 28);
    ST(run_pipeline(q));
    colq_status st = fetch_results(q, out_bitmask, bitmask_capacity_words, out_indices, indices_capacity, out_count, out_timing);
    if (st == RERUN_GROUP)
        return fail(ctx, COLQ_ERR_CAPACITY, "the result block of this rank is too small and the query must be re-run on every rank of the local communicator: use colq_execute_group + colq_fetch_group");
    return st;
}

colq_status colq_profile(const colq_query* q, colq_stage* out_stages, int capacity, int* out_n_stages) {
    if (!q || !out_n_stages) return COLQ_THROW_NULL;
    *out_n_stages = (int)q->stages.size();
    for (int i = 0; i < capacity && i < (int)q->stages.size(); ++i) out_stages[i] = q->stages[i];
    return COLQ_OK;
}

}  // extern "C"

// ---- result materialisation (K5) --------------------------------------------------------------------

namespace {

colq_status result_prologue(colq_ctx* ctx, colq_query* q, int ordinal, const Table** T, const Column** col) {
    if (q->ctx != ctx) return fail(ctx, COLQ_THROW_ILLEGAL_ARG, "query belongs to another context");
    if (q->local_count < 0) return fail(ctx, COLQ_THROW_ILLEGAL_STATE, "no fetched result: call colq_execute or colq_fetch first");
    CU(ctx, cudaSetDevice(ctx->device));
    *T = &ctx->tables[q->root_table];
    if (ordinal < 0 || (size_t)ordinal >= (*T)->cols.size()) return fail(ctx, COLQ_THROW_INDEX_OOB, "Index %d out of bounds for length %d", ordinal, (int)(*T)->cols.size());
    *col = &(*T)->cols[ordinal];
    if ((*col)->kind == COL_UNSET) return fail(ctx, COLQ_THROW_ILLEGAL_STATE, "column %d was never set", ordinal);
    return COLQ_OK;
}

template <typename T>
colq_status result_fixed(colq_ctx* ctx, colq_query* q, const Table& tb, const void* src, T* out, int64_t capacity, int64_t* out_count) {
    const int64_t n = q->local_count;
    if (out_count) *out_count = n;
    if (n == 0) return COLQ_OK;
    if (!out || capacity < n) return fail(ctx, COLQ_ERR_CAPACITY, "result capacity %lld < %lld rows", (long long)capacity, (long long)n);
    if (q->mat_a.bytes < (size_t)n * sizeof(T)) ST(dev_alloc(ctx, q->mat_a, (size_t)n * sizeof(T)));
    cudaStream_t s = ctx->stream;
    gather_values_kernel<T><<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, s>>>((const T*)src, q->d_idx, tb.row_base, n, (T*)q->mat_a.ptr);
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out, q->mat_a.ptr, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaStreamSynchronize(s));
    return COLQ_OK;
}

// lengths -> offsets -> payload copy; OutOffT is the host offset type (uint32 strings, int64 CSR)
template <typename OffT, typename ElemT, typename OutOffT>
colq_status result_var(colq_ctx* ctx, colq_query* q, const Table& tb, const OffT* src_off, const int32_t* codes, const ElemT* src,
                       OutOffT* out_off, int64_t off_capacity, ElemT* out, int64_t capacity, int64_t* out_count, int64_t* out_total,
                       uint64_t max_total) {
    const int64_t n = q->local_count;
    if (out_count) *out_count = n;
    cudaStream_t s = ctx->stream;
    if (q->mat_a.bytes < (size_t)(n + 1) * 8) ST(dev_alloc(ctx, q->mat_a, (size_t)(n + 1) * 8));
    GatherVarParams<OffT> P{src_off, codes, q->d_idx, tb.row_base, n, (u64*)q->mat_a.ptr};
    gather_var_lens_kernel<OffT><<<grid_for(std::max<int64_t>(n, 1), 256, ctx->sm_count, 8), 256, 0, s>>>(P);
    if (n > 0) scan_u64_inplace_kernel<<<1, 1024, 0, s>>>(P.out_off, n);
    CU(ctx, cudaGetLastError());
    u64 total = 0;
    CU(ctx, cudaMemcpyAsync(&total, P.out_off + n, 8, cudaMemcpyDeviceToHost, s));
    CU(ctx, cudaStreamSynchronize(s));
    if (out_total) *out_total = (int64_t)total;
    if (total > max_total) return fail(ctx, COLQ_ERR_CAPACITY, "result payload of %llu elements exceeds the offset range of this column type", (unsigned long long)total);
    if (!out_off || off_capacity < n + 1 || (total > 0 && (!out || capacity < (int64_t)total)))
        return fail(ctx, COLQ_ERR_CAPACITY, "result capacity too small: need %lld offsets and %llu payload elements", (long long)(n + 1), (unsigned long long)total);
    std::vector<u64> host_off((size_t)n + 1);
    CU(ctx, cudaMemcpyAsync(host_off.data(), P.out_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (total > 0) {
        if (q->mat_b.bytes < (size_t)total * sizeof(ElemT)) ST(dev_alloc(ctx, q->mat_b, (size_t)total * sizeof(ElemT)));
        gather_var_copy_kernel<OffT, ElemT><<<grid_for((n + 31) / 32 * 32, 256, ctx->sm_count, 8), 256, 0, s>>>(P, src, (ElemT*)q->mat_b.ptr);
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaMemcpyAsync(out, q->mat_b.ptr, (size_t)total * sizeof(ElemT), cudaMemcpyDeviceToHost, s));
    }
    CU(ctx, cudaStreamSynchronize(s));
    for (int64_t i = 0; i <= n; ++i) out_off[i] = (OutOffT)host_off[(size_t)i];
    return COLQ_OK;
}

}  // namespace

extern "C" {

colq_status colq_result_count(colq_ctx* ctx, colq_query* q, int64_t* out_rows) {
    if (!ctx || !q || !out_rows) return COLQ_THROW_NULL;
    if (q->local_count < 0) return fail(ctx, COLQ_THROW_ILLEGAL_STATE, "no fetched result: call colq_execute or colq_fetch first");
    *out_rows = q->local_count;
    return COLQ_OK;
}

colq_status colq_result_i32(colq_ctx* ctx, colq_query* q, int ordinal, int32_t* out_values, int64_t capacity, int64_t* out_count) {
    if (!ctx || !q) return COLQ_THROW_NULL;
    const Table* T = nullptr; const Column* c = nullptr;
    ST(result_prologue(ctx, q, ordinal, &T, &c));
    const bool to_one = c->kind == COL_ASSOC && c->forward && c->is_fk;
    if (c->kind != COL_I32 && !to_one)
        return fail(ctx, COLQ_FAILURE, "column %d is neither an integer column nor a stored to-one association column", ordinal);
    if (c->kind == COL_I32 && c->dict) {  // decode: codes at the matching rows -> distinct values
        const int64_t n = q->local_count;
        if (out_count) *out_count = n;
        if (n == 0) return COLQ_OK;
        if (!out_values || capacity < n) return fail(ctx, COLQ_ERR_CAPACITY, "result capacity %lld < %lld rows", (long long)capacity, (long long)n);
        if (q->mat_a.bytes < (size_t)n * 4) ST(dev_alloc(ctx, q->mat_a, (size_t)n * 4));
        gather_decode_kernel<<<grid_for(n, 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>((const int32_t*)c->data.ptr, (const int32_t*)c->dict->data.ptr,
                                                                                          q->d_idx, T->row_base, n, (int32_t*)q->mat_a.ptr);
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaMemcpyAsync(out_values, q->mat_a.ptr, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        return COLQ_OK;
    }
    return result_fixed<int32_t>(ctx, q, *T, c->data.ptr, out_values, capacity, out_count);
}

colq_status colq_result_bool(colq_ctx* ctx, colq_query* q, int ordinal, uint8_t* out_values, int64_t capacity, int64_t* out_count) {
    if (!ctx || !q) return COLQ_THROW_NULL;
    const Table* T = nullptr; const Column* c = nullptr;
    ST(result_prologue(ctx, q, ordinal, &T, &c));
    if (c->kind != COL_BOOL) return fail(ctx, COLQ_FAILURE, "column %d is not a boolean column", ordinal);
    return result_fixed<uint8_t>(ctx, q, *T, c->data.ptr, out_values, capacity, out_count);
}

colq_status colq_result_str(colq_ctx* ctx, colq_query* q, int ordinal, uint32_t* out_offsets, int64_t offsets_capacity, uint8_t* out_bytes,
                            int64_t bytes_capacity, int64_t* out_count, int64_t* out_n_bytes) {
    if (!ctx || !q) return COLQ_THROW_NULL;
    const Table* T = nullptr; const Column* c = nullptr;
    ST(result_prologue(ctx, q, ordinal, &T, &c));
    if (c->kind != COL_STR) return fail(ctx, COLQ_FAILURE, "column %d is not a string column", ordinal);
    const Column* payload = c->dict ? c->dict.get() : c;
    return result_var<u32, uint8_t, uint32_t>(ctx, q, *T, (const u32*)payload->offsets.ptr, c->dict ? (const int32_t*)c->data.ptr : nullptr,
                                              (const uint8_t*)payload->data.ptr, out_offsets, offsets_capacity, out_bytes, bytes_capacity,
                                              out_count, out_n_bytes, 0xfffffff0ull);
}

colq_status colq_result_csr(colq_ctx* ctx, colq_query* q, int ordinal, int64_t* out_offsets, int64_t offsets_capacity, int32_t* out_targets,
                            int64_t targets_capacity, int64_t* out_count, int64_t* out_nnz) {
    if (!ctx || !q) return COLQ_THROW_NULL;
    const Table* T = nullptr; const Column* c = nullptr;
    ST(result_prologue(ctx, q, ordinal, &T, &c));
    if (!(c->kind == COL_ASSOC && c->forward && !c->is_fk))
        return fail(ctx, COLQ_FAILURE, "column %d is not a stored to-many association column", ordinal);
    return result_var<int64_t, int32_t, int64_t>(ctx, q, *T, (const int64_t*)c->offsets.ptr, nullptr, (const int32_t*)c->targets.ptr,
                                                 out_offsets, offsets_capacity, out_targets, targets_capacity, out_count, out_nnz,
                                                 ~0ull >> 1);
}

colq_status colq_profile_hot(colq_query* q, colq_stage* out_stage, int* out_samples) {
    if (!q || !out_stage || !out_samples) return COLQ_THROW_NULL;
    colq_ctx* ctx = q->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    double sum = 0;
    for (size_t i = 0; i < q->hot_used; ++i) {
        float t = 0;
        CU(ctx, cudaEventElapsedTime(&t, q->hot_ring[i].first, q->hot_ring[i].second));
        sum += t;
    }
    *out_stage = q->hot_stage;
    out_stage->ms = q->hot_used ? sum / (double)q->hot_used : -1.0;
    *out_samples = (int)q->hot_used;
    q->hot_used = 0;
    return COLQ_OK;
}

colq_status colq_node_cardinalities(colq_ctx* ctx, const colq_query* q, int64_t* out, int capacity, int* out_n) {
    if (!ctx || !q || !out_n) return COLQ_THROW_NULL;
    if (!q->executed) return fail(ctx, COLQ_THROW_ILLEGAL_STATE, "query was never executed");
    CU(ctx, cudaSetDevice(ctx->device));
    *out_n = (int)q->xnodes.size();
    DevBuf acc;
    ST(dev_alloc(ctx, acc, 8));
    for (int i = 0; i < capacity && i < (int)q->xnodes.size(); ++i) {
        const XNode& x = q->xnodes[i];
        const int64_t n = ctx->tables[x.table].n_rows;
        if (x.fused && !x.bits) { out[i] = -1; continue; }
        if (x.all_ones || !x.bits) { out[i] = x.all_ones ? n : -1; continue; }
        CU(ctx, cudaMemsetAsync(acc.ptr, 0, 8, ctx->stream));
        popc_total_kernel<<<grid_for(bitmap_words(n), 256, ctx->sm_count, 8), 256, 0, ctx->stream>>>(x.bits, bitmap_words(n), (u64*)acc.ptr);
        u64 c = 0;
        CU(ctx, cudaMemcpyAsync(&c, acc.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        out[i] = (int64_t)c;
    }
    return COLQ_OK;
}

}  // extern "C"

#ifdef COLQ_RF_DEBUG
// debug builds only (not part of include/colq.h): the per-CTA phase timestamps of the last root_fused launch
extern "C" int colq_debug_rf_times(colq_query* q, uint64_t* out, int cap_ctas) {
    if (!q || !q->rf_dbg) return 0;
    cudaStreamSynchronize(q->ctx->stream);
    const int n = std::min(cap_ctas, q->rf_dbg_ctas);
    cudaMemcpy(out, q->rf_dbg, (size_t)n * 64, cudaMemcpyDeviceToHost);
    return n;
}
#endif
