// colq_kernels.cuh -- hand-written sm_100a kernels of the query hot path.
//
// Each kernel names the reference loop it replaces.  Citations are relative to the reference checkout:
//   E = data-system-serial-indices-arrays/src/main/java/dgroomes/data_system_serial_indices_arrays
//   M = data-model-in-memory/src/main/java/dgroomes/in_memory
//
// Everything here is HBM-bound integer / byte work: the design rules are coalesced 128-bit loads,
// TMA (cp.async.bulk) staging for the variable-length string column, many loads in flight per SM and
// grids sized from the SM count.  There is no GEMM-shaped work, hence no tcgen05.
//
// Bitmask layout everywhere: java.util.BitSet words, row i <-> bit (i & 63) of uint64 word (i >> 6).
// Kernels address the same memory as little-endian uint32 halves (row i <-> bit (i & 31) of u32 word (i >> 5)).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace colq {

typedef uint32_t u32;
typedef uint64_t u64;

constexpr u32 FULL_MASK = 0xffffffffu;
constexpr u32 ST_NO_TILE = 0xffffffffu;  // sentinel stage of the TMA rings: the producer has run out of tiles

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------

// streaming 128-bit load: read-only path, do not allocate in L1 (every input element is touched once)
__device__ __forceinline__ int4 ldg_stream_v4(const int32_t* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_v4u(const u32* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ bool bit_test(const u32* bits, int64_t i) { return (bits[i >> 5] >> (i & 31)) & 1u; }

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while the kernel in front of it in the stream is still draining, once every CTA of that kernel has executed
// pdl_launch_dependents() or exited; it must call pdl_wait() before it touches anything the earlier kernel wrote
// (pdl_wait returns when that kernel has completed and its writes are visible).  Both are no-ops in ordinary launches.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ u64 ld_volatile_u64(const u64* p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(u64* p, u64 v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// push epilogue: ExecutionContext.Node.filterParent for a to-one association (E/ExecutionContext.java:100-122,
// Association.One branch :114).  A matching child row sets the bit of its parent row in `reach`.
// Small parents (<= PUSH_SMEM_BITS rows, e.g. the 51-row states table) are accumulated in shared memory and
// flushed once per CTA so that same-address atomics do not serialise in L2.
// ---------------------------------------------------------------------------------------------

constexpr int PUSH_SMEM_BITS = 4096;
constexpr int PUSH_SMEM_WORDS = PUSH_SMEM_BITS / 32;

struct PushD {
    const int32_t* fk;  // forward to-one column on the CHILD table (child row -> parent row, -1 = None); null = no push
    u32* reach;         // parent-sized bitmask, zero-initialised
    int64_t n_parent;
    u32* oob;           // nullable: set to 1 when a walked target is outside the parent table (host-resident columns
                        // are range-checked lazily, on the rows a query walks, instead of at ingest)
};

__device__ __forceinline__ void push_init(const PushD& ps, u32* s_reach) {
    if (ps.fk != nullptr && ps.n_parent <= PUSH_SMEM_BITS) {
        for (int i = threadIdx.x; i < PUSH_SMEM_WORDS; i += blockDim.x) s_reach[i] = 0;
    }
}
__device__ __forceinline__ void push_row(const PushD& ps, u32* s_reach, int64_t row) {
    int32_t t = ps.fk[row];
    if (t < 0 || t >= ps.n_parent) {  // None, or (never via associateTo) out of range: vanishes at the AND
        if (t != -1 && ps.oob != nullptr) *ps.oob = 1u;
        return;
    }
    u32 m = 1u << (t & 31);
    if (ps.n_parent <= PUSH_SMEM_BITS) {
        if (!(s_reach[t >> 5] & m)) atomicOr(&s_reach[t >> 5], m);
    } else {
        if (!(ps.reach[t >> 5] & m)) atomicOr(&ps.reach[t >> 5], m);
    }
}
// call after a __syncthreads()
__device__ __forceinline__ void push_flush(const PushD& ps, const u32* s_reach) {
    if (ps.fk != nullptr && ps.n_parent <= PUSH_SMEM_BITS) {
        int words = (int)((ps.n_parent + 31) >> 5);
        for (int i = threadIdx.x; i < words; i += blockDim.x) {
            u32 w = s_reach[i];
            if (w) atomicOr(&ps.reach[i], w);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1  scan_rows: int32 range predicates + foreign-key chain gathers -> bitmask
//
// Replaces ExecutionContext.Node.filterSelf (E/ExecutionContext.java:79-94) over IntegerColumn.where
// (M/InMemoryColumn.java:53-56) and, fused into the same pass, the to-one branch of filterParent for every
// child reached through a forward foreign key (pull form: parent row r keeps its bit iff the child row fk[r]
// matches; equals the reference's push through the transposed reverse column, M/InMemoryTable.java:55-85).
//
// Mapping: one warp owns 512 consecutive rows per tile (4 x 128-bit loads per lane and column, all issued
// before first use), a 256-thread CTA owns 4096 rows.  Lane l of vector j holds rows base + 128 j + 4 l .. +3;
// its 4 result bits are merged into u32 words with an 8-lane shuffle butterfly and the warp stores its 16 words
// (64 B) with one coalesced instruction.
// ---------------------------------------------------------------------------------------------

constexpr int SR_THREADS = 256;
constexpr int SR_V = 4;                                  // 128-bit vectors per lane per column
constexpr int SR_WARP_ROWS = 128 * SR_V;                 // 512
constexpr int SR_BLOCK_ROWS = SR_WARP_ROWS * (SR_THREADS / 32);  // 4096
constexpr int SR_MAX_PRED = 2;
constexpr int SR_MAX_GATHER = 2;
constexpr int GATHER_MAX_DEPTH = 3;

struct IntPredD {
    const int32_t* col;
    int32_t lo;
    u32 span;  // row passes iff (u32)(v - lo) <= span, i.e. lo <= v <= lo + span (closed interval, no overflow)
    int32_t* promote;  // nullable: `col` is pinned HOST memory streamed over PCIe; every value read is also stored
                       // here so that the column is HBM-resident for the next query (first-touch promotion)
};

// A chain of to-one hops r0 -fk[0]-> r1 -fk[1]-> ... ending in a bitmask test.
struct GatherD {
    const int32_t* fk[GATHER_MAX_DEPTH];
    int64_t n[GATHER_MAX_DEPTH];  // n[d] = rows of the table that fk[d]'s values index into
    int depth;
    const u32* bits;  // final test; null = every row of the last table matches
    u32* oob;         // nullable: see PushD::oob
};

struct ScanRowsParams {
    int64_t n;
    const u32* in_bits;  // nullable: AND-in (bits of this node computed so far)
    u32* out_bits;       // nullable when only the push epilogue consumes the result
    IntPredD pred[SR_MAX_PRED];
    GatherD gather[SR_MAX_GATHER];
    PushD push;
    // LIST instantiations (the root's scan in front of root_finish_kernel): every warp also leaves the surviving rows of
    // its 512-row chunk, in row order, in lists[chunk][list_cap] and their number in ucount[chunk] (> list_cap: overflow)
    u32* lists;
    u32* ucount;
    int list_cap;
};

__device__ __forceinline__ bool gather_eval(const GatherD& g, int64_t r, int level) {
    for (int d = level; d < g.depth; ++d) {
        int32_t t = g.fk[d][r];
        if (t < 0 || t >= g.n[d]) {  // Association.None, or a target that vanishes at the AND
            if (t != -1 && g.oob != nullptr) *g.oob = 1u;
            return false;
        }
        r = t;
    }
    return g.bits == nullptr ? true : bit_test(g.bits, r);
}

__device__ __forceinline__ u32 range4(const int4& v, int32_t lo, u32 span) {
    u32 m = 0;
    m |= ((u32)(v.x - lo) <= span) ? 1u : 0u;
    m |= ((u32)(v.y - lo) <= span) ? 2u : 0u;
    m |= ((u32)(v.z - lo) <= span) ? 4u : 0u;
    m |= ((u32)(v.w - lo) <= span) ? 8u : 0u;
    return m;
}

// NP predicates, NG gather chains.  EAGER: the first hop of every chain is loaded with coalesced 128-bit loads
// for all rows (right when no selective predicate precedes it); otherwise chains are walked lazily, only for rows
// that survived the predicates (the 0.16 %-selective population filter skips almost every FK sector).
template <int NP, int NG, bool EAGER, bool LIST = false>
__global__ void __launch_bounds__(SR_THREADS) scan_rows_kernel(const ScanRowsParams P) {
    __shared__ u32 s_reach[PUSH_SMEM_WORDS];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t wbase = (int64_t)blockIdx.x * SR_BLOCK_ROWS + (int64_t)warp * SR_WARP_ROWS;
    const bool do_push = P.push.fk != nullptr;
    if (do_push) {
        push_init(P.push, s_reach);
        __syncthreads();
    }

    if (wbase < P.n) {
        u32 nib[SR_V];
        int4 fkv[NG > 0 && EAGER ? NG : 1][SR_V];
        const bool full = wbase + SR_WARP_ROWS <= P.n;
        if (full) {
            // ---- fast path: every vector is inside the column; issue all loads first
            int4 v[NP > 0 ? NP : 1][SR_V];
#pragma unroll
            for (int p = 0; p < NP; ++p)
#pragma unroll
                for (int j = 0; j < SR_V; ++j) v[p][j] = ldg_stream_v4(P.pred[p].col + wbase + j * 128 + lane * 4);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if (P.pred[p].promote != nullptr) {
#pragma unroll
                    for (int j = 0; j < SR_V; ++j)
                        *reinterpret_cast<int4*>(P.pred[p].promote + wbase + j * 128 + lane * 4) = v[p][j];
                }
            }
            if (EAGER) {
#pragma unroll
                for (int g = 0; g < NG; ++g)
#pragma unroll
                    for (int j = 0; j < SR_V; ++j)
                        fkv[g][j] = ldg_stream_v4(P.gather[g].fk[0] + wbase + j * 128 + lane * 4);
            }
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 m = 0xFu;
#pragma unroll
                for (int p = 0; p < NP; ++p) m &= range4(v[p][j], P.pred[p].lo, P.pred[p].span);
                nib[j] = m;
            }
        } else {
            // ---- tail warp: scalar guarded loads
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 m = 0;
                int32_t f[NG > 0 ? NG : 1][4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int64_t r = wbase + j * 128 + lane * 4 + e;
                    bool ok = r < P.n;
#pragma unroll
                    for (int p = 0; p < NP; ++p) {
                        if (r < P.n) {  // (every row is read even when an earlier predicate failed: promotion copies it)
                            int32_t x = P.pred[p].col[r];
                            if (P.pred[p].promote != nullptr) P.pred[p].promote[r] = x;
                            ok = ok && (u32)(x - P.pred[p].lo) <= P.pred[p].span;
                        }
                    }
                    if (EAGER) {
#pragma unroll
                        for (int g = 0; g < NG; ++g) f[g][e] = (r < P.n) ? P.gather[g].fk[0][r] : -1;
                    }
                    m |= ok ? (1u << e) : 0u;
                }
                if (EAGER) {
#pragma unroll
                    for (int g = 0; g < NG; ++g) fkv[g][j] = make_int4(f[g][0], f[g][1], f[g][2], f[g][3]);
                }
                nib[j] = m;
            }
        }

        // ---- AND-in the bits this node already has
        if (P.in_bits != nullptr) {
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 w = P.in_bits[((wbase + j * 128) >> 5) + (lane >> 3)];
                nib[j] &= (w >> ((lane & 7) * 4)) & 0xFu;
            }
        }

        // ---- association hops through forward foreign keys
#pragma unroll
        for (int g = 0; g < NG; ++g) {
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 m = nib[j];
                if (m == 0) continue;
                const int64_t r0 = wbase + j * 128 + lane * 4;
                if (EAGER) {
                    const int32_t f[4] = {fkv[g][j].x, fkv[g][j].y, fkv[g][j].z, fkv[g][j].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (m & (1u << e)) {
                            bool ok = f[e] >= 0 && f[e] < P.gather[g].n[0];
                            if (!ok && f[e] != -1 && P.gather[g].oob != nullptr) *P.gather[g].oob = 1u;
                            ok = ok && gather_eval(P.gather[g], f[e], 1);
                            if (!ok) m &= ~(1u << e);
                        }
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (m & (1u << e)) {
                            if (!gather_eval(P.gather[g], r0 + e, 0)) m &= ~(1u << e);
                        }
                    }
                }
                nib[j] = m;
            }
        }

        // ---- push epilogue (this node is the child of a reverse-side hop)
        if (do_push) {
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 m = nib[j];
                while (m) {
                    int e = __ffs(m) - 1;
                    m &= m - 1;
                    push_row(P.push, s_reach, wbase + j * 128 + lane * 4 + e);
                }
            }
        }

        // ---- pack nibbles into u32 words and store 64 B per warp
        if (P.out_bits != nullptr) {
            u32 y[SR_V];
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 x = nib[j] << ((lane & 7) * 4);
                x |= __shfl_xor_sync(FULL_MASK, x, 1);
                x |= __shfl_xor_sync(FULL_MASK, x, 2);
                x |= __shfl_xor_sync(FULL_MASK, x, 4);
                y[j] = __shfl_sync(FULL_MASK, x, (lane & 3) * 8);  // lane gets word (lane & 3) of vector j
            }
            u32 out = y[0];
#pragma unroll
            for (int j = 1; j < SR_V; ++j) out = ((lane >> 2) == j) ? y[j] : out;
            if (lane < 4 * SR_V) P.out_bits[(wbase >> 5) + lane] = out;
        }

        // ---- LIST: the chunk's survivors, in row order (vector j, then lane, then element), for root_finish_kernel
        if (LIST) {
            const int64_t chunk = wbase / SR_WARP_ROWS;
            u32* my_list = P.lists + (size_t)chunk * P.list_cap;
            u32 cnt = 0;
            if (__ballot_sync(FULL_MASK, (nib[0] | nib[1] | nib[2] | nib[3]) != 0) != 0) {
#pragma unroll
                for (int j = 0; j < SR_V; ++j) {
                    const u32 m = nib[j];
                    const u32 mine = __popc(m);
                    if (__ballot_sync(FULL_MASK, mine != 0) == 0) continue;
                    u32 incl = mine;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const u32 t = __shfl_up_sync(FULL_MASK, incl, d);
                        if (lane >= d) incl += t;
                    }
                    const u32 tot = __shfl_sync(FULL_MASK, incl, 31);
                    if (cnt + tot <= (u32)P.list_cap) {
                        u32 pos = cnt + incl - mine, mm = m;
                        while (mm) {
                            const int e = __ffs(mm) - 1;
                            mm &= mm - 1;
                            my_list[pos++] = (u32)(wbase + j * 128 + lane * 4 + e);
                        }
                    }
                    cnt += tot;  // keeps counting past the cap: cnt > list_cap marks the overflow
                }
            }
            if (lane == 0) P.ucount[chunk] = cnt;
        }
    }
    // LIST launches sit between the string scan (whose programmatic dependent they are: they read nothing of it) and
    // root_finish_kernel, which does: this launch must not COMPLETE before that scan has, so that the finish kernel's
    // own dependency wait covers both.  A no-op in ordinary launches.
    if (LIST) pdl_wait();

    if (do_push) {
        __syncthreads();
        push_flush(P.push, s_reach);
    }
}

// ---------------------------------------------------------------------------------------------
// K1-TMA  scan_rows_tma: the single-predicate scan with cp.async.bulk staging (north_star: "128-bit vectorised coalesced
// loads and TMA/shared-memory staging").  A/B variant of scan_rows<1, 0>, selected with COLQ_SCAN_ROWS_TMA=1: persistent
// CTAs, one producer lane claims 4096-row tiles from the device-wide counter and copies each (16 KB) into a ring of
// SRT_STAGES shared-memory slots with one bulk copy; eight consumer warps read their 512 rows with conflict-free LDS.128
// and do the same compare / pack / 64-byte mask store as scan_rows.  Measured against the LDG.128 kernel in
// profiles/r02_scan_rows_tma_ab.txt; the faster one is the default.
// ---------------------------------------------------------------------------------------------
constexpr int SRT_STAGES = 4;
constexpr int SRT_THREADS = SR_THREADS + 32;
constexpr int SRT_TILE_BYTES = SR_BLOCK_ROWS * 4;  // 16 KB

struct ScanRowsTmaParams {
    int64_t n;
    const int32_t* col;
    int32_t lo;
    u32 span;
    u32* out_bits;
    u32* tile_counter;  // [0] next unclaimed tile, [1] finished CTAs (zero between launches)
};

__global__ void __launch_bounds__(SRT_THREADS) scan_rows_tma_kernel(const ScanRowsTmaParams P) {
    extern __shared__ __align__(128) uint8_t srt_smem[];
    __shared__ u64 s_full[SRT_STAGES], s_empty[SRT_STAGES];
    __shared__ u32 s_tile[SRT_STAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 n_tiles = (u32)((P.n + SR_BLOCK_ROWS - 1) / SR_BLOCK_ROWS);
    if (tid == 0) {
        for (int s = 0; s < SRT_STAGES; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], SR_THREADS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == SR_THREADS / 32) {
        if (lane == 0) {
            u32 batch = atomicAdd(P.tile_counter, 2u), batch_next = atomicAdd(P.tile_counter, 2u), j = 0;
            int s = 0;
            u32 round = 0;
            while (true) {
                const u32 cur = batch + j;
                if (round > 0) mbar_wait(&s_empty[s], (round - 1) & 1);
                if (cur >= n_tiles) {
                    s_tile[s] = ST_NO_TILE;
                    mbar_arrive(&s_full[s]);
                    break;
                }
                s_tile[s] = cur;
                const int64_t r0 = (int64_t)cur * SR_BLOCK_ROWS;
                const int64_t rows = (P.n - r0) < SR_BLOCK_ROWS ? (P.n - r0) : SR_BLOCK_ROWS;
                const u32 bytes = (u32)((rows * 4 + 15) & ~(int64_t)15);
                mbar_arrive_expect_tx(&s_full[s], bytes);
                tma_bulk_g2s(srt_smem + (size_t)s * SRT_TILE_BYTES, P.col + r0, bytes, &s_full[s]);
                if (++j == 2) {
                    j = 0;
                    batch = batch_next;
                    batch_next = atomicAdd(P.tile_counter, 2u);
                }
                if (++s == SRT_STAGES) {
                    s = 0;
                    ++round;
                }
            }
        }
    } else {
        u32 s = 0, parity = 0;
        while (true) {
            mbar_wait(&s_full[s], parity);
            const u32 tile = s_tile[s];
            if (tile == ST_NO_TILE) break;
            const int64_t wbase = (int64_t)tile * SR_BLOCK_ROWS + (int64_t)warp * SR_WARP_ROWS;
            const u32 sbase = smem_u32(srt_smem) + s * SRT_TILE_BYTES + (u32)warp * SR_WARP_ROWS * 4;
            u32 nib[SR_V];
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                int4 v;
                asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sbase + (j * 128 + lane * 4) * 4));
                u32 m = range4(v, P.lo, P.span);
                const int64_t r = wbase + j * 128 + lane * 4;
                if (r + 3 >= P.n) {  // the table's last rows
                    u32 valid = 0;
#pragma unroll
                    for (int e = 0; e < 4; ++e) valid |= (r + e < P.n) ? (1u << e) : 0u;
                    m &= valid;
                }
                nib[j] = m;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);
            if (wbase < P.n) {
                u32 y[SR_V];
#pragma unroll
                for (int j = 0; j < SR_V; ++j) {
                    u32 x = nib[j] << ((lane & 7) * 4);
                    x |= __shfl_xor_sync(FULL_MASK, x, 1);
                    x |= __shfl_xor_sync(FULL_MASK, x, 2);
                    x |= __shfl_xor_sync(FULL_MASK, x, 4);
                    y[j] = __shfl_sync(FULL_MASK, x, (lane & 3) * 8);
                }
                u32 out = y[0];
#pragma unroll
                for (int j = 1; j < SR_V; ++j) out = ((lane >> 2) == j) ? y[j] : out;
                if (lane < 4 * SR_V) P.out_bits[(wbase >> 5) + lane] = out;
            }
            if (++s == SRT_STAGES) {
                s = 0;
                parity ^= 1;
            }
        }
    }
    __syncthreads();
    if (tid == 0 && atomicAdd(P.tile_counter + 1, 1u) == gridDim.x - 1) {
        P.tile_counter[0] = 0;
        P.tile_counter[1] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Peer-memory exchange over NVLink (multi-GPU, one process per GPU).
//
// Every rank owns a MAILBOX in its HBM that all peers have mapped through CUDA IPC.  A collective is: store my
// contribution straight into every peer's mailbox slot [parity][my rank] (plain st.global over NVLink), fence, then
// publish an epoch flag with st.release.sys; the consumer spins on the flags in its OWN memory (ld.acquire.sys) and
// reads the slots locally.  Slots are double-buffered by epoch parity: a rank can only be one exchange ahead of any
// peer (it needs that peer's flag to finish the current one), so a slot is never overwritten while it is still read.
// These replace ncclAllGather on the data path: ~5 us instead of ~25-70 us for the 8-byte state mask.
// A spin that exceeds PEER_TIMEOUT_NS sets *status and gives up (the host reports COLQ_ERR_DEVICE) -- no hangs.
// ---------------------------------------------------------------------------------------------
constexpr int MAX_RANKS = 64;
constexpr int MASK_SLOT_BYTES = 4096;
constexpr int MASK_WORDS_MAX = (MASK_SLOT_BYTES - 16) / 8;  // the mask travels as flag-in-data words: 8 bytes per 32 rows
constexpr size_t PEER_GATHER_AREA_OFFSET = (size_t)2 * MAX_RANKS * MASK_SLOT_BYTES;
constexpr int GATHER_SLOT_HEADER = 256;  // [u64 {epoch | count}] + pad, keeps the index words 256-byte aligned
constexpr unsigned long long PEER_TIMEOUT_NS = 4000000000ull;

__device__ __forceinline__ u64 ld_acquire_sys(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(u64* p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 global_timer_ns() {
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool peer_wait(const u64* flag, u64 epoch, u32* status) {
    const u64 t0 = global_timer_ns();
    while (ld_acquire_sys(flag) != epoch) {
        if (global_timer_ns() - t0 > PEER_TIMEOUT_NS) {
            atomicExch(status, 1u);
            return false;
        }
    }
    return true;
}

// Flag-in-data words ("LL", as in NCCL's low-latency protocol): one naturally atomic 8-byte store carries 32 bits of payload
// and the low 32 bits of the exchange's epoch, so the small-mask exchange needs NO fence and NO separate flag -- a
// system-scope fence behind stores into eight peers' memory cost ~5 us on the tail of the string scan, every step, on every
// rank (r02: scan_str+publish 78 us vs 68 us for the same shard without the publish).  The receiver polls the word itself.
__device__ __forceinline__ void ll_store(u64* p, u32 data, u64 epoch) {
    const u64 v = ((u64)(u32)epoch << 32) | data;
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u32 ll_load(const u64* p, u64 epoch, u32* status) {
    const u64 t0 = global_timer_ns();
    while (true) {
        u64 v;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if ((u32)(v >> 32) == (u32)epoch) return (u32)v;
        if (global_timer_ns() - t0 > PEER_TIMEOUT_NS) {
            atomicExch(status, 1u);
            return 0u;
        }
    }
}

struct PeerMaskParams {
    u32* reach;            // in: this rank's mask; out: OR over all ranks
    int n_words;           // <= MASK_WORDS_MAX
    int n_ranks, rank;
    uint8_t* const* peers; // device array: mailbox base of every rank (own entry = local pointer)
    u64 epoch;
    u32* status;
};

// OR-allreduce of a small replicated-table mask -- the collective half of a sharded -> replicated filterParent hop --
// split in two so that independent work can run between the halves (the planner puts the root table's predicate scan
// there): PUBLISH stores this rank's words into every peer's mailbox and raises the flags; COLLECT waits for all
// ranks' flags in local memory and ORs the slots.  COLLECT also runs as the prologue of a single-block csr_pull.
__global__ void __launch_bounds__(256) peer_mask_publish_kernel(const PeerMaskParams P) {
    const size_t area = (size_t)(P.epoch & 1) * MAX_RANKS * MASK_SLOT_BYTES;
    for (int r = 0; r < P.n_ranks; ++r) {
        u64* dst = reinterpret_cast<u64*>(P.peers[r] + area + (size_t)P.rank * MASK_SLOT_BYTES + 16);
        for (int w = threadIdx.x; w < P.n_words; w += blockDim.x) ll_store(dst + w, P.reach[w], P.epoch);
    }
}

// PUBLISH from the tail of the kernel that produced the mask (scan_str / scan_codes push epilogue): every CTA has
// flushed its bits with atomicOr; the LAST CTA to arrive (device-wide counter, re-armed by that CTA) reads the finished
// mask from L2 and publishes it.  Saves the separate one-block launch.  Call from all threads, after push_flush.
__device__ __forceinline__ void peer_mask_publish_tail(const PeerMaskParams& P, u32* done) {
    __shared__ u32 s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(done, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *done = 0;
    __threadfence();
    const size_t area = (size_t)(P.epoch & 1) * MAX_RANKS * MASK_SLOT_BYTES;
    // one thread per (rank, word): all stores of the exchange leave in one wave
    for (int i = threadIdx.x; i < P.n_ranks * P.n_words; i += blockDim.x) {
        const int r = i / P.n_words, w = i - r * P.n_words;
        ll_store(reinterpret_cast<u64*>(P.peers[r] + area + (size_t)P.rank * MASK_SLOT_BYTES + 16) + w, __ldcg(P.reach + w), P.epoch);
    }
}

// block-wide; ends with a __syncthreads()
__device__ __forceinline__ void peer_mask_collect(const PeerMaskParams& P) {
    const size_t area = (size_t)(P.epoch & 1) * MAX_RANKS * MASK_SLOT_BYTES;
    const uint8_t* mine = P.peers[P.rank] + area;
    for (int w = threadIdx.x; w < P.n_words; w += blockDim.x) P.reach[w] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < P.n_words * P.n_ranks; i += blockDim.x) {  // one thread per (rank, word): all polls in flight at once
        const int r = i / P.n_words, w = i - r * P.n_words;
        const u32 v = ll_load(reinterpret_cast<const u64*>(mine + (size_t)r * MASK_SLOT_BYTES + 16) + w, P.epoch, P.status);
        if (v) atomicOr(&P.reach[w], v);
    }
    __threadfence_block();
    __syncthreads();
}

__global__ void __launch_bounds__(256) peer_mask_collect_kernel(const PeerMaskParams P) { peer_mask_collect(P); }

// ---------------------------------------------------------------------------------------------
// Cross-shard association hops (SURVEY.md 8e "not supported" / 8f4: sharded <-> sharded hops whose targets leave the
// rank's shard; ExecutionContext.Node.filterParent, E/ExecutionContext.java:100-122, for ANY table pair).
//
// The association column holds GLOBAL row indices of the sharded target table.  Every rank keeps a bitmap over the
// target's GLOBAL rows in the peer-mapped HEAP behind its mailbox (same offset on every rank, double-buffered by the
// parity of the execution so that a rank one step ahead never overwrites what a peer still reads):
//   pull (parent holds the key):  the child's local bits are ALL-GATHERED into every rank's global bitmap
//        (peer_bits_allgather: plain stores over NVLink into each peer's copy, then the epoch flags of the mask
//        mailbox), and the parent's gather / csr_pull kernels test bit [global key] exactly as they test a local one;
//   push (child holds the key):   the child's kernels set bits in the rank's OWN global-sized reach bitmap with the
//        usual local atomics; peer_bits_reduce then ORs, for this rank's slice only, the copies of all ranks (remote
//        loads over NVLink) into the parent's local reach mask -- an OR-reduce-scatter without remote atomics.
// Partition bounds are multiples of 64 rows (except the last), so slices are whole BitSet words and never share one.
// ---------------------------------------------------------------------------------------------
struct PeerBitsParams {
    const u32* src;        // allgather: this rank's local bits
    u32* dst;              // reduce: this rank's local reach mask
    size_t heap_off;       // byte offset of the global bitmap inside every rank's mailbox allocation
    int64_t word_base;     // first u32 word of this rank's slice in the global bitmap
    int64_t n_words;       // u32 words of this rank's slice
    int n_ranks, rank;
    int n_src;             // reduce: how many ranks' copies are OR-ed (n_ranks; 1 = only this rank's own: replicated child)
    uint8_t* const* peers;
    u64 epoch;
    u32* status;
    u32* done;
};

__device__ __forceinline__ u32 ld_volatile_u32(const u32* p) {
    u32 v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) peer_bits_allgather_kernel(const PeerBitsParams P) {
    __shared__ u32 s_last;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int r = 0; r < P.n_ranks; ++r) {
        u32* dst = reinterpret_cast<u32*>(P.peers[r] + P.heap_off) + P.word_base;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n_words; i += stride) dst[i] = __ldcg(P.src + i);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(P.done, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *P.done = 0;
    __threadfence_system();
    const size_t area = (size_t)(P.epoch & 1) * MAX_RANKS * MASK_SLOT_BYTES;
    if ((int)threadIdx.x < P.n_ranks)
        st_release_sys(reinterpret_cast<u64*>(P.peers[threadIdx.x] + area + (size_t)P.rank * MASK_SLOT_BYTES), P.epoch);
    // the launch ends only when every rank's slice has arrived here: the consumers are ordinary later launches
    const uint8_t* mine = P.peers[P.rank] + area;
    if ((int)threadIdx.x < P.n_ranks)
        peer_wait(reinterpret_cast<const u64*>(mine + (size_t)threadIdx.x * MASK_SLOT_BYTES), P.epoch, P.status);
}

__global__ void __launch_bounds__(256) peer_bits_reduce_kernel(const PeerBitsParams P) {
    const size_t area = (size_t)(P.epoch & 1) * MAX_RANKS * MASK_SLOT_BYTES;
    if (P.n_src > 1) {
        // every push into this rank's copy was made by earlier launches of this stream: tell the peers it is complete,
        // then wait until theirs are
        if (blockIdx.x == 0 && (int)threadIdx.x < P.n_ranks) {
            __threadfence_system();
            st_release_sys(reinterpret_cast<u64*>(P.peers[threadIdx.x] + area + (size_t)P.rank * MASK_SLOT_BYTES), P.epoch);
        }
        const uint8_t* mine = P.peers[P.rank] + area;
        if ((int)threadIdx.x < P.n_ranks)
            peer_wait(reinterpret_cast<const u64*>(mine + (size_t)threadIdx.x * MASK_SLOT_BYTES), P.epoch, P.status);
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n_words; i += stride) {
        u32 v = 0;
        if (P.n_src > 1) {
            for (int r = 0; r < P.n_ranks; ++r) v |= ld_volatile_u32(reinterpret_cast<const u32*>(P.peers[r] + P.heap_off) + P.word_base + i);
        } else {
            v = __ldcg(reinterpret_cast<const u32*>(P.peers[P.rank] + P.heap_off) + P.word_base + i);
        }
        P.dst[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// K2'  scan_codes: string predicate over a DICTIONARY-ENCODED column -> bitmask
//
// Replaces ExecutionContext.Node.filterSelf (E/ExecutionContext.java:79-94) over StringColumn.where
// (M/InMemoryColumn.java:71-74) when the column is stored as int32 dictionary codes.  The predicate was evaluated
// once per DISTINCT value -- by scan_str over the dictionary, or on the host for an opaque lambda -- into an
// n_dict-bit accept mask; the row scan streams 4 bytes per row and tests bit `code`.  The mask is staged in shared
// memory (random 4-byte gathers: ~3 bank-conflict cycles per warp instead of up to 16 L1 wavefronts); dictionaries
// beyond SC_SMEM_MASK_WORDS * 32 entries are looked up through L1/L2.  When the staged mask has at most SC_FEW bits
// set -- an equality predicate accepts exactly ONE dictionary entry -- the CTA extracts those codes once and the row
// test is a handful of register compares with no shared-memory traffic at all (one compare per row for the single
// accepted code of an equality).  Same row mapping as scan_rows, SC_ITER consecutive 4096-row tiles per CTA so that
// the mask staging is amortised.
// ---------------------------------------------------------------------------------------------

constexpr int SC_ITER = 8;
constexpr int SC_BLOCK_ROWS = SR_BLOCK_ROWS * SC_ITER;  // 32768 rows per CTA
constexpr int SC_SMEM_MASK_WORDS = 8192;                // 32 KB: dictionaries of up to 262144 entries
constexpr int SC_FEW = 4;                               // accepted codes that are tested by register compares

struct ScanCodesParams {
    int64_t n;
    const int32_t* codes;
    int32_t* promote;     // nullable: first-touch promotion of host-resident codes (see IntPredD::promote)
    const u32* accept;    // n_dict-bit mask, allocation padded to whole 16-byte lines
    u32 n_dict;
    int mask_words;       // words staged in shared memory (0: look the mask up in global memory)
    const u32* in_bits;
    u32* out_bits;
    PushD push;
    PeerMaskParams pub;   // pub.n_words > 0: the last CTA publishes push.reach to the peers (multi-GPU mask exchange)
    u32* pub_done;
};

template <bool SMEM>
__global__ void __launch_bounds__(SR_THREADS) scan_codes_kernel(const ScanCodesParams P) {
    extern __shared__ __align__(16) u32 s_dyn[];  // [PUSH_SMEM_WORDS reach | mask_words accept]
    u32* s_reach = s_dyn;
    u32* s_accept = s_dyn + PUSH_SMEM_WORDS;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool do_push = P.push.fk != nullptr;
    if (do_push) push_init(P.push, s_reach);
    if (SMEM) {
        for (int i = threadIdx.x; i < P.mask_words; i += SR_THREADS) s_accept[i] = __ldg(P.accept + i);
    }
    __shared__ u32 s_few[SC_FEW];
    __shared__ u32 s_nfew;
    if (threadIdx.x == 0) s_nfew = 0;
    __syncthreads();
    if (SMEM) {
        // how many dictionary entries does the predicate accept?  (bits past n_dict are zero by construction)
        for (int i = threadIdx.x; i < P.mask_words; i += SR_THREADS) {
            u32 w = s_accept[i];
            while (w) {
                const int b = __ffs(w) - 1;
                w &= w - 1;
                const u32 slot = atomicAdd(&s_nfew, 1u);
                if (slot < SC_FEW) s_few[slot] = ((u32)i << 5) + b;
            }
        }
        __syncthreads();
    }
    const u32 n_few = SMEM ? s_nfew : SC_FEW + 1;
    const bool few = n_few <= SC_FEW;
    u32 k[SC_FEW];
#pragma unroll
    for (int i = 0; i < SC_FEW; ++i) k[i] = (few && i < (int)n_few) ? s_few[i] : 0xffffffffu;  // codes are >= 0: never equal
    const u32 span = P.n_dict - 1u;
    auto test = [&](int32_t c) -> bool {
        if (few) {
            bool m = false;
#pragma unroll
            for (int i = 0; i < SC_FEW; ++i) m = m || ((u32)c == k[i]);
            return m;
        }
        if ((u32)c > span) return false;  // a code outside the dictionary matches nothing
        const u32 w = SMEM ? s_accept[(u32)c >> 5] : __ldg(P.accept + ((u32)c >> 5));
        return (w >> ((u32)c & 31)) & 1u;
    };

    // SC_ITER consecutive 4096-row tiles per CTA amortise the prologue above (measured on B200: better than both one
    // tile per CTA and a persistent grid)
#pragma unroll 1
    for (int it = 0; it < SC_ITER; ++it) {
        const int64_t wbase = ((int64_t)blockIdx.x * SC_ITER + it) * SR_BLOCK_ROWS + (int64_t)warp * SR_WARP_ROWS;
        if (wbase >= P.n) break;
        u32 nib[SR_V];
        if (wbase + SR_WARP_ROWS <= P.n) {
            int4 v[SR_V];
#pragma unroll
            for (int j = 0; j < SR_V; ++j) v[j] = ldg_stream_v4(P.codes + wbase + j * 128 + lane * 4);
            if (P.promote != nullptr) {
#pragma unroll
                for (int j = 0; j < SR_V; ++j) *reinterpret_cast<int4*>(P.promote + wbase + j * 128 + lane * 4) = v[j];
            }
            if (n_few <= 1) {  // "X"::equals accepts one dictionary entry: a single compare per row
#pragma unroll
                for (int j = 0; j < SR_V; ++j)
                    nib[j] = ((u32)v[j].x == k[0] ? 1u : 0u) | ((u32)v[j].y == k[0] ? 2u : 0u) | ((u32)v[j].z == k[0] ? 4u : 0u) |
                             ((u32)v[j].w == k[0] ? 8u : 0u);
            } else {
#pragma unroll
                for (int j = 0; j < SR_V; ++j)
                    nib[j] = (test(v[j].x) ? 1u : 0u) | (test(v[j].y) ? 2u : 0u) | (test(v[j].z) ? 4u : 0u) | (test(v[j].w) ? 8u : 0u);
            }
        } else {
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 m = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t r = wbase + j * 128 + lane * 4 + e;
                    if (r < P.n) {
                        const int32_t c = P.codes[r];
                        if (P.promote != nullptr) P.promote[r] = c;
                        m |= test(c) ? (1u << e) : 0u;
                    }
                }
                nib[j] = m;
            }
        }
        if (P.in_bits != nullptr) {
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 w = P.in_bits[((wbase + j * 128) >> 5) + (lane >> 3)];
                nib[j] &= (w >> ((lane & 7) * 4)) & 0xFu;
            }
        }
        if (do_push) {
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 m = nib[j];
                while (m) {
                    int e = __ffs(m) - 1;
                    m &= m - 1;
                    push_row(P.push, s_reach, wbase + j * 128 + lane * 4 + e);
                }
            }
        }
        if (P.out_bits != nullptr) {
            u32 y[SR_V];
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 x = nib[j] << ((lane & 7) * 4);
                x |= __shfl_xor_sync(FULL_MASK, x, 1);
                x |= __shfl_xor_sync(FULL_MASK, x, 2);
                x |= __shfl_xor_sync(FULL_MASK, x, 4);
                y[j] = __shfl_sync(FULL_MASK, x, (lane & 3) * 8);
            }
            u32 out = y[0];
#pragma unroll
            for (int j = 1; j < SR_V; ++j) out = ((lane >> 2) == j) ? y[j] : out;
            if (lane < 4 * SR_V) P.out_bits[(wbase >> 5) + lane] = out;
        }
    }
    if (do_push) {
        __syncthreads();
        push_flush(P.push, s_reach);
    }
    if (P.pub.n_words > 0) peer_mask_publish_tail(P.pub, P.pub_done);
}

// ---------------------------------------------------------------------------------------------
// K1c  scan_bool: Predicate<Boolean> over a BooleanColumn (M/InMemoryColumn.java:28-44; the where() the reference
// declares in DS/ColumnFilterable.java:20-22 and never implements, E/Verifier.java:82-84).  One byte per row; the
// predicate is its two-entry truth table.  A thread takes 4 x 16 consecutive rows with 128-bit streaming loads (a warp
// reads 4 x 512 contiguous bytes, all loads in flight before the first use), turns each 16 bytes into 16 bits with a
// SWAR nonzero-byte test and a gathering multiply, and lane pairs assemble the BitSet words.  1 byte per row in, 1 bit
// out: HBM-bound (6.3 TB/s at 2^30 rows).  The grid covers the whole padded bitmap, words past n are zeroed.
// ---------------------------------------------------------------------------------------------

struct ScanBoolParams {
    int64_t n;
    int64_t n_alloc_words;  // words of out_bits (a multiple of 16)
    const uint8_t* values;  // allocation padded to whole 16-byte lines
    u32 accept_false, accept_true;
    const u32* in_bits;
    u32* out_bits;
};

constexpr int SB_THREADS = 256;
constexpr int SB_V = 4;                            // 16-byte loads in flight per thread
constexpr int SB_WARP_ROWS = 32 * 16 * SB_V;       // 2048
constexpr int SB_BLOCK_ROWS = (SB_THREADS / 32) * SB_WARP_ROWS;

// 16 bytes -> 16 bits (bit k: byte k != 0).  Per 32-bit word: the SWAR nonzero-byte test leaves bit 7 of every nonzero
// byte set, and one multiply gathers bits 7 / 15 / 23 / 31 into four adjacent bits (the partial products land on
// distinct bit positions, so nothing carries).
__device__ __forceinline__ u32 sb_truth16(const int4 v) {
    const u32 w[4] = {(u32)v.x, (u32)v.y, (u32)v.z, (u32)v.w};
    u32 truth = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const u32 nz = ((((w[e] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w[e]) & 0x80808080u) >> 7;  // bits 0 / 8 / 16 / 24
        truth |= (((nz * 0x00204081u) >> 21) & 0xFu) << (4 * e);
    }
    return truth;
}

__global__ void __launch_bounds__(SB_THREADS) scan_bool_kernel(const ScanBoolParams P) {
    const int lane = threadIdx.x & 31;
    const int64_t wbase = ((int64_t)blockIdx.x * (SB_THREADS / 32) + (threadIdx.x >> 5)) * SB_WARP_ROWS;
    const u32 flip = P.accept_false ? 0xFFFFu : 0u;          // m = (truth ^ flip) & keep covers the four truth tables
    const u32 keep = P.accept_false != P.accept_true ? 0xFFFFu : (P.accept_true ? 0xFFFFu : 0u);
    const bool constant = (P.accept_false != 0) == (P.accept_true != 0);
    u32 m[SB_V];
    if (wbase + SB_WARP_ROWS <= P.n) {
        int4 v[SB_V];
#pragma unroll
        for (int j = 0; j < SB_V; ++j) v[j] = ldg_stream_v4(reinterpret_cast<const int32_t*>(P.values + wbase + j * 512 + lane * 16));
#pragma unroll
        for (int j = 0; j < SB_V; ++j) m[j] = constant ? keep : ((sb_truth16(v[j]) ^ flip) & 0xFFFFu);
    } else {
#pragma unroll
        for (int j = 0; j < SB_V; ++j) {
            const int64_t r0 = wbase + j * 512 + lane * 16;
            const int64_t left = P.n - r0;
            m[j] = 0;
            if (left > 0) {
                const int4 v = ldg_stream_v4(reinterpret_cast<const int32_t*>(P.values + r0));
                m[j] = constant ? keep : ((sb_truth16(v) ^ flip) & 0xFFFFu);
                if (left < 16) m[j] &= (1u << (int)left) - 1u;
            }
        }
    }
    u32* out_words = P.out_bits + (wbase >> 5) + (lane >> 1);
    const u32* in_words = P.in_bits != nullptr ? P.in_bits + (wbase >> 5) + (lane >> 1) : nullptr;
    const bool in_range = (wbase >> 5) + SB_WARP_ROWS / 32 <= P.n_alloc_words;  // whole-warp tile inside the padded bitmap
#pragma unroll
    for (int j = 0; j < SB_V; ++j) {
        const u32 hi = __shfl_down_sync(FULL_MASK, m[j], 1);
        if ((lane & 1) == 0 && (in_range || (wbase >> 5) + j * 16 + (lane >> 1) < P.n_alloc_words)) {
            u32 out = m[j] | (hi << 16);
            if (in_words != nullptr && out != 0) out &= in_words[j * 16];
            out_words[j * 16] = out;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2  scan_str: string predicate over an (offsets, bytes) column -> bitmask, TMA-staged, warp-specialised
//
// Replaces ExecutionContext.Node.filterSelf (E/ExecutionContext.java:79-94) over StringColumn.where
// (M/InMemoryColumn.java:71-74) with "X"::equals / s.contains("X") / s.compareTo("X") predicates
// (app/.../Runner.java:236,255-259; QueryTest.java:124-125).
//
// Persistent CTAs (grid = SMs x CTAs/SM) walk 1024-row tiles round-robin.  One PRODUCER warp (one elected lane)
// issues, per tile, two TMA bulk copies -- the tile's offsets slice and its contiguous byte range, both widened to
// 16-byte lines -- into a STAGES-deep shared-memory ring; "full" mbarriers carry the transaction bytes, "empty"
// mbarriers hand slots back.  Eight CONSUMER warps never touch global memory for input: each thread takes 4
// consecutive rows (one 128-bit + one 32-bit shared load for its 5 offsets), tests them out of shared memory, and the
// warp merges the 4-bit results into bitmask words with a shuffle butterfly.  HBM therefore only sees long, perfectly
// coalesced bursts whatever the row lengths are.  A tile whose bytes do not fit a ring slot is read from global
// memory instead (correct, slower).
//
// Two instantiation families keep the per-row instruction count low:
//   scan_str_fixed<NW>   EQ / NE / STARTS_WITH / ENDS_WITH with a needle of at most 4*NW <= 16 bytes: branch-free,
//                        every row does NW+1 shared loads, NW funnel shifts and NW xors, no divergence
//   scan_str_generic<OP> everything else (contains, compareTo, long needles): per-row loops
// ---------------------------------------------------------------------------------------------

constexpr int ST_CONSUMER_WARPS = 8;
constexpr int ST_CONSUMER_THREADS = ST_CONSUMER_WARPS * 32;  // 256
constexpr int ST_THREADS = ST_CONSUMER_THREADS + 32;         // + 1 producer warp
constexpr int ST_ROWS = 4 * ST_CONSUMER_THREADS;             // 1024 rows per tile
constexpr int ST_MAX_STAGES = 8;  // ring depth is a launch parameter (ScanStrParams::stages)
constexpr int ST_OFF_BYTES = (ST_ROWS + 4) * 4;  // offsets slice incl. the closing offset, padded to 16 B
constexpr int ST_SLACK = 32;                     // readable bytes behind the staged byte range
constexpr int ST_MAX_NEEDLE = 16384;             // needle bytes are staged in shared memory next to the ring
__host__ __device__ inline int st_needle_region(int needle_len) { return (needle_len + 16 + 15) & ~15; }
__host__ __device__ inline int st_stage_bytes(int cap) { return ST_OFF_BYTES + cap + ST_SLACK; }

enum StrOp : int {
    OP_EQ = 0, OP_CONTAINS = 1, OP_CMP_GT = 2, OP_CMP_LT = 3, OP_CMP_GE = 4, OP_CMP_LE = 5, OP_NE = 6,
    OP_STARTS_WITH = 7, OP_ENDS_WITH = 8
};

struct ScanStrParams {
    int64_t n;
    const u32* offsets;      // n + 1 entries, allocation padded to a 16-byte multiple past entry n
    const uint8_t* bytes;
    int64_t bytes_capacity;  // usable allocation size (multiple of 16, >= n_bytes + ST_SLACK)
    const uint8_t* needle;   // device copy of the needle
    int needle_len;
    int op;
    int cap;                 // bytes slot size of one ring stage (multiple of 16)
    int stages;              // ring depth, 2..ST_MAX_STAGES
    int64_t n_tiles;
    const u32* in_bits;
    u32* out_bits;
    PushD push;
    // nullable pair: `offsets` / `bytes` are pinned HOST memory streamed over PCIe; every staged tile is also written
    // to these HBM copies so that the column is resident for the next query (first-touch promotion)
    u32* promote_offsets;
    uint8_t* promote_bytes;
    PeerMaskParams pub;   // pub.n_words > 0: the last CTA publishes push.reach to the peers (multi-GPU mask exchange)
    u32* pub_done;
    u32* tile_counter;    // [0] next unclaimed tile, [1] finished CTAs; both are zero between launches
    // 1: pipelined behind the previous execution of the same query (COLQ_OPT_PIPELINE): this launch is a programmatic
    // dependent of that execution's root kernel and may start while it drains.  Until pdl_wait() it only reads immutable
    // columns, claims tiles (the previous string scan has completed: the root kernel triggers only after its own wait)
    // and sets bits in shared memory; the flush, the publish and the trigger for the next kernel come after the wait.
    u32 early;
};

#ifndef COLQ_ST_CLAIM
#define COLQ_ST_CLAIM 2
#endif
constexpr int ST_CLAIM = COLQ_ST_CLAIM;   // tiles per claim

struct StrTileMeta {
    u32 a0;    // 16-byte-aligned global byte offset the staged slice starts at (offsets are uint32)
    u32 fast;  // 1: bytes were staged in shared memory, 0: read them from global memory
    u32 sz;    // bytes of the 16-byte-aligned slice [a0, a0 + sz) covering the tile's strings
    u32 tile;  // which tile this stage holds (tiles are claimed dynamically), or ST_NO_TILE
};

// ---- word loaders: the same row tests run over the staged slice (ld.shared) or the column itself (ld.global)
struct SmemWords {
    u32 base;  // shared-state-space address of word 0
    __device__ __forceinline__ u32 operator()(u32 word_index) const {
        u32 v;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(base + word_index * 4));
        return v;
    }
};
struct GlobalWords {
    const u32* base;
    __device__ __forceinline__ u32 operator()(u32 word_index) const { return __ldg(base + word_index); }
};

__device__ __forceinline__ u32 low_mask(int nbytes) { return nbytes >= 4 ? 0xffffffffu : ((1u << (nbytes * 8)) - 1u); }

// 4 bytes starting at byte position pos (any alignment) of a little-endian word array
template <class W>
__device__ __forceinline__ u32 extract32(const W& words, u32 pos) {
    u32 i = pos >> 2;
    return __funnelshift_r(words(i), words(i + 1), (pos & 3) * 8);
}

// bytes [pos, pos+len) equal needle[0, len)
template <class W>
__device__ __forceinline__ bool bytes_equal(const W& hay, u32 pos, const u32* needle_w, int len) {
    for (int c = 0; c < len; c += 4) {
        u32 h = extract32(hay, pos + c);
        if ((h ^ needle_w[c >> 2]) & low_mask(len - c)) return false;
    }
    return true;
}

// String.compareTo sign (UTF-16 code-unit order on UTF-8 bytes; a supplementary code point, lead byte >= 0xF0, sorts
// below U+E000..U+FFFF, lead bytes 0xEE/0xEF, and above everything with a lead byte <= 0xED; two supplementary code
// points keep their byte order = code-point order = surrogate order; a differing continuation byte implies equal lead
// bytes)
__device__ __forceinline__ int utf16_key(u32 b) { return b >= 0xF0u ? 0xED * 512 + 256 + (int)(b - 0xF0u) : (int)b * 512; }
template <class W>
__device__ __forceinline__ int compare_to(const W& hay, u32 pos, int len, const u32* needle_w, int nlen) {
    int m = len < nlen ? len : nlen;
    for (int c = 0; c < m; c += 4) {
        u32 h = extract32(hay, pos + c);
        u32 d = (h ^ needle_w[c >> 2]) & low_mask(m - c);
        if (d) {
            int bi = (__ffs(d) - 1) >> 3;
            u32 hb = (h >> (bi * 8)) & 0xffu, nb = (needle_w[c >> 2] >> (bi * 8)) & 0xffu;
            return utf16_key(hb) < utf16_key(nb) ? -1 : 1;
        }
    }
    return len < nlen ? -1 : (len > nlen ? 1 : 0);
}

template <class W>
__device__ __forceinline__ bool contains(const W& hay, u32 pos, int len, const u32* needle_w, int nlen) {
    if (nlen == 0) return true;
    if (len < nlen) return false;
    const u32 first_mask = low_mask(nlen);
    const u32 first = needle_w[0] & first_mask;
    // rolling window over aligned words: one load per 4 candidate positions
    u32 wi = pos >> 2;
    u32 cur = hay(wi), nxt = hay(wi + 1);
    int sh = (int)(pos & 3);
    const int last = len - nlen;  // last candidate start
    for (int p = 0; p <= last; ++p) {
        u32 h = __funnelshift_r(cur, nxt, sh * 8);
        if (((h ^ first) & first_mask) == 0) {
            if (nlen <= 4 || bytes_equal(hay, pos + p + 4, needle_w + 1, nlen - 4)) return true;
        }
        if (++sh == 4) {
            sh = 0;
            ++wi;
            cur = nxt;
            nxt = hay(wi + 1);
        }
    }
    return false;
}

template <int OP, class W>
__device__ __forceinline__ bool str_test(const W& hay, u32 pos, int len, const u32* needle_w, int nlen) {
    if (OP == OP_EQ) return len == nlen && bytes_equal(hay, pos, needle_w, nlen);
    if (OP == OP_NE) return !(len == nlen && bytes_equal(hay, pos, needle_w, nlen));
    if (OP == OP_CONTAINS) return contains(hay, pos, len, needle_w, nlen);
    if (OP == OP_CMP_GT) return compare_to(hay, pos, len, needle_w, nlen) > 0;
    if (OP == OP_CMP_LT) return compare_to(hay, pos, len, needle_w, nlen) < 0;
    if (OP == OP_CMP_GE) return compare_to(hay, pos, len, needle_w, nlen) >= 0;
    if (OP == OP_CMP_LE) return compare_to(hay, pos, len, needle_w, nlen) <= 0;
    if (OP == OP_STARTS_WITH) return len >= nlen && bytes_equal(hay, pos, needle_w, nlen);
    if (OP == OP_ENDS_WITH) return len >= nlen && bytes_equal(hay, pos + (u32)(len - nlen), needle_w, nlen);
    return false;
}

// out-of-line copy of every operator for tiles that were not staged (one instance per kernel keeps the code small)
__device__ __noinline__ bool slow_row_test(int op, GlobalWords hay, u32 pos, int len, const u32* needle_w, int nlen) {
    switch (op) {
        case OP_EQ: return str_test<OP_EQ>(hay, pos, len, needle_w, nlen);
        case OP_CONTAINS: return str_test<OP_CONTAINS>(hay, pos, len, needle_w, nlen);
        case OP_CMP_GT: return str_test<OP_CMP_GT>(hay, pos, len, needle_w, nlen);
        case OP_CMP_LT: return str_test<OP_CMP_LT>(hay, pos, len, needle_w, nlen);
        case OP_CMP_GE: return str_test<OP_CMP_GE>(hay, pos, len, needle_w, nlen);
        case OP_CMP_LE: return str_test<OP_CMP_LE>(hay, pos, len, needle_w, nlen);
        case OP_NE: return str_test<OP_NE>(hay, pos, len, needle_w, nlen);
        case OP_STARTS_WITH: return str_test<OP_STARTS_WITH>(hay, pos, len, needle_w, nlen);
        default: return str_test<OP_ENDS_WITH>(hay, pos, len, needle_w, nlen);
    }
}

// branch-free test of one row against a needle of at most 4*NW bytes.  EQ_ONLY: "X"::equals; otherwise the operator
// (EQ / NE / STARTS_WITH / ENDS_WITH) is a run-time, warp-uniform value.
template <int NW, bool EQ_ONLY, class W>
__device__ __forceinline__ bool fixed_test(const W& hay, u32 pos, int len, const u32 (&needle)[NW], u32 last_mask, int nlen,
                                          int op) {
    u32 p = pos;
    bool len_ok;
    if (EQ_ONLY) {
        len_ok = len == nlen;
    } else {
        len_ok = (op == OP_EQ || op == OP_NE) ? (len == nlen) : (len >= nlen);
        if (op == OP_ENDS_WITH) p += (u32)(len >= nlen ? len - nlen : 0);
    }
    const u32 wi = p >> 2, sh = (p & 3) * 8;
    u32 w[NW + 1];
#pragma unroll
    for (int i = 0; i <= NW; ++i) w[i] = hay(wi + i);
    u32 diff = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        u32 h = __funnelshift_r(w[i], w[i + 1], sh);
        diff |= (h ^ needle[i]) & (i == NW - 1 ? last_mask : 0xffffffffu);
    }
    const bool m = len_ok && diff == 0;
    return (!EQ_ONLY && op == OP_NE) ? !m : m;
}

// MODE >= 0: generic path for operator MODE.  MODE in -1..-4: fixed path, EQ only, NW = -MODE needle words.
// MODE in -5..-8: fixed path, run-time operator (EQ / NE / STARTS_WITH / ENDS_WITH), NW = -MODE - 4.
// PROMOTE: the column is pinned host memory; every tile is also copied into the HBM buffers P.promote_*.
template <int MODE, bool PROMOTE = false>
__global__ void __launch_bounds__(ST_THREADS, (MODE == -1 || MODE == -2) ? 4 : 3) scan_str_kernel(const ScanStrParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    // layout: [stage: offsets | bytes(cap + slack)] x STAGES | needle words | full[] | empty[] | metas | reach
    const int stage_bytes = st_stage_bytes(P.cap);
    const int needle_region = st_needle_region(P.needle_len);
    const int n_stages = P.stages;
    u32* s_needle = reinterpret_cast<u32*>(smem + (size_t)n_stages * stage_bytes);
    u64* s_full = reinterpret_cast<u64*>(reinterpret_cast<uint8_t*>(s_needle) + needle_region);
    u64* s_empty = s_full + ST_MAX_STAGES;
    StrTileMeta* s_meta = reinterpret_cast<StrTileMeta*>(s_empty + ST_MAX_STAGES);
    u32* s_reach = reinterpret_cast<u32*>(s_meta + ST_MAX_STAGES);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool do_push = P.push.fk != nullptr;
    // the next kernel of the plan (the root's fused scan) may move onto an SM as soon as this kernel's CTAs leave it
    if (!P.early) pdl_launch_dependents();

    for (int i = tid; i < needle_region / 4; i += ST_THREADS) {
        u32 w = 0;
        for (int b = 0; b < 4; ++b) {
            int k = i * 4 + b;
            if (k < P.needle_len) w |= (u32)P.needle[k] << (8 * b);
        }
        s_needle[i] = w;
    }
    if (do_push) push_init(P.push, s_reach);
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], ST_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // Tiles are CLAIMED from a device-wide counter (r02): with a static round-robin every CTA got the same number of tiles,
    // but SMs do not get the same share of HBM bandwidth, so the slowest CTA finished 10-20 us after the median one.
    const u32 n_tiles = (u32)P.n_tiles;

    if (warp == ST_CONSUMER_WARPS) {
        // =========================== producer warp: one lane drives the TMA ring ===========================
        if (lane == 0) {
            u32 nb0 = 0, nb1 = 0;
            auto tile_bounds = [&](u32 t, u32& gb0, u32& gb1) {
                const int64_t r0 = (int64_t)t * ST_ROWS;
                const int64_t r1 = (r0 + ST_ROWS) < P.n ? (r0 + ST_ROWS) : P.n;
                gb0 = __ldg(P.offsets + r0);
                gb1 = __ldg(P.offsets + r1);
            };
            // claims are batches of ST_CLAIM consecutive tiles, and the NEXT batch is claimed while the current one is
            // being issued: neither the atomic's nor the bounds load's latency sits between two TMA issues
            // (the first tile of every CTA is static -- tile blockIdx.x, the grid never exceeds the tile count -- so the first
            // copy leaves without waiting for an atomic; the counter hands out the tiles from gridDim.x on)
            u32 next = blockIdx.x;
            tile_bounds(next, nb0, nb1);
            u32 batch = gridDim.x + atomicAdd(P.tile_counter, (u32)ST_CLAIM);
            u32 batch_next = gridDim.x + atomicAdd(P.tile_counter, (u32)ST_CLAIM);
            u32 j = ST_CLAIM - 1;  // the static tile counts as the last of a (virtual) batch: the next one is batch + 0
            int s = 0;
            u32 round = 0;  // how many times the ring has wrapped
            while (true) {
                const u32 cur = next;
                const u32 gb0 = nb0, gb1 = nb1;
                if (cur < n_tiles) {  // the tile after this one: its bounds are in flight while we wait for the slot
                    if (++j == ST_CLAIM) {
                        j = 0;
                        if (cur >= gridDim.x) {   // (not right after the static tile: `batch` is still unused then)
                            batch = batch_next;
                            batch_next = gridDim.x + atomicAdd(P.tile_counter, (u32)ST_CLAIM);
                        }
                    }
                    next = batch + j;
                    if (next < n_tiles) tile_bounds(next, nb0, nb1);
                }
                if (round > 0) mbar_wait(&s_empty[s], (round - 1) & 1);
                if (cur >= n_tiles) {  // out of tiles: tell the consumers
                    s_meta[s].tile = ST_NO_TILE;
                    mbar_arrive(&s_full[s]);
                    break;
                }
                s_meta[s].tile = cur;
                const int64_t r0 = (int64_t)cur * ST_ROWS;
                const int nr = (int)((P.n - r0) < ST_ROWS ? (P.n - r0) : ST_ROWS);
                uint8_t* base = smem + (size_t)s * stage_bytes;
                const u64 a0 = (u64)gb0 & ~(u64)15;
                const u64 a1 = ((u64)gb1 + 15) & ~(u64)15;
                const u64 sz = a1 - a0;
                const u32 off_bytes = (u32)(((nr + 1) * 4 + 15) & ~15);
                const bool fast = sz <= (u64)P.cap && (int64_t)a1 <= P.bytes_capacity;
                s_meta[s].a0 = (u32)a0;
                s_meta[s].fast = fast ? 1u : 0u;
                s_meta[s].sz = (int64_t)a1 <= P.bytes_capacity ? (u32)sz : (u32)(P.bytes_capacity - (int64_t)a0);
                mbar_arrive_expect_tx(&s_full[s], off_bytes + (fast ? (u32)sz : 0u));
                tma_bulk_g2s(base, P.offsets + r0, off_bytes, &s_full[s]);
                if (fast && sz > 0) tma_bulk_g2s(base + ST_OFF_BYTES, P.bytes + a0, (u32)sz, &s_full[s]);
                if (++s == n_stages) {
                    s = 0;
                    ++round;
                }
            }
        }
    } else {
        // =========================== consumer warps ===========================
        // needle words of the fixed path live in registers
        constexpr int NW = MODE < -4 ? -MODE - 4 : (MODE < 0 ? -MODE : 1);
        constexpr bool EQ_ONLY = MODE >= -4;
        u32 needle_r[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) needle_r[i] = s_needle[i];
        const int nlen = P.needle_len;
        const u32 last_mask = low_mask(nlen - 4 * (NW - 1));
        const int op = P.op;
        const u32* in_bits = P.in_bits;
        u32* out_bits = P.out_bits;
        const u32 n_words_out = (u32)(((P.n + 63) >> 6) << 1);  // whole 64-row BitSet words, as u32 halves

        u32 s = 0, parity = 0;
        while (true) {
            mbar_wait(&s_full[s], parity);
            const u32 tile = s_meta[s].tile;  // global tile index; tile * 1024 < 2^31 because n < 2^31
            if (tile == ST_NO_TILE) break;

            const u32 so = smem_u32(smem) + s * (u32)stage_bytes;
            const u32 r0 = tile * ST_ROWS;
            const u32 rem = (u32)P.n - r0;
            const int nr = rem < (u32)ST_ROWS ? (int)rem : ST_ROWS;
            const bool fast = s_meta[s].fast != 0;
            const u32 a0 = s_meta[s].a0;

            // this thread's 4 consecutive rows: offsets o[0..4]
            u32 o[5];
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3])
                         : "r"(so + tid * 16));
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(o[4]) : "r"(so + tid * 16 + 16));

            u32 nib = 0;
            if (fast && nr == ST_ROWS) {
                // full staged tile: no per-row validity checks
                const SmemWords hay{so + ST_OFF_BYTES};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const u32 pos = o[j] - a0;
                    const int len = (int)(o[j + 1] - o[j]);
                    bool m;
                    if (MODE < 0) m = fixed_test<NW, EQ_ONLY>(hay, pos, len, needle_r, last_mask, nlen, op);
                    else m = str_test<(MODE < 0 ? 0 : MODE)>(hay, pos, len, s_needle, nlen);
                    nib |= m ? (1u << j) : 0u;
                }
            } else if (fast) {
                const SmemWords hay{so + ST_OFF_BYTES};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool valid = tid * 4 + j < nr;
                    const u32 pos = valid ? o[j] - a0 : 0u;
                    const int len = valid ? (int)(o[j + 1] - o[j]) : -1;
                    bool m;
                    if (MODE < 0) m = fixed_test<NW, EQ_ONLY>(hay, pos, len, needle_r, last_mask, nlen, op);
                    else m = str_test<(MODE < 0 ? 0 : MODE)>(hay, pos, len, s_needle, nlen);
                    nib |= (valid && m) ? (1u << j) : 0u;
                }
            } else {
                const GlobalWords hay{reinterpret_cast<const u32*>(P.bytes)};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool valid = tid * 4 + j < nr;
                    if (!valid) continue;
                    const u32 pos = o[j];  // absolute byte position in the column (< 4 GiB)
                    const int len = (int)(o[j + 1] - o[j]);
                    const bool m = slow_row_test(op, hay, pos, len, s_needle, nlen);
                    nib |= m ? (1u << j) : 0u;
                }
            }
            if (PROMOTE) {
                // first-touch promotion: the tile just crossed PCIe into shared memory -- leave a copy in HBM
                const u32 off_lines = (u32)(((nr + 1) * 4 + 15) >> 4);
                uint4* dst_o = reinterpret_cast<uint4*>(P.promote_offsets + r0);
                for (u32 i = tid; i < off_lines; i += ST_CONSUMER_THREADS) {
                    uint4 x;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(so + i * 16));
                    dst_o[i] = x;
                }
                const u32 lines = s_meta[s].sz >> 4;
                uint4* dst_b = reinterpret_cast<uint4*>(P.promote_bytes + a0);
                if (fast) {
                    for (u32 i = tid; i < lines; i += ST_CONSUMER_THREADS) {
                        uint4 x;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(so + ST_OFF_BYTES + i * 16));
                        dst_b[i] = x;
                    }
                } else {
                    const uint4* src_b = reinterpret_cast<const uint4*>(P.bytes + a0);
                    for (u32 i = tid; i < lines; i += ST_CONSUMER_THREADS) dst_b[i] = src_b[i];
                }
            }
            // all shared-memory reads of this slot are done: hand it back to the producer
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);

            // this warp owns rows [r0 + 128 warp, +128) = 4 bitmask words
            const u32 w0 = (r0 >> 5) + warp * 4;
            if (in_bits != nullptr) {
                u32 w = in_bits[w0 + (lane >> 3)];
                nib &= (w >> ((lane & 7) * 4)) & 0xFu;
            }
            if (do_push) {
                u32 m = nib;
                while (m) {
                    int e = __ffs(m) - 1;
                    m &= m - 1;
                    push_row(P.push, s_reach, (int64_t)r0 + tid * 4 + e);
                }
            }
            if (out_bits != nullptr) {
                u32 x = nib << ((lane & 7) * 4);
                x |= __shfl_xor_sync(FULL_MASK, x, 1);
                x |= __shfl_xor_sync(FULL_MASK, x, 2);
                x |= __shfl_xor_sync(FULL_MASK, x, 4);
                u32 y = __shfl_sync(FULL_MASK, x, (lane & 3) * 8);  // lane l < 4 gets word l
                // write every word that overlaps the table, rounded up to whole 64-row BitSet words
                if (lane < 4 && (w0 + lane) < n_words_out) out_bits[w0 + lane] = y;
            }
            if (++s == (u32)n_stages) {
                s = 0;
                parity ^= 1;
            }
        }
    }

    __syncthreads();  // every claim of this CTA has been made
    if (P.early) {
        pdl_wait();  // the previous execution's root kernel has completed: it has read and re-zeroed the push target
        pdl_launch_dependents();
    }
    if (do_push) push_flush(P.push, s_reach);
    if (tid == 0 && atomicAdd(P.tile_counter + 1, 1u) == gridDim.x - 1) {  // last CTA: re-arm the claim counter
        P.tile_counter[0] = 0;
        P.tile_counter[1] = 0;
    }
    if (P.pub.n_words > 0) peer_mask_publish_tail(P.pub, P.pub_done);
}

// ---------------------------------------------------------------------------------------------
// K4b  csr_pull: to-many association hop on the forward side
//
// Replaces ExecutionContext.Node.filterParent, Association.Many branch (E/ExecutionContext.java:111-113), in pull
// form: parent row r keeps its bit iff any of its targets matches in the child.  Used for the 51-row state
// adjacency (219 edges) and the QueryTest garden grid.
// ---------------------------------------------------------------------------------------------

struct CsrPullParams {
    int64_t n;               // parent rows
    const int64_t* offsets;  // n + 1
    const int32_t* targets;
    const u32* child_bits;   // null = every child row matches
    int64_t n_child;
    const u32* in_bits;      // nullable
    u32* out_bits;
    PushD push;
    PeerMaskParams pm;       // pm.n_words > 0 (single-block launches only): collect the child mask exchange first
    int64_t nnz;             // number of edges (targets)
};

constexpr int CSR_SMEM_EDGES = 2048;  // single-block launches stage up to this many targets (and a <= 4096-row child mask)

__global__ void __launch_bounds__(256) csr_pull_kernel(const CsrPullParams P) {
    __shared__ u32 s_reach[PUSH_SMEM_WORDS];
    __shared__ int32_t s_tgt[CSR_SMEM_EDGES];
    __shared__ u32 s_child[PUSH_SMEM_WORDS];
    if (P.pm.n_words > 0) peer_mask_collect(P.pm);
    const bool do_push = P.push.fk != nullptr;
    if (do_push) push_init(P.push, s_reach);
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    // the 51-row state adjacency (219 edges) is one block of pure latency: issue every global load at once -- row bounds,
    // all targets and the child mask -- instead of a chain of dependent loads per edge
    const bool staged = gridDim.x == 1 && P.nnz <= CSR_SMEM_EDGES && P.n_child <= PUSH_SMEM_BITS;
    int64_t e0 = 0, e1 = 0;
    if (r < P.n) {
        e0 = P.offsets[r];
        e1 = P.offsets[r + 1];
    }
    if (staged) {
        for (int i = threadIdx.x; i < (int)P.nnz; i += blockDim.x) s_tgt[i] = P.targets[i];
        const int cw = (int)((P.n_child + 31) >> 5);
        for (int i = threadIdx.x; i < cw; i += blockDim.x) s_child[i] = P.child_bits != nullptr ? P.child_bits[i] : 0xffffffffu;
    }
    if (do_push || staged) __syncthreads();
    bool ok = false;
    if (staged) {
        for (int64_t e = e0; e < e1 && !ok; ++e) {
            const int32_t t = s_tgt[e];
            if (t >= 0 && t < P.n_child) ok = (s_child[t >> 5] >> (t & 31)) & 1u;
        }
    } else if (r < P.n) {
        for (int64_t e = e0; e < e1 && !ok; ++e) {
            int32_t t = P.targets[e];
            if (t >= 0 && t < P.n_child) ok = P.child_bits == nullptr ? true : bit_test(P.child_bits, t);
        }
    }
    u32 word = __ballot_sync(FULL_MASK, ok);
    const int64_t wbase = r - lane;  // first row of this warp
    if (wbase < ((P.n + 63) & ~(int64_t)63)) {
        if (P.in_bits != nullptr) word &= P.in_bits[wbase >> 5];
        if (do_push && ((word >> lane) & 1u)) push_row(P.push, s_reach, r);
        if (lane == 0 && P.out_bits != nullptr) P.out_bits[wbase >> 5] = word;
    }
    if (do_push) {
        __syncthreads();
        push_flush(P.push, s_reach);
    }
}

// ---------------------------------------------------------------------------------------------
// K4a/K4c  push from a materialised child bitmask (reverse-side hop whose child was not fused), AND, fill
// Replaces the One / Many / None switch of filterParent (E/ExecutionContext.java:105-119) and BitSet.and (:121).
// ---------------------------------------------------------------------------------------------

struct PushBitsParams {
    int64_t n_child;
    const u32* child_bits;   // null = all child rows
    const int32_t* fk;       // to-one forward column on the child, or null when CSR
    const int64_t* offsets;  // CSR forward column on the child
    const int32_t* targets;
    u32* reach;
    int64_t n_parent;
    u32* oob;                // nullable: see PushD::oob
};

__global__ void __launch_bounds__(256) push_bits_kernel(const PushBitsParams P) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < P.n_child; r += stride) {
        if (P.child_bits != nullptr && !bit_test(P.child_bits, r)) continue;
        if (P.fk != nullptr) {
            int32_t t = P.fk[r];
            if (t >= 0 && t < P.n_parent) {
                u32 m = 1u << (t & 31);
                if (!(P.reach[t >> 5] & m)) atomicOr(&P.reach[t >> 5], m);
            } else if (t != -1 && P.oob != nullptr) {
                *P.oob = 1u;
            }
        } else {
            for (int64_t e = P.offsets[r]; e < P.offsets[r + 1]; ++e) {
                int32_t t = P.targets[e];
                if (t >= 0 && t < P.n_parent) {
                    u32 m = 1u << (t & 31);
                    if (!(P.reach[t >> 5] & m)) atomicOr(&P.reach[t >> 5], m);
                }
            }
        }
    }
}

// dst &= src  (BitSet.and, E/ExecutionContext.java:121)
__global__ void __launch_bounds__(256) and_words_kernel(u32* dst, const u32* src, int64_t n_words) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) dst[i] &= src[i];
}

// matchingBits.set(0, table.size()) (E/ExecutionContext.java:83-87); words past n stay zero
__global__ void __launch_bounds__(256) fill_ones_kernel(u32* dst, int64_t n_rows, int64_t n_words_alloc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words_alloc; i += stride) {
        int64_t lo = i << 5;
        u32 w = 0;
        if (lo + 32 <= n_rows) w = 0xffffffffu;
        else if (lo < n_rows) w = (1u << (n_rows - lo)) - 1u;
        dst[i] = w;
    }
}

// dst[w] = OR over ranks of gathered[rank * n_words + w]: the consumer half of the state-mask OR-allreduce
__global__ void __launch_bounds__(256) or_ranks_kernel(u32* dst, const u32* gathered, int64_t n_words, int n_ranks) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) {
        u32 w = 0;
        for (int r = 0; r < n_ranks; ++r) w |= gathered[(int64_t)r * n_words + i];
        dst[i] = w;
    }
}

// ---------------------------------------------------------------------------------------------
// K3  stream compaction: bitmask -> ascending row indices
//
// Replaces the index half of InMemoryTable.subset (M/InMemoryTable.java:121-131: cardinality(), then an ordered
// copy).  Three launches: per-block popcount, single-block exclusive scan of block counts, ordered write with a
// warp-shuffle scan inside each block.  Output order is ascending by construction.
// ---------------------------------------------------------------------------------------------

constexpr int CP_THREADS = 256;
constexpr int CP_WORDS_PER_BLOCK = CP_THREADS * 4;  // one 128-bit load per thread: 32768 rows per block

__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32* s_warp, u32& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(FULL_MASK, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
        u32 wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 t = __shfl_up_sync(FULL_MASK, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < (int)(blockDim.x >> 5)) s_warp[lane] = wi - w;
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    block_total = s_warp[32];
    return s_warp[warp] + incl - v;
}

__device__ __forceinline__ uint4 load_words4(const u32* bits, int64_t w0, int64_t n_words) {
    // bitmask allocations are padded to whole compaction blocks, so the vector load is always in bounds
    // L2 (.cg) load: compact_fused_kernel<NG > 0> clears bits with atomics between its two passes over a tile
    uint4 v = __ldcg(reinterpret_cast<const uint4*>(bits + w0));
    if (w0 + 0 >= n_words) v.x = 0;
    if (w0 + 1 >= n_words) v.y = 0;
    if (w0 + 2 >= n_words) v.z = 0;
    if (w0 + 3 >= n_words) v.w = 0;
    return v;
}

__global__ void __launch_bounds__(CP_THREADS) popc_blocks_kernel(const u32* bits, int64_t n_words, u32* block_counts) {
    __shared__ u32 s_warp[33];
    const int64_t w0 = ((int64_t)blockIdx.x * CP_THREADS + threadIdx.x) * 4;
    uint4 v = load_words4(bits, w0, n_words);
    u32 c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    u32 total;
    block_exclusive_scan(c, s_warp, total);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// single block: exclusive scan of block_counts -> block_offsets, total -> *out_total
__global__ void __launch_bounds__(1024) scan_counts_kernel(const u32* block_counts, int64_t n_blocks, u64* block_offsets,
                                                          u64* out_total) {
    __shared__ u32 s_warp[33];
    __shared__ u64 s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_blocks; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        u32 v = i < n_blocks ? block_counts[i] : 0;
        u32 total;
        u32 ex = block_exclusive_scan(v, s_warp, total);
        const u64 carry = s_carry;
        if (i < n_blocks) block_offsets[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_total = s_carry;
}

__global__ void __launch_bounds__(CP_THREADS) compact_kernel(const u32* bits, int64_t n_words, const u64* block_offsets,
                                                            int32_t* out_idx, int64_t capacity, int64_t row_base) {
    __shared__ u32 s_warp[33];
    const int64_t w0 = ((int64_t)blockIdx.x * CP_THREADS + threadIdx.x) * 4;
    uint4 v = load_words4(bits, w0, n_words);
    u32 c = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    u32 total;
    u32 ex = block_exclusive_scan(c, s_warp, total);
    if (total == 0 || c == 0) return;
    int64_t pos = (int64_t)block_offsets[blockIdx.x] + ex;
    const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        u32 m = w[k];
        const int64_t rb = row_base + ((w0 + k) << 5);
        while (m) {
            int b = __ffs(m) - 1;
            m &= m - 1;
            if (pos < capacity) out_idx[pos] = (int32_t)(rb + b);
            ++pos;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// final gather (multi-GPU): every rank contributes one fixed-size block [u64 count | pad | cap indices]; after the
// all-gather this kernel concatenates the valid prefixes in rank order (= ascending global row order, because
// ranks own ascending universe ranges).  info[0] = rows written, info[1] = largest per-rank count (overflow check),
// info[2] = true total.
// ---------------------------------------------------------------------------------------------
constexpr int GATHER_HEADER_WORDS = 4;
constexpr int RESULT_FLAGS_WORD = 2;  // result block = [u64 count | u32 flags | u32 pad | indices...]; flags bit 0: a walked
                                      // to-one target was outside its table (lazy range check of host-resident columns)

__global__ void __launch_bounds__(256) unpack_gather_kernel(const int32_t* blocks, int n_ranks, int64_t cap, int32_t* out,
                                                          u64* info) {
    __shared__ int64_t s_off[MAX_RANKS + 1];
    const int64_t stride_words = cap + GATHER_HEADER_WORDS;
    if (threadIdx.x == 0) {
        int64_t off = 0;
        u64 maxc = 0, total = 0;
        for (int r = 0; r < n_ranks; ++r) {
            u64 c = *reinterpret_cast<const u64*>(blocks + r * stride_words);
            s_off[r] = off;
            off += (int64_t)(c < (u64)cap ? c : (u64)cap);
            maxc = c > maxc ? c : maxc;
            total += c;
        }
        s_off[n_ranks] = off;
        if (blockIdx.x == 0) {
            info[0] = (u64)off;
            info[1] = maxc;
            info[2] = total;
        }
    }
    __syncthreads();
    const int64_t n = s_off[n_ranks];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int r = 0;
        while (r + 1 < n_ranks && i >= s_off[r + 1]) ++r;
        out[i] = blocks[r * stride_words + GATHER_HEADER_WORDS + (i - s_off[r])];
    }
}

struct PeerGatherParams {
    const u64* count;       // this rank's match count (device)
    const int32_t* idx;     // this rank's ascending global row indices
    int64_t idx_capacity;   // how many of them were actually written
    int64_t slot_cap;       // EFFECTIVE per-rank cap: min(index capacity, indices one mailbox slot can hold) -- the same
                            // number on the sending and on the receiving side
    size_t slot_bytes;
    int n_ranks, rank;
    uint8_t* const* peers;
    u64 epoch;
    u32* done;              // [n_ranks] block-completion counters (zero between launches)
    int blocks_per_peer;
    u32* status;
    int32_t* out;           // concatenation of all ranks' indices in rank order
    u64* info;              // [0] rows written, [1] largest per-rank count, [2] true total
    int wait;               // fused gather: 1 = the launch ends only when every peer's flag has arrived (gather_tail);
                            // 0 = publish only, the flags are awaited when the host fetches (peer_gather_recv_kernel)
};

// Gather slots hold flag-in-data words like the mask slots: [u64 {epoch | count}] + pad to GATHER_SLOT_HEADER, then one
// u64 {epoch | index} per matched row.  Nothing in the gather needs a fence: every word carries its own epoch tag and the
// receiver polls the words it needs.
__device__ __forceinline__ u64* gather_slot(const PeerGatherParams& P, int dst_rank, int src_rank) {
    return reinterpret_cast<u64*>(P.peers[dst_rank] + PEER_GATHER_AREA_OFFSET + ((size_t)(P.epoch & 1) * P.n_ranks + src_rank) * P.slot_bytes);
}

// final gather as a launch of its own (COLQ_OPT_FUSED_GATHER=0 and the non-cooperative compactions): every rank stores its
// indices into its slot of every peer's mailbox
__global__ void __launch_bounds__(256) peer_gather_send_kernel(const PeerGatherParams P) {
    const int peer = blockIdx.x / P.blocks_per_peer, part = blockIdx.x % P.blocks_per_peer;
    u64* slot = gather_slot(P, peer, P.rank);
    const u64 true_count = *P.count;
    int64_t n = (int64_t)true_count;
    if (n > P.idx_capacity) n = P.idx_capacity;
    if (n > P.slot_cap) n = P.slot_cap;
    u64* dst = slot + GATHER_SLOT_HEADER / 8;
    const int64_t per = (n + P.blocks_per_peer - 1) / P.blocks_per_peer;
    const int64_t lo = part * per, hi = (lo + per) < n ? (lo + per) : n;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) ll_store(dst + i, (u32)P.idx[i], P.epoch);
    if (part == 0 && threadIdx.x == 0) ll_store(slot, (u32)true_count, P.epoch);
}

// final gather, receive half: wait for every rank's count word, then concatenate the valid prefixes in rank order (each
// index word is polled for its own epoch tag).  info[0] = rows written, info[1] = largest per-rank count, info[2] = total.
__global__ void __launch_bounds__(256) peer_gather_recv_kernel(const PeerGatherParams P) {
    __shared__ int64_t s_off[MAX_RANKS + 1];
    __shared__ u32 s_cnt[MAX_RANKS];
    if ((int)threadIdx.x < P.n_ranks) s_cnt[threadIdx.x] = ll_load(gather_slot(P, P.rank, threadIdx.x), P.epoch, P.status);
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t off = 0;
        u64 maxc = 0, total = 0;
        for (int r = 0; r < P.n_ranks; ++r) {
            const u64 c = s_cnt[r];
            s_off[r] = off;
            off += (int64_t)(c < (u64)P.slot_cap ? c : (u64)P.slot_cap);
            maxc = c > maxc ? c : maxc;
            total += c;
        }
        s_off[P.n_ranks] = off;
        if (blockIdx.x == 0) {
            P.info[0] = (u64)off;
            P.info[1] = maxc;
            P.info[2] = total;
        }
    }
    __syncthreads();
    const int64_t n = s_off[P.n_ranks];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int r = 0;
        while (r + 1 < P.n_ranks && i >= s_off[r + 1]) ++r;
        P.out[i] = (int32_t)ll_load(gather_slot(P, P.rank, r) + GATHER_SLOT_HEADER / 8 + (i - s_off[r]), P.epoch, P.status);
    }
}

// ---- final gather fused into the kernel that writes the indices (compact_fused<.., true>, root_fused) --------------
// The writer of result position `pos` stores the index word straight into slot [parity][my rank] of EVERY rank's mailbox
// (8-byte flag-in-data stores over NVLink, coalesced per warp because neighbouring threads hold neighbouring positions):
// no separate send launch, no second pass over the index list, no fence.  P.slot_cap is the EFFECTIVE per-rank cap
// min(index capacity, mailbox slot capacity): sender and receiver clamp to the same number.
__device__ __forceinline__ void gather_store(const PeerGatherParams& P, int64_t pos, int32_t v) {
    if (pos >= P.slot_cap) return;
    for (int r = 0; r < P.n_ranks; ++r) ll_store(gather_slot(P, r, P.rank) + GATHER_SLOT_HEADER / 8 + pos, (u32)v, P.epoch);
}

// Run by ONE block (all of its threads) after every block of the launch has issued its gather_store()s and the total is
// known: publishes this rank's true count in every rank's slot header (its arrival tells a peer that this rank's launch
// has reached its end).  P.wait: also wait until every rank's count has arrived here and summarise them -- that keeps
// ranks within one execution of each other when nothing else in the plan does; otherwise the counts are awaited when the
// host fetches (peer_gather_recv_kernel), which is also where the slots are concatenated into one list.
__device__ __forceinline__ void gather_tail(const PeerGatherParams& P, const u64* total) {
    __shared__ u32 s_gt_cnt[MAX_RANKS];
    const u64 true_count = *reinterpret_cast<const volatile u64*>(total);
    if ((int)threadIdx.x < P.n_ranks) ll_store(gather_slot(P, threadIdx.x, P.rank), (u32)true_count, P.epoch);
    if (!P.wait) return;
    if ((int)threadIdx.x < P.n_ranks) s_gt_cnt[threadIdx.x] = ll_load(gather_slot(P, P.rank, threadIdx.x), P.epoch, P.status);
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 rows = 0, maxc = 0, sum = 0;
        for (int r = 0; r < P.n_ranks; ++r) {
            const u64 c = s_gt_cnt[r];
            rows += c < (u64)P.slot_cap ? c : (u64)P.slot_cap;
            maxc = c > maxc ? c : maxc;
            sum += c;
        }
        P.info[0] = rows;
        P.info[1] = maxc;
        P.info[2] = sum;
    }
}

// ---------------------------------------------------------------------------------------------
// K3'  single-launch compaction (cooperative): per-tile popcount, grid barrier, every block scans the tile counts it
// needs and writes its indices.  Same result as popc_blocks + scan_counts + compact with one launch instead of three;
// the grid is sized to be co-resident (cudaLaunchCooperativeKernel) so the hand-rolled barrier cannot deadlock.
// ---------------------------------------------------------------------------------------------

constexpr int CF_VEC = 4;                                   // 128-bit loads per thread per tile
constexpr int CF_WORDS_PER_TILE = CP_THREADS * 4 * CF_VEC;  // 4096 words = 131072 rows

constexpr int CF_MAX_GATHER = 2;

struct CompactFusedParams {
    u32* bits;             // root bitmask; phase 1 clears the bits whose deferred FK chain fails (NG > 0)
    int64_t n_words;
    int64_t n_tiles;       // tiles of CF_WORDS_PER_TILE words
    u32* tile_counts;      // [n_tiles]
    u32* barrier;          // {arrival count, generation}: self-resetting, no host-side state
    u64* total;
    int32_t* out_idx;
    int64_t capacity;
    int64_t row_base;
    int64_t n_rows;
    GatherD gather[CF_MAX_GATHER];  // deferred to-one chains of the root node (see compact_fused_kernel)
    PeerGatherParams pg;            // GATHER kernels: the multi-GPU final gather runs as phases 3 and 4 of this launch
};

// sense-reversal grid barrier on {arrival count, generation}; self-resetting.  The grid must be co-resident
// (cooperative launch).
__device__ __forceinline__ void grid_barrier(u32* bar) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile u32* gen = bar + 1;
        const u32 my_gen = *gen;
        if (atomicAdd(bar, 1u) == gridDim.x - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == my_gen) {}
        }
        __threadfence();
    }
    __syncthreads();
}

// thread t of a tile owns words [16 t, 16 t + 16): four consecutive 128-bit loads
__device__ __forceinline__ u32 cf_load(const CompactFusedParams& P, int64_t tile, uint4 (&v)[CF_VEC]) {
    const int64_t w0 = tile * CF_WORDS_PER_TILE + (int64_t)threadIdx.x * 4 * CF_VEC;
    u32 c = 0;
#pragma unroll
    for (int k = 0; k < CF_VEC; ++k) {
        v[k] = load_words4(P.bits, w0 + 4 * k, P.n_words);
        c += __popc(v[k].x) + __popc(v[k].y) + __popc(v[k].z) + __popc(v[k].w);
    }
    return c;
}

// NG > 0: the root node's criteria-free to-one FK chains (ExecutionContext.Node.filterParent, One branch,
// E/ExecutionContext.java:114, in pull form) were DEFERRED by the planner: the root's row scan ran as a pure
// coalesced predicate scan, and the chains are walked here, in phase 1, only for the rows whose bit is still set.
// After a selective predicate that is a fraction of a percent of the rows, spread evenly over every resident
// thread of the grid (all walks in flight at once) instead of stalling the streaming warps of the scan.
constexpr int CF_LIST_CAP = 4096;  // survivors of one 131072-row tile that are resolved block-wide (3 % selectivity)

// GATHER (multi-GPU, sharded root): the ordered write of phase 2 also stores every index into all ranks' mailbox slots
// over NVLink (gather_store) and the last block to finish publishes the flags and waits for the peers (gather_tail):
// no extra grid barrier, no send / receive launches (r01 ran the gather as two more barrier-separated phases of this
// kernel, which was slower than separate launches; writing from phase 2 is not).
template <int NG, bool GATHER = false>
__global__ void __launch_bounds__(CP_THREADS) compact_fused_kernel(const CompactFusedParams P) {
    __shared__ u32 s_warp[33];
    __shared__ u64 s_base;
    __shared__ u32 s_list[NG > 0 ? CF_LIST_CAP : 1];
    __shared__ u32 s_removed;
    uint4 v[CF_VEC];
    {
        // phase 1: (resolve deferred chains,) popcount my tiles
        for (int64_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
            u32 c = cf_load(P, t, v);
            u32 total;
            const u32 ex = block_exclusive_scan(c, s_warp, total);
            if (NG > 0 && total != 0) {
                const u32 lw0 = threadIdx.x * 4 * CF_VEC;  // first word of this thread inside the tile
                const int64_t tile_w0 = t * CF_WORDS_PER_TILE;
                if (total <= CF_LIST_CAP) {
                    // the tile's surviving rows go into one shared list and are walked round-robin by the whole
                    // block: every chain of the tile is in flight at once, whatever thread found the bit
                    if (threadIdx.x == 0) s_removed = 0;
                    u32 pos = ex;
#pragma unroll
                    for (int k = 0; k < CF_VEC; ++k) {
                        const u32 w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            u32 m = w[j];
                            while (m) {
                                const int b = __ffs(m) - 1;
                                m &= m - 1;
                                s_list[pos++] = ((lw0 + 4 * k + j) << 5) + b;
                            }
                        }
                    }
                    __syncthreads();
                    u32 removed = 0;
                    for (u32 i = threadIdx.x; i < total; i += CP_THREADS) {
                        const u32 lr = s_list[i];
                        const int64_t row = (tile_w0 << 5) + lr;
                        bool ok = row < P.n_rows;
#pragma unroll
                        for (int g = 0; g < NG; ++g) ok = ok && gather_eval(P.gather[g], row, 0);
                        if (!ok) {
                            atomicAnd(&P.bits[tile_w0 + (lr >> 5)], ~(1u << (lr & 31)));
                            ++removed;
                        }
                    }
                    if (removed) atomicAdd(&s_removed, removed);
                    __syncthreads();
                    total -= s_removed;
                } else {
                    // dense tile: every thread walks the bits of its own 16 words
                    c = 0;
#pragma unroll
                    for (int k = 0; k < CF_VEC; ++k) {
                        const u32 w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            u32 m = w[j], keep = w[j];
                            const int64_t rb = (tile_w0 + lw0 + 4 * k + j) << 5;
                            while (m) {
                                const int b = __ffs(m) - 1;
                                m &= m - 1;
                                bool ok = rb + b < P.n_rows;
#pragma unroll
                                for (int g = 0; g < NG; ++g) ok = ok && gather_eval(P.gather[g], rb + b, 0);
                                if (!ok) keep &= ~(1u << b);
                            }
                            if (keep != w[j]) P.bits[tile_w0 + lw0 + 4 * k + j] = keep;
                            c += __popc(keep);
                        }
                    }
                    __syncthreads();
                    block_exclusive_scan(c, s_warp, total);
                }
            }
            if (threadIdx.x == 0) P.tile_counts[t] = total;
            __syncthreads();
        }
        grid_barrier(P.barrier);
    }
    // phase 2: prefix of each of my tiles = sum of the counts of all earlier tiles (carried forward), then write
    int64_t prev_tile = 0;
    u64 running = 0;
    for (int64_t t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        u32 part = 0;
        for (int64_t i = prev_tile + threadIdx.x; i < t; i += CP_THREADS) part += P.tile_counts[i];
        u32 seg_total;
        block_exclusive_scan(part, s_warp, seg_total);
        if (threadIdx.x == 0) s_base = running + seg_total;
        __syncthreads();
        running = s_base;
        prev_tile = t;
        if (P.tile_counts[t] != 0) {
            u32 c = cf_load(P, t, v);
            u32 total;
            u32 ex = block_exclusive_scan(c, s_warp, total);
            int64_t pos = (int64_t)running + ex;
            const int64_t w0 = t * CF_WORDS_PER_TILE + (int64_t)threadIdx.x * 4 * CF_VEC;
#pragma unroll
            for (int k = 0; k < CF_VEC; ++k) {
                const u32 w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    u32 m = w[j];
                    const int64_t rb = P.row_base + ((w0 + 4 * k + j) << 5);
                    while (m) {
                        int b = __ffs(m) - 1;
                        m &= m - 1;
                        if (pos < P.capacity) P.out_idx[pos] = (int32_t)(rb + b);
                        if (GATHER) gather_store(P.pg, pos, (int32_t)(rb + b));
                        ++pos;
                    }
                }
            }
        }
        __syncthreads();
    }
    // the block that owns the last tile knows the grand total
    if ((P.n_tiles - 1) % gridDim.x == blockIdx.x && threadIdx.x == 0) *P.total = running + P.tile_counts[P.n_tiles - 1];

    if (GATHER) {
        // the indices already went to every rank's mailbox slot in phase 2; the last block to finish publishes the count
        // and the epoch flags, then waits for every peer's flag (gather_tail)
        const PeerGatherParams& G = P.pg;
        __shared__ u32 s_last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();  // (the total, for the last block; the gathered index words carry their own tags)
            s_last = (atomicAdd(&G.done[0], 1u) == gridDim.x - 1) ? 1u : 0u;
        }
        __syncthreads();
        if (s_last) {
            if (threadIdx.x == 0) G.done[0] = 0;
            gather_tail(G, P.total);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1+K4+K3  root_fused: the whole ROOT node in one persistent launch
//
//   filterSelf of the root (E/ExecutionContext.java:79-94, int predicates as in scan_rows)
//   + the root's deferred to-one chains (filterParent One branch, :114, pull form)
//   + (optional) a tiny to-many hop feeding such a chain, e.g. the 51-row / 219-edge state adjacency (:111-113), with the
//     multi-GPU mask COLLECT in front of it
//   + the ascending index half of Table.subset (M/InMemoryTable.java:121-131)
//   + (multi-GPU) the final gather of the matched indices into every rank's mailbox.
//
// r01 ran these as scan_rows -> csr_pull -> compact_fused -> peer_gather_send -> peer_gather_recv: the compaction
// re-read the dense root mask twice across a grid barrier (57 us at 38 % of DRAM bandwidth for 0.17 GB) and the
// launch-latency-class kernels were a third of the 8-GPU step.  Here:
//
// Phase A (bandwidth): every WARP owns a contiguous range of 512-row chunks (virtual warp id = CTA ticket * 8 + warp,
//   tickets handed out in CTA start order).  It streams the predicate columns with the same 128-bit streaming loads as
//   scan_rows, stores the mask words (the root BitSet stays available to the host), and appends the surviving rows --
//   0.16 % for the population predicate -- to its own ordered candidate list in global memory (ballot + warp prefix, no
//   block synchronisation anywhere in the streaming loop).  Ascending warp id == ascending row order.
// Phase B (latency, ~400 candidates per CTA): (1) collect the exchanged mask and redo the tiny to-many hop into shared
//   memory -- every CTA does it redundantly, it is 219 edges; (2) walk the chains of all candidates of the CTA at once,
//   one thread per candidate, clearing failed bits in the mask with atomicAnd; (3) publish the CTA's survivor count and
//   sum the counts of all lower tickets (decoupled look-back: a CTA only waits for CTAs that started before it, so an
//   ordinary launch cannot deadlock); (4) ordered write of the indices, locally and (multi-GPU) into all peers' slots.
//   The last CTA to finish re-arms the counters and runs gather_tail.
// A CTA whose candidate list overflowed (selectivity above ~3 %) falls back, for its own row range only, to the dense
//   per-word walk over the mask it just wrote (same as compact_fused's dense tiles).
// ---------------------------------------------------------------------------------------------

constexpr int RF_THREADS = 256;
constexpr int RF_WARPS = RF_THREADS / 32;
constexpr int RF_ILP = 2;                     // candidates whose chains one lane walks concurrently
constexpr u32 RF_NO_LEAF = 0xffffffffu;       // a chain that ended in Association.None (or outside its table)
constexpr int RF_PRE_ROWS = PUSH_SMEM_BITS;   // a folded to-many hop has at most this many parent / child rows ...
constexpr int RF_PRE_EDGES = CSR_SMEM_EDGES;  // ... and this many edges

struct RootFusedParams {
    int64_t n;                   // root rows
    IntPredD pred[SR_MAX_PRED];  // NP of them (promote pointers unused: the planner does not fuse a promoting scan)
    const u32* in_bits;          // nullable: bits the root already has from other launches
    u32* bits;                   // root mask, always written
    int64_t n_chunks;            // ceil(n / 512)
    int64_t chunks_per_warp;
    u32* lists;                  // [gridDim.x * RF_WARPS][list_cap][1 + ng]: candidate row (bit 31: dropped), chain leaf rows
    int list_cap;
    u32* counters;               // [0] CTA tickets, [1] finished CTAs; both are zero between launches
    u64* cta_state;              // [gridDim.x]  epoch << 32 | survivors of that ticket
    u32 epoch;                   // differs from the previous launch on the same cta_state (host resets at wrap)
    int ng;
    GatherD gather[CF_MAX_GATHER];
    u32 pre_mask;                // bit g: gather[g].bits is the output of `pre` (held in shared memory)
    CsrPullParams pre;           // pre.n > 0: folded tiny to-many hop (pm.n_words > 0: collect the mask exchange first)
    u64* total;
    int32_t* out_idx;
    int64_t capacity;
    int64_t row_base;
    PeerGatherParams pg;         // pg.n_ranks > 0: final gather fused in
    u32* clean;                  // nullable: the push target behind the folded hop; the last CTA re-zeroes it, so the next
    int clean_words;             //   execution needs no memset in front of its first scan (COLQ_OPT_PIPELINE)
    u64* dbg;                    // (COLQ_RF_DEBUG builds) 8 globaltimer stamps per CTA
    // root_finish_kernel only (the candidates were listed per 512-row chunk by scan_rows<.., LIST>):
    u32* ucount;                 // [n_chunks] candidates of the chunk (> list_cap: its list overflowed)
    int64_t units_per_cta;       // ticket v finishes the chunks [v * units_per_cta, (v + 1) * units_per_cta); <= RF_MAX_UPC
};

__device__ __forceinline__ bool rf_chain(const GatherD& g, const u32* bits, int64_t r) {
    for (int d = 0; d < g.depth; ++d) {
        const int32_t t = g.fk[d][r];
        if (t < 0 || t >= g.n[d]) {
            if (t != -1 && g.oob != nullptr) *g.oob = 1u;
            return false;
        }
        r = t;
    }
    return bits == nullptr ? true : bit_test(bits, r);
}

// The folded tiny to-many hop of root_fused / root_finish (and the COLLECT of the mask exchange in front of it) into shared
// memory: s_pre receives the hop's output bits.  Block-wide (RF_THREADS threads); every CTA runs it redundantly.
__device__ __forceinline__ void rf_run_pre(const CsrPullParams& C, u32 vcta, u32* s_pre, u32* s_child, int32_t* s_tgt) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int cw = (int)((C.n_child + 31) >> 5);
    if (C.pm.n_words > 0) {
        // COLLECT half of the OR-exchange, straight into shared memory (every CTA polls its own rank's mailbox)
        const PeerMaskParams& M = C.pm;
        const size_t area = (size_t)(M.epoch & 1) * MAX_RANKS * MASK_SLOT_BYTES;
        const uint8_t* mine = M.peers[M.rank] + area;
        // one thread per (rank, word): all polls are in flight at once -- eight ranks polled one after the other by the
        // same thread put 5 us of L2 round trips on the critical path of every step
        for (int w = tid; w < cw; w += RF_THREADS) s_child[w] = 0;
        __syncthreads();
        for (int i = tid; i < cw * M.n_ranks; i += RF_THREADS) {
            const int r = i / cw, w = i - r * cw;
            const u32 v = ll_load(reinterpret_cast<const u64*>(mine + (size_t)r * MASK_SLOT_BYTES + 16) + w, M.epoch, M.status);
            if (v) atomicOr(&s_child[w], v);
        }
        __syncthreads();
        if (vcta == 0 && M.reach != nullptr)
            for (int w = tid; w < cw; w += RF_THREADS) M.reach[w] = s_child[w];  // the reduced mask stays readable (node cardinalities)
    } else {
        for (int w = tid; w < cw; w += RF_THREADS) s_child[w] = C.child_bits != nullptr ? __ldcg(C.child_bits + w) : 0xffffffffu;
    }
    for (int i = tid; i < (int)C.nnz; i += RF_THREADS) s_tgt[i] = C.targets[i];
    __syncthreads();
    const int rows_pad = (int)((C.n + 31) & ~(int64_t)31);
    for (int r = tid; r < rows_pad; r += RF_THREADS) {
        bool ok = false;
        if (r < C.n) {
            const int64_t e0 = C.offsets[r], e1 = C.offsets[r + 1];
            for (int64_t e = e0; e < e1 && !ok; ++e) {
                const int32_t t = s_tgt[e];
                if (t >= 0 && t < C.n_child) ok = (s_child[t >> 5] >> (t & 31)) & 1u;
            }
        }
        u32 word = __ballot_sync(FULL_MASK, ok);
        if (C.in_bits != nullptr) word &= C.in_bits[r >> 5];
        if (lane == 0) {
            s_pre[r >> 5] = word;
            if (vcta == 0 && C.out_bits != nullptr) C.out_bits[r >> 5] = word;
        }
    }
}

#ifdef COLQ_RF_DEBUG
#define RF_STAMP(k) do { if (threadIdx.x == 0 && P.dbg != nullptr) P.dbg[(size_t)blockIdx.x * 8 + (k)] = global_timer_ns(); } while (0)
#else
#define RF_STAMP(k) do { } while (0)
#endif

template <int NP>
__global__ void __launch_bounds__(RF_THREADS, 4) root_fused_kernel(const RootFusedParams P) {
    __shared__ u32 s_warp[33];
    __shared__ u32 s_ticket, s_last, s_overflow;
    __shared__ u32 s_wcnt[RF_WARPS], s_woff[RF_WARPS + 1];
    __shared__ u64 s_base;
    __shared__ u32 s_pre[PUSH_SMEM_WORDS];     // output bits of the folded hop
    __shared__ u32 s_child[PUSH_SMEM_WORDS];   // its child mask
    __shared__ int32_t s_tgt[RF_PRE_EDGES];    // its targets
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_ticket = atomicAdd(&P.counters[0], 1u);
        s_overflow = 0;
    }
    __syncthreads();
    const u32 vcta = s_ticket;
    RF_STAMP(0);

    // Ordering against the kernel in front of this one (P.pdl: launched as a programmatic dependent of it, typically the
    // string scan that produces the mask the folded hop reads): phase A reads only the root's own predicate columns, so
    // it may run while that kernel drains; everything it produced is touched only after pdl_wait() below.  If the root
    // already carries bits from earlier launches (in_bits), phase A does depend on them: wait first.
    if (P.in_bits != nullptr) pdl_wait();

    // ======================= phase A: stream, test, store mask words, list the survivors =======================
    const int64_t vwarp = (int64_t)vcta * RF_WARPS + warp;
    const int64_t c_lo = vwarp * P.chunks_per_warp;
    const int64_t c_hi = (c_lo + P.chunks_per_warp) < P.n_chunks ? (c_lo + P.chunks_per_warp) : P.n_chunks;
    const int stride = 1 + P.ng;  // list entry: [row | leaf row of chain 0 | leaf row of chain 1]
    u32* my_list = P.lists + (size_t)vwarp * P.list_cap * stride;
    u32 cnt = 0;
    for (int64_t c = c_lo; c < c_hi; ++c) {
        const int64_t wbase = c * SR_WARP_ROWS;
        u32 nib[SR_V];
        if (wbase + SR_WARP_ROWS <= P.n) {
            int4 v[NP][SR_V];
#pragma unroll
            for (int p = 0; p < NP; ++p)
#pragma unroll
                for (int j = 0; j < SR_V; ++j) v[p][j] = ldg_stream_v4(P.pred[p].col + wbase + j * 128 + lane * 4);
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 m = 0xFu;
#pragma unroll
                for (int p = 0; p < NP; ++p) m &= range4(v[p][j], P.pred[p].lo, P.pred[p].span);
                nib[j] = m;
            }
        } else {
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 m = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t r = wbase + j * 128 + lane * 4 + e;
                    bool ok = r < P.n;
#pragma unroll
                    for (int p = 0; p < NP; ++p)
                        if (r < P.n) ok = ok && (u32)(P.pred[p].col[r] - P.pred[p].lo) <= P.pred[p].span;
                    m |= ok ? (1u << e) : 0u;
                }
                nib[j] = m;
            }
        }
        if (P.in_bits != nullptr) {
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                const u32 w = P.in_bits[((wbase + j * 128) >> 5) + (lane >> 3)];
                nib[j] &= (w >> ((lane & 7) * 4)) & 0xFu;
            }
        }
        // mask words: same packing as scan_rows (64 B per warp and chunk)
        {
            u32 y[SR_V];
#pragma unroll
            for (int j = 0; j < SR_V; ++j) {
                u32 x = nib[j] << ((lane & 7) * 4);
                x |= __shfl_xor_sync(FULL_MASK, x, 1);
                x |= __shfl_xor_sync(FULL_MASK, x, 2);
                x |= __shfl_xor_sync(FULL_MASK, x, 4);
                y[j] = __shfl_sync(FULL_MASK, x, (lane & 3) * 8);
            }
            u32 out = y[0];
#pragma unroll
            for (int j = 1; j < SR_V; ++j) out = ((lane >> 2) == j) ? y[j] : out;
            if (lane < 4 * SR_V) P.bits[(wbase >> 5) + lane] = out;
        }
        // survivors, in row order: vector j, then lane, then element
        if (__ballot_sync(FULL_MASK, (nib[0] | nib[1] | nib[2] | nib[3]) != 0) == 0) continue;
#pragma unroll
        for (int j = 0; j < SR_V; ++j) {
            const u32 m = nib[j];
            const u32 mine = __popc(m);
            if (__ballot_sync(FULL_MASK, mine != 0) == 0) continue;
            u32 incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 t = __shfl_up_sync(FULL_MASK, incl, d);
                if (lane >= d) incl += t;
            }
            const u32 tot = __shfl_sync(FULL_MASK, incl, 31);
            if (cnt + tot <= (u32)P.list_cap) {
                u32 pos = cnt + incl - mine, mm = m;
                while (mm) {
                    const int e = __ffs(mm) - 1;
                    mm &= mm - 1;
                    my_list[(pos++) * stride] = (u32)(wbase + j * 128 + lane * 4 + e);
                }
            }
            cnt += tot;  // keeps counting past the cap: cnt > list_cap marks the overflow
        }
    }
    // The warp walks the key levels of its OWN candidates' chains right away (RF_ILP candidates per lane, level by level)
    // and leaves each chain's leaf row next to the candidate: these dependent DRAM reads overlap the other warps' and
    // CTAs' streaming instead of sitting on the kernel's tail (r02 timeline: 7-13 us on the critical CTA).  Only the
    // final bit test -- which needs the exchanged mask -- is left for phase B.
    if (P.ng > 0 && cnt <= (u32)P.list_cap) {
        __syncwarp();
        for (u32 i0 = 0; i0 < cnt; i0 += 32 * RF_ILP) {
            u32 row[RF_ILP];
            bool ok[RF_ILP];
#pragma unroll
            for (int k = 0; k < RF_ILP; ++k) {
                const u32 i = i0 + k * 32 + lane;
                ok[k] = i < cnt;
                row[k] = ok[k] ? my_list[(size_t)i * stride] : 0u;
            }
            for (int g = 0; g < P.ng; ++g) {
                const GatherD& G = P.gather[g];
                int64_t r[RF_ILP];
                bool pass[RF_ILP];
#pragma unroll
                for (int k = 0; k < RF_ILP; ++k) { r[k] = row[k]; pass[k] = ok[k]; }
                for (int d = 0; d < G.depth; ++d) {
                    int32_t t[RF_ILP];
#pragma unroll
                    for (int k = 0; k < RF_ILP; ++k) t[k] = pass[k] ? G.fk[d][r[k]] : 0;
#pragma unroll
                    for (int k = 0; k < RF_ILP; ++k) {
                        if (pass[k] && (t[k] < 0 || t[k] >= G.n[d])) {
                            if (t[k] != -1 && G.oob != nullptr) *G.oob = 1u;
                            pass[k] = false;
                        }
                        r[k] = t[k];
                    }
                }
#pragma unroll
                for (int k = 0; k < RF_ILP; ++k)
                    if (ok[k]) my_list[(size_t)(i0 + k * 32 + lane) * stride + 1 + g] = pass[k] ? (u32)r[k] : RF_NO_LEAF;
            }
        }
    }
    if (lane == 0) {
        s_wcnt[warp] = cnt;
        if (cnt > (u32)P.list_cap) s_overflow = 1;
    }
    // from here on the kernel reads what earlier launches produced (the mask behind the folded hop, chain bitmaps)
    pdl_wait();
    // the next execution's first scan may move onto the SMs this kernel leaves (it orders itself with its own pdl_wait);
    // triggering only now keeps the chain simple: whatever starts early starts after the kernel in front of this one is done
    pdl_launch_dependents();

    RF_STAMP(1);
    // ======================= phase B =======================
    // ---- (1) the folded hop (multi-GPU: the COLLECT of the mask exchange first -- the wait for the slowest rank has been
    //      hiding behind phase A)
    if (P.pre.n > 0) rf_run_pre(P.pre, vcta, s_pre, s_child, s_tgt);
    __syncthreads();  // s_wcnt, s_overflow, s_pre
    RF_STAMP(2);
    if (tid == 0) {
        u32 o = 0;
        for (int w = 0; w < RF_WARPS; ++w) {
            s_woff[w] = o;
            o += s_wcnt[w];
        }
        s_woff[RF_WARPS] = o;
    }
    __syncthreads();
    const bool dense = s_overflow != 0;
    const u32 n_c = s_woff[RF_WARPS];
    const int64_t w_lo = (int64_t)vcta * RF_WARPS * P.chunks_per_warp * (SR_WARP_ROWS / 32);  // this CTA's mask words
    const int64_t w_end = ((int64_t)vcta + 1) * RF_WARPS * P.chunks_per_warp * (SR_WARP_ROWS / 32);
    const int64_t n_words = (P.n + 31) >> 5;
    const int64_t w_hi = w_end < n_words ? w_end : n_words;
    const u32* gbits[CF_MAX_GATHER];
#pragma unroll
    for (int g = 0; g < CF_MAX_GATHER; ++g) gbits[g] = ((P.pre_mask >> g) & 1u) ? s_pre : P.gather[g].bits;

    // ---- (2) chains of the surviving rows; count what is left
    u32 kept = 0;
    if (!dense) {
        if (P.ng == 0) {
            kept = tid == 0 ? n_c : 0;
        } else {
            // the key levels were walked in phase A: test the leaf bits (shared memory for a folded hop)
            for (u32 i = tid; i < n_c; i += RF_THREADS) {
                int w = 0;
#pragma unroll
                for (int k = 1; k < RF_WARPS; ++k) w += (i >= s_woff[k]) ? 1 : 0;
                u32* e = P.lists + (((size_t)vcta * RF_WARPS + w) * P.list_cap + (i - s_woff[w])) * stride;
                bool ok = true;
                for (int g = 0; g < P.ng; ++g) {
                    const u32 leaf = __ldcg(e + 1 + g);
                    ok = ok && leaf != RF_NO_LEAF && (gbits[g] == nullptr || bit_test(gbits[g], leaf));
                }
                if (ok) ++kept;
                else {
                    const u32 row = __ldcg(e);
                    atomicAnd(&P.bits[row >> 5], ~(1u << (row & 31)));
                    *e = row | 0x80000000u;
                }
            }
        }
    } else {
        for (int64_t w0 = w_lo + (int64_t)tid * 4; w0 < w_hi; w0 += RF_THREADS * 4) {
            uint4 v = __ldcg(reinterpret_cast<const uint4*>(P.bits + w0));
            u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (w0 + j >= w_hi) { w[j] = 0; continue; }
                u32 m = w[j], keep = w[j];
                if (P.ng > 0) {
                    const int64_t rb = (w0 + j) << 5;
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        bool ok = true;
                        for (int g = 0; g < P.ng; ++g) ok = ok && rf_chain(P.gather[g], gbits[g], rb + b);
                        if (!ok) keep &= ~(1u << b);
                    }
                    if (keep != w[j]) P.bits[w0 + j] = keep;
                }
                kept += __popc(keep);
            }
        }
    }
    u32 cta_count;
    block_exclusive_scan(kept, s_warp, cta_count);
    RF_STAMP(3);

    // ---- (3) publish, then sum the counts of every lower ticket
    if (tid == 0) st_volatile_u64(P.cta_state + vcta, ((u64)P.epoch << 32) | cta_count);
    u64 part = 0;
    for (u32 j = tid; j < vcta; j += RF_THREADS) {
        u64 st;
        do { st = ld_volatile_u64(P.cta_state + j); } while ((u32)(st >> 32) != P.epoch);
        part += (u32)st;
    }
    // block reduction of a u64 (the grand total of a 2^31-row table fits, a u32 lane sum might not)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(FULL_MASK, part, d);
    __shared__ u64 s_part[RF_WARPS];
    __syncthreads();  // block_exclusive_scan's readers of s_warp are done; s_part is fresh
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    if (tid == 0) {
        u64 b = 0;
        for (int w = 0; w < RF_WARPS; ++w) b += s_part[w];
        s_base = b;
    }
    __syncthreads();
    const u64 base = s_base;
    RF_STAMP(4);

    // ---- (4) ordered write
    const bool gather = P.pg.n_ranks > 0;
    if (!dense) {
        u64 running = base;
        for (u32 i0 = 0; i0 < n_c; i0 += RF_THREADS) {
            const u32 i = i0 + tid;
            u32 row = 0x80000000u;
            if (i < n_c) {
                int w = 0;
#pragma unroll
                for (int k = 1; k < RF_WARPS; ++k) w += (i >= s_woff[k]) ? 1 : 0;
                row = __ldcg(P.lists + (((size_t)vcta * RF_WARPS + w) * P.list_cap + (i - s_woff[w])) * stride);
            }
            const u32 valid = (row & 0x80000000u) ? 0u : 1u;
            u32 tot;
            const u32 ex = block_exclusive_scan(valid, s_warp, tot);
            if (valid) {
                const int64_t pos = (int64_t)(running + ex);
                const int32_t v = (int32_t)(P.row_base + row);
                if (pos < P.capacity) P.out_idx[pos] = v;
                if (gather) gather_store(P.pg, pos, v);
            }
            running += tot;
            __syncthreads();
        }
    } else {
        u64 running = base;
        for (int64_t t0 = w_lo; t0 < w_hi; t0 += RF_THREADS * 4) {
            const int64_t w0 = t0 + (int64_t)tid * 4;
            u32 w[4] = {0, 0, 0, 0};
            if (w0 < w_hi) {
                const uint4 v = __ldcg(reinterpret_cast<const uint4*>(P.bits + w0));
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (w0 + j >= w_hi) w[j] = 0;
            }
            const u32 c = __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
            u32 tot;
            const u32 ex = block_exclusive_scan(c, s_warp, tot);
            int64_t pos = (int64_t)(running + ex);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                u32 m = w[j];
                const int64_t rb = P.row_base + ((w0 + j) << 5);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    if (pos < P.capacity) P.out_idx[pos] = (int32_t)(rb + b);
                    if (gather) gather_store(P.pg, pos, (int32_t)(rb + b));
                    ++pos;
                }
            }
            running += tot;
            __syncthreads();
        }
    }

    RF_STAMP(5);
    // ---- tail: the highest ticket knows the grand total; the last CTA to finish re-arms the counters
    if (tid == 0 && vcta == gridDim.x - 1) *P.total = base + cta_count;
    // the total reaches the last CTA through the finished-CTA counter (barrier, then thread 0's fence, then the atomic); the
    // gathered index words carry their own epoch tags, so no system-scope fence is needed for them
    __syncthreads();
    if (tid == 0) __threadfence();
    if (tid == 0) s_last = (atomicAdd(&P.counters[1], 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    RF_STAMP(6);
    if (s_last) {
        if (tid == 0) {
            P.counters[0] = 0;
            P.counters[1] = 0;
        }
        // every CTA has read the push target (before its arrival at the counter above): leave it zeroed
        if (P.clean != nullptr)
            for (int w = tid; w < P.clean_words; w += RF_THREADS) P.clean[w] = 0;
        if (gather) gather_tail(P.pg, P.total);
    }
    RF_STAMP(7);
}

// ---------------------------------------------------------------------------------------------
// K4+K3  root_finish: the latency half of the root node as a launch of its own (COLQ_OPT_ROOT_FUSED=2).
//
// The bandwidth half is the ordinary NON-persistent scan_rows<NP, 0, false, LIST> -- 71 k short-lived CTAs walking the
// column in address order at 8 CTAs/SM and 32 registers, the access pattern that reaches 7.0 TB/s, where the persistent
// root_fused_kernel's phase A reaches 6.4 (r02 timelines) -- which additionally leaves every 512-row chunk's survivors in
// lists[chunk][list_cap] and their number in ucount[chunk].  This kernel is the rest: ticket v takes the chunks
// [v * units_per_cta, ...), builds the exclusive prefix of their counts in shared memory, walks the chains of its ~800
// candidates RF_ILP per thread (failed bits cleared in the mask with atomicAnd), publishes its survivor count, sums the counts
// of all lower tickets (decoupled look-back), and writes the indices in order -- locally and, multi-GPU, as flag-in-data
// words into every peer's slot.  The folded to-many hop, the mask COLLECT, the dense fallback (a chunk whose list
// overflowed: this CTA walks its part of the mask word by word) and the tail are those of root_fused_kernel.
// r01's compact_fused did this work in 57.6 us by reading the dense 37 MB mask twice across a grid barrier.
// ---------------------------------------------------------------------------------------------
constexpr int RF_MAX_UPC = 2048;  // chunks one CTA finishes (shared-memory prefix array)

__global__ void __launch_bounds__(RF_THREADS, 4) root_finish_kernel(const RootFusedParams P) {
    __shared__ u32 s_warp[33];
    __shared__ u32 s_ticket, s_last, s_overflow;
    __shared__ u32 s_uoff[RF_MAX_UPC + 1];     // exclusive prefix of this CTA's chunk counts
    __shared__ u64 s_base;
    __shared__ u64 s_part[RF_WARPS];
    __shared__ u32 s_pre[PUSH_SMEM_WORDS];
    __shared__ u32 s_child[PUSH_SMEM_WORDS];
    __shared__ int32_t s_tgt[RF_PRE_EDGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_ticket = atomicAdd(&P.counters[0], 1u);
        s_overflow = 0;
    }
    pdl_wait();  // everything below was produced by earlier launches (lists, counts, mask words, the exchanged mask)
    __syncthreads();
    const u32 vcta = s_ticket;
    RF_STAMP(0);
    if (P.pre.n > 0) rf_run_pre(P.pre, vcta, s_pre, s_child, s_tgt);
    RF_STAMP(1);

    // ---- this CTA's chunks and the exclusive prefix of their candidate counts
    const int64_t u_lo = (int64_t)vcta * P.units_per_cta;
    const int64_t u_end = u_lo + P.units_per_cta;
    const int64_t u_hi = u_end < P.n_chunks ? u_end : P.n_chunks;
    const int nu = u_hi > u_lo ? (int)(u_hi - u_lo) : 0;
    {
        u32 carry = 0;
        for (int i0 = 0; i0 < nu; i0 += RF_THREADS) {
            const int i = i0 + tid;
            u32 c = 0;
            if (i < nu) {
                c = __ldcg(P.ucount + u_lo + i);
                if (c > (u32)P.list_cap) {
                    s_overflow = 1;
                    c = 0;
                }
            }
            u32 tot;
            const u32 ex = block_exclusive_scan(c, s_warp, tot);
            if (i < nu) s_uoff[i] = carry + ex;
            carry += tot;
            __syncthreads();
        }
        if (tid == 0) s_uoff[nu] = carry;
    }
    __syncthreads();  // s_uoff, s_overflow, s_pre
    RF_STAMP(2);
    const bool dense = s_overflow != 0;
    const u32 n_c = s_uoff[nu];
    const int64_t words_per_unit = SR_WARP_ROWS / 32;
    const int64_t n_words = (P.n + 31) >> 5;
    const int64_t w_lo = u_lo * words_per_unit;  // this CTA's mask words
    const int64_t w_end = u_end * words_per_unit;
    const int64_t w_hi = w_end < n_words ? w_end : n_words;
    const u32* gbits[CF_MAX_GATHER];
#pragma unroll
    for (int g = 0; g < CF_MAX_GATHER; ++g) gbits[g] = ((P.pre_mask >> g) & 1u) ? s_pre : P.gather[g].bits;
    // candidate i of this CTA -> its list entry (the chunk is found by bisection of the shared prefix array)
    auto entry = [&](u32 i) -> u32* {
        int lo = 0, hi = nu;  // s_uoff[lo] <= i < s_uoff[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_uoff[mid] <= i) lo = mid;
            else hi = mid;
        }
        return P.lists + ((size_t)u_lo + lo) * P.list_cap + (i - s_uoff[lo]);
    };

    // ---- chains of the candidates; count what is left
    u32 kept = 0;
    if (!dense) {
        if (P.ng == 0) {
            kept = tid == 0 ? n_c : 0;
        } else {
            // RF_ILP candidates per thread at a time, walked level by level: the dependent loads of one chain are
            // serial, those of different candidates are all in flight together
            for (u32 i0 = 0; i0 < n_c; i0 += RF_THREADS * RF_ILP) {
                u32* ep[RF_ILP];
                u32 row[RF_ILP];
                bool ok[RF_ILP], pass[RF_ILP];
#pragma unroll
                for (int k = 0; k < RF_ILP; ++k) {
                    const u32 i = i0 + k * RF_THREADS + tid;
                    ok[k] = i < n_c;
                    ep[k] = ok[k] ? entry(i) : P.lists;
                    row[k] = ok[k] ? __ldcg(ep[k]) : 0u;
                    pass[k] = ok[k];
                }
                for (int g = 0; g < P.ng; ++g) {
                    const GatherD& G = P.gather[g];
                    int64_t r[RF_ILP];
#pragma unroll
                    for (int k = 0; k < RF_ILP; ++k) r[k] = row[k];
                    for (int d = 0; d < G.depth; ++d) {
                        int32_t t[RF_ILP];
#pragma unroll
                        for (int k = 0; k < RF_ILP; ++k) t[k] = pass[k] ? G.fk[d][r[k]] : 0;
#pragma unroll
                        for (int k = 0; k < RF_ILP; ++k) {
                            if (pass[k] && (t[k] < 0 || t[k] >= G.n[d])) {
                                if (t[k] != -1 && G.oob != nullptr) *G.oob = 1u;
                                pass[k] = false;
                            }
                            r[k] = t[k];
                        }
                    }
                    if (gbits[g] != nullptr) {
#pragma unroll
                        for (int k = 0; k < RF_ILP; ++k) pass[k] = pass[k] && bit_test(gbits[g], pass[k] ? r[k] : 0);
                    }
                }
#pragma unroll
                for (int k = 0; k < RF_ILP; ++k) {
                    if (!ok[k]) continue;
                    if (pass[k]) ++kept;
                    else {
                        atomicAnd(&P.bits[row[k] >> 5], ~(1u << (row[k] & 31)));
                        *ep[k] = row[k] | 0x80000000u;
                    }
                }
            }
        }
    } else {
        for (int64_t w0 = w_lo + (int64_t)tid * 4; w0 < w_hi; w0 += RF_THREADS * 4) {
            uint4 v = __ldcg(reinterpret_cast<const uint4*>(P.bits + w0));
            u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (w0 + j >= w_hi) { w[j] = 0; continue; }
                u32 m = w[j], keep = w[j];
                if (P.ng > 0) {
                    const int64_t rb = (w0 + j) << 5;
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        bool ok = true;
                        for (int g = 0; g < P.ng; ++g) ok = ok && rf_chain(P.gather[g], gbits[g], rb + b);
                        if (!ok) keep &= ~(1u << b);
                    }
                    if (keep != w[j]) P.bits[w0 + j] = keep;
                }
                kept += __popc(keep);
            }
        }
    }
    u32 cta_count;
    block_exclusive_scan(kept, s_warp, cta_count);
    RF_STAMP(3);

    // ---- publish, then sum the counts of every lower ticket
    if (tid == 0) st_volatile_u64(P.cta_state + vcta, ((u64)P.epoch << 32) | cta_count);
    u64 part = 0;
    for (u32 j = tid; j < vcta; j += RF_THREADS) {
        u64 st;
        do { st = ld_volatile_u64(P.cta_state + j); } while ((u32)(st >> 32) != P.epoch);
        part += (u32)st;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(FULL_MASK, part, d);
    __syncthreads();
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    if (tid == 0) {
        u64 b = 0;
        for (int w = 0; w < RF_WARPS; ++w) b += s_part[w];
        s_base = b;
    }
    __syncthreads();
    const u64 base = s_base;
    RF_STAMP(4);

    // ---- ordered write
    const bool gather = P.pg.n_ranks > 0;
    if (!dense) {
        u64 running = base;
        for (u32 i0 = 0; i0 < n_c; i0 += RF_THREADS) {
            const u32 i = i0 + tid;
            u32 row = 0x80000000u;
            if (i < n_c) row = __ldcg(entry(i));
            const u32 valid = (row & 0x80000000u) ? 0u : 1u;
            u32 tot;
            const u32 ex = block_exclusive_scan(valid, s_warp, tot);
            if (valid) {
                const int64_t pos = (int64_t)(running + ex);
                const int32_t v = (int32_t)(P.row_base + row);
                if (pos < P.capacity) P.out_idx[pos] = v;
                if (gather) gather_store(P.pg, pos, v);
            }
            running += tot;
            __syncthreads();
        }
    } else {
        u64 running = base;
        for (int64_t t0 = w_lo; t0 < w_hi; t0 += RF_THREADS * 4) {
            const int64_t w0 = t0 + (int64_t)tid * 4;
            u32 w[4] = {0, 0, 0, 0};
            if (w0 < w_hi) {
                const uint4 v = __ldcg(reinterpret_cast<const uint4*>(P.bits + w0));
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (w0 + j >= w_hi) w[j] = 0;
            }
            const u32 c = __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
            u32 tot;
            const u32 ex = block_exclusive_scan(c, s_warp, tot);
            int64_t pos = (int64_t)(running + ex);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                u32 m = w[j];
                const int64_t rb = P.row_base + ((w0 + j) << 5);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    if (pos < P.capacity) P.out_idx[pos] = (int32_t)(rb + b);
                    if (gather) gather_store(P.pg, pos, (int32_t)(rb + b));
                    ++pos;
                }
            }
            running += tot;
            __syncthreads();
        }
    }
    RF_STAMP(5);

    // ---- tail: the highest ticket knows the grand total; the last CTA to finish re-arms the counters
    if (tid == 0 && vcta == gridDim.x - 1) *P.total = base + cta_count;
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = (atomicAdd(&P.counters[1], 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    RF_STAMP(6);
    if (s_last) {
        if (tid == 0) {
            P.counters[0] = 0;
            P.counters[1] = 0;
        }
        // every CTA has read the push target (before its arrival at the counter above): leave it zeroed
        if (P.clean != nullptr)
            for (int w = tid; w < P.clean_words; w += RF_THREADS) P.clean[w] = 0;
        if (gather) gather_tail(P.pg, P.total);
    }
    RF_STAMP(7);
}

// ---------------------------------------------------------------------------------------------
// K5  result materialisation: the value half of Table.subset(BitSet) (M/InMemoryTable.java:106-159)
//
// The reference re-scans every column for the set bits (hot loop 3: `for i in [0, size): if bits.get(i) copy`).  Here
// the compacted ascending index list already exists in HBM, so a column of the result is a gather: O(matches), not
// O(rows).  Int / boolean / to-one association columns (indices un-remapped, :143-154) are plain gathers; string
// columns (plain or dictionary-encoded) take lengths -> exclusive scan -> byte copy; to-many association columns the
// same over (offsets, targets).
// ---------------------------------------------------------------------------------------------

template <typename T>
__global__ void __launch_bounds__(256) gather_values_kernel(const T* col, const int32_t* idx, int64_t row_base, int64_t n, T* out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = col[(int64_t)idx[i] - row_base];
}

// dictionary-encoded int column: row -> code -> distinct value
__global__ void __launch_bounds__(256) gather_decode_kernel(const int32_t* codes, const int32_t* dict, const int32_t* idx, int64_t row_base,
                                                           int64_t n, int32_t* out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = dict[codes[(int64_t)idx[i] - row_base]];
}

// variable-length columns: element e of the source spans src_off[e] .. src_off[e + 1] (OffT = u32 for strings, int64
// for CSR associations); `codes` (nullable) maps a row to its dictionary entry
template <typename OffT>
struct GatherVarParams {
    const OffT* src_off;
    const int32_t* codes;
    const int32_t* idx;
    int64_t row_base;
    int64_t n;
    u64* out_off;  // n + 1 entries: lengths first, then (after scan_u64_inplace_kernel) exclusive offsets
};

template <typename OffT>
__global__ void __launch_bounds__(256) gather_var_lens_kernel(const GatherVarParams<OffT> P) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        int64_t e = (int64_t)P.idx[i] - P.row_base;
        if (P.codes != nullptr) e = P.codes[e];
        P.out_off[i + 1] = (u64)(P.src_off[e + 1] - P.src_off[e]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) P.out_off[0] = 0;
}

// single block: a[1..n] becomes its inclusive prefix sum (a[0] = 0 stays), i.e. a[i] = start of element i
__global__ void __launch_bounds__(1024) scan_u64_inplace_kernel(u64* a, int64_t n) {
    __shared__ u64 s_warp[33];
    __shared__ u64 s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 1; base <= n; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        u64 v = i <= n ? a[i] : 0, incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 t = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            u64 w = s_warp[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 t = __shfl_up_sync(FULL_MASK, wi, d);
                if (lane >= d) wi += t;
            }
            s_warp[lane] = wi - w;
            if (lane == 31) s_warp[32] = wi;
        }
        __syncthreads();
        const u64 carry = s_carry;
        if (i <= n) a[i] = carry + s_warp[warp] + incl;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_warp[32];
        __syncthreads();
    }
}

// copy the payloads: ELEM = 1 (string bytes) or 4 (CSR targets).  One lane per result row for short rows; rows
// longer than 64 elements are copied by the whole warp.
template <typename OffT, typename ElemT>
__global__ void __launch_bounds__(256) gather_var_copy_kernel(const GatherVarParams<OffT> P, const ElemT* src, ElemT* out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t n_groups = (P.n + 31) >> 5;
    for (int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < n_groups; g += warp_stride) {
        const int64_t i = g * 32 + lane;
        u64 s0 = 0, d0 = 0, len = 0;
        if (i < P.n) {
            int64_t e = (int64_t)P.idx[i] - P.row_base;
            if (P.codes != nullptr) e = P.codes[e];
            s0 = (u64)P.src_off[e];
            len = (u64)P.src_off[e + 1] - s0;
            d0 = P.out_off[i];
        }
        const bool is_long = len > 64;
        if (!is_long)
            for (u64 b = 0; b < len; ++b) out[d0 + b] = src[s0 + b];
        u32 longs = __ballot_sync(FULL_MASK, is_long);
        while (longs) {
            const int l = __ffs(longs) - 1;
            longs &= longs - 1;
            const u64 ls = __shfl_sync(FULL_MASK, s0, l), ld = __shfl_sync(FULL_MASK, d0, l), ll = __shfl_sync(FULL_MASK, len, l);
            for (u64 b = lane; b < ll; b += 32) out[ld + b] = src[ls + b];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3''  single-PASS compaction with decoupled look-back (COLQ_OPT_FUSED_COMPACT=2).  One CTA per 131072-row tile, tile ids handed out
// by an atomic counter in CTA start order, so a CTA only ever waits for tiles whose CTAs are already running.  Each CTA:
// loads its 16 KB of mask once, (NG > 0) resolves the root's deferred FK chains block-wide exactly like
// compact_fused_kernel, publishes its count, finds its output offset by looking back over its predecessors' published
// {aggregate | inclusive prefix} words (32 at a time, one warp), and writes its indices from registers.  Against the
// cooperative two-phase kernel: no grid barrier, no second read of the mask, an ordinary (non-cooperative) launch.
// Tile states carry a per-launch epoch, so nothing has to be cleared between launches.
// MEASURED SLOWER than the cooperative kernel on B200 (r01: 74 vs 62 us at 293.5 M rows with chains, 125 vs 94 us at
// 1 B rows; tile sizes of 64 K / 128 K / 256 K rows tried): with one 16 KB tile per CTA the launch is bound by CTA
// lifetime x waves, not by bandwidth.  Kept selectable and parity-tested; not the default.
// ---------------------------------------------------------------------------------------------

struct CompactLookbackParams {
    u32* bits;
    int64_t n_words;
    int64_t n_tiles;    // == gridDim.x
    u64* tile_state;    // [n_tiles]: epoch << 34 | flag << 32 | value; flag 1 = tile count, 2 = inclusive prefix
    u32* counters;      // [0] next tile id, [1] finished CTAs; both return to zero at the end of the launch
    u32 epoch;          // 1 .. 2^30 - 1, different from the previous launch on the same tile_state
    u64* total;
    int32_t* out_idx;
    int64_t capacity;
    int64_t row_base;
    int64_t n_rows;
    GatherD gather[CF_MAX_GATHER];
};


template <int NG>
__global__ void __launch_bounds__(CP_THREADS) compact_lookback_kernel(const CompactLookbackParams P) {
    __shared__ u32 s_warp[33];
    __shared__ u32 s_list[NG > 0 ? CF_LIST_CAP : 1];
    __shared__ u32 s_tile;
    __shared__ u64 s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(&P.counters[0], 1u);
    __syncthreads();
    const int64_t t = s_tile;
    const int64_t tile_w0 = t * CF_WORDS_PER_TILE;
    const u32 lw0 = threadIdx.x * 4 * CF_VEC;  // first word of this thread inside the tile

    uint4 v[CF_VEC];
    u32 c = 0;
#pragma unroll
    for (int k = 0; k < CF_VEC; ++k) {
        v[k] = load_words4(P.bits, tile_w0 + lw0 + 4 * k, P.n_words);
        c += __popc(v[k].x) + __popc(v[k].y) + __popc(v[k].z) + __popc(v[k].w);
    }

    if (NG > 0) {
        u32 total;
        const u32 ex = block_exclusive_scan(c, s_warp, total);
        if (total != 0) {
            const bool listed = total <= CF_LIST_CAP;
            if (listed) {
                // every surviving row of the tile into one shared list, walked round-robin by the whole CTA
                u32 pos = ex;
#pragma unroll
                for (int k = 0; k < CF_VEC; ++k) {
                    const u32 w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        u32 m = w[j];
                        while (m) {
                            const int b = __ffs(m) - 1;
                            m &= m - 1;
                            s_list[pos++] = ((lw0 + 4 * k + j) << 5) + b;
                        }
                    }
                }
                __syncthreads();
                for (u32 i = threadIdx.x; i < total; i += CP_THREADS) {
                    const u32 lr = s_list[i];
                    const int64_t row = (tile_w0 << 5) + lr;
                    bool ok = row < P.n_rows;
#pragma unroll
                    for (int g = 0; g < NG; ++g) ok = ok && gather_eval(P.gather[g], row, 0);
                    if (!ok) s_list[i] = lr | 0x80000000u;
                }
                __syncthreads();
            }
            // back to the owner of each word: drop the failed bits (list verdicts, or walk the chains here when the
            // tile is too dense for the list), and write changed words back -- the root mask is part of the result
            u32 pos = ex;
            c = 0;
#pragma unroll
            for (int k = 0; k < CF_VEC; ++k) {
                u32 w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    u32 m = w[j], keep = w[j];
                    const int64_t rb = (tile_w0 + lw0 + 4 * k + j) << 5;
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        bool ok;
                        if (listed) ok = !(s_list[pos++] & 0x80000000u);
                        else {
                            ok = rb + b < P.n_rows;
#pragma unroll
                            for (int g = 0; g < NG; ++g) ok = ok && gather_eval(P.gather[g], rb + b, 0);
                        }
                        if (!ok) keep &= ~(1u << b);
                    }
                    if (keep != w[j]) P.bits[tile_w0 + lw0 + 4 * k + j] = keep;
                    w[j] = keep;
                    c += __popc(keep);
                }
                v[k] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        __syncthreads();  // s_warp is reused by the next scan
    }

    u32 tile_total;
    const u32 ex = block_exclusive_scan(c, s_warp, tile_total);

    // ---- decoupled look-back (warp 0)
    if (warp == 0) {
        const u64 tag = (u64)P.epoch << 34;
        if (lane == 0) {
            __threadfence();
            st_volatile_u64(P.tile_state + t, tag | ((u64)(t == 0 ? 2 : 1) << 32) | tile_total);
        }
        u64 base = 0;
        if (t > 0) {
            int64_t hi = t - 1;  // newest predecessor of the current window
            while (true) {
                const int64_t idx = hi - lane;
                u64 st;
                bool valid;
                do {
                    st = idx >= 0 ? ld_volatile_u64(P.tile_state + idx) : (tag | (2ull << 32));  // before tile 0: prefix 0
                    valid = (st >> 34) == (u64)P.epoch && ((st >> 32) & 3u) != 0;
                } while (!__all_sync(FULL_MASK, valid));
                const u32 flag = (u32)(st >> 32) & 3u, val = (u32)st;
                const u32 has_prefix = __ballot_sync(FULL_MASK, flag == 2);
                const int first = has_prefix ? __ffs(has_prefix) - 1 : 31;  // nearest tile that knows its inclusive prefix
                u32 contrib = lane <= first ? val : 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(FULL_MASK, contrib, d);
                base += contrib;
                if (has_prefix) break;
                hi -= 32;
            }
            if (lane == 0) {
                __threadfence();
                st_volatile_u64(P.tile_state + t, tag | (2ull << 32) | (u64)(u32)(base + tile_total));
            }
        }
        if (lane == 0) s_base = base;
    }
    __syncthreads();

    // ---- ordered write from registers
    if (c != 0) {
        int64_t pos = (int64_t)s_base + ex;
#pragma unroll
        for (int k = 0; k < CF_VEC; ++k) {
            const u32 w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                u32 m = w[j];
                const int64_t rb = P.row_base + ((tile_w0 + lw0 + 4 * k + j) << 5);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    if (pos < P.capacity) P.out_idx[pos] = (int32_t)(rb + b);
                    ++pos;
                }
            }
        }
    }
    if (threadIdx.x == 0) {
        if (t == P.n_tiles - 1) *P.total = s_base + tile_total;
        const u32 prev = atomicAdd(&P.counters[1], 1u);
        if (prev == gridDim.x - 1) {  // every CTA has taken its tile id long ago: rearm the counters
            P.counters[0] = 0;
            P.counters[1] = 0;
        }
    }
}

// popcount of a whole bitmask into one u64 (node cardinalities; not on the timed path)
__global__ void __launch_bounds__(256) popc_total_kernel(const u32* bits, int64_t n_words, u64* out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    u64 c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) c += __popc(bits[i]);
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(FULL_MASK, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd((unsigned long long*)out, (unsigned long long)c);
}

// string ingest: the largest byte payload of any 1024-row scan tile; sizes the TMA ring slots of scan_str exactly
__global__ void __launch_bounds__(256) tile_payload_max_kernel(const u32* offsets, int64_t n, u32* out_max) {
    const int64_t n_tiles = (n + ST_ROWS - 1) / ST_ROWS;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    u32 m = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += stride) {
        const int64_t r0 = t * ST_ROWS, r1 = (r0 + ST_ROWS) < n ? (r0 + ST_ROWS) : n;
        const u32 b = offsets[r1] - offsets[r0];
        m = b > m ? b : m;
    }
    m = __reduce_max_sync(FULL_MASK, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out_max, m);
}

// association ingest check: min / max of a to-one column (M/InMemoryTable.java:70-71 would NPE on a bad target)
__global__ void __launch_bounds__(256) fk_minmax_kernel(const int32_t* fk, int64_t n, int32_t* out_min, int32_t* out_max) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int32_t lo = INT32_MAX, hi = INT32_MIN;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t v = fk[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
    for (int d = 16; d > 0; d >>= 1) {
        int32_t a = __shfl_down_sync(FULL_MASK, lo, d), b = __shfl_down_sync(FULL_MASK, hi, d);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out_min, lo);
        atomicMax(out_max, hi);
    }
}

}  // namespace colq
