// colq_ingest.cuh -- load-time kernels (the ingest path Table graph -> device).
//
// The reference builds its tables on the host: InMemoryTable.associateTo classifies and transposes Association[]
// objects (M/InMemoryTable.java:44-90), the app copies String[] columns around (app/.../Runner.java:89-196).  Round 1
// replaced those with host loops in the shims (dictionary building, None / One / Many flattening, per-edge range checks).
// Here the flat arrays are shipped as they are and the GPU does the rest:
//   * dictionary encoding of a string column: hash insert with the FIRST row of every distinct value as its
//     representative, a verification pass (bytes really equal -- a 64-bit hash collision restarts with another seed),
//     ordered compaction of the first-occurrence flags, code assignment in first-appearance order (identical to the
//     host's encode_dictionary), gather of the distinct values;
//   * association ingest: degree / order / target-range statistics of a CSR in one pass, the None / One / Many
//     classification (E/ExecutionContext.java:110-118's switch, decided once per column instead of once per row), and
//     the dense to-one form (-1 = None) when every row has at most one target.
#pragma once

#include "colq_kernels.cuh"

namespace colq {

struct DictSlot {
    u64 key;      // 64-bit hash of the value, never 0; 0 = empty slot
    u32 min_row;  // smallest row that holds the value (0xffffffff until the first insert lands)
    u32 code;     // dictionary code = rank of min_row among all representatives
};

// The strings are read as little-endian 32-bit words of the (16-byte aligned, padded) bytes buffer and re-aligned with a
// funnel shift -- 2-3 loads per city name instead of 8.5 single-byte loads.
struct StrWords {
    const u32* words;
    u32 wi, sh, cur;
    __device__ __forceinline__ StrWords(const uint8_t* bytes, u32 o0) : words(reinterpret_cast<const u32*>(bytes)), wi(o0 >> 2), sh((o0 & 3) * 8) {
        cur = words[wi];
    }
    // the next 4 bytes of the string (bytes past `remaining` are zeroed)
    __device__ __forceinline__ u32 next(u32 remaining) {
        const u32 nxt = words[++wi];  // at most 4 bytes past the string: inside the buffer's slack
        u32 w = __funnelshift_r(cur, nxt, sh);
        cur = nxt;
        if (remaining < 4) w &= (1u << (remaining * 8)) - 1u;
        return w;
    }
};

__device__ __forceinline__ u64 hash_bytes(const uint8_t* bytes, u32 o0, u32 len, u64 seed) {
    u64 h = 0xcbf29ce484222325ull ^ seed;
    if (len > 0) {
        StrWords sw(bytes, o0);
        for (u32 c = 0; c < len; c += 4) {
            h ^= sw.next(len - c);
            h *= 0x100000001b3ull;
            h ^= h >> 29;
        }
    }
    h ^= (u64)len * 0x9E3779B97F4A7C15ull;
    h = (h ^ (h >> 30)) * 0xBF58476D1CE4E5B9ull;
    h = (h ^ (h >> 27)) * 0x94D049BB133111EBull;
    h ^= h >> 31;
    return h ? h : 1;
}

__global__ void __launch_bounds__(256) dict_init_kernel(DictSlot* slots, int64_t n_slots) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {
        slots[i].key = 0;
        slots[i].min_row = 0xffffffffu;
        slots[i].code = 0;
    }
}

// status[0]: 1 = table more than half full (the host retries with a larger one), status[1]: 1 = hash collision found
// by the verification pass, status[2]: distinct values inserted so far.  slot_of_row[i] receives the slot that represents row i's value.
__global__ void __launch_bounds__(256) dict_insert_kernel(const u32* offsets, const uint8_t* bytes, int64_t n, DictSlot* slots, u32 mask,
                                                         u64 seed, int32_t* slot_of_row, u32* status) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u32 o0 = offsets[i], len = offsets[i + 1] - o0;
        const u64 h = hash_bytes(bytes, o0, len, seed);
        u32 s = (u32)(h >> 7) & mask;
        u32 probes = 0;
        while (true) {
            u64 k = ld_volatile_u64(&slots[s].key);
            if (k == 0) {
                const u64 prev = atomicCAS((unsigned long long*)&slots[s].key, 0ull, (unsigned long long)h);
                k = prev == 0 ? h : prev;
                // a new distinct value: keep the table at most half full (the host retries with a larger one)
                if (prev == 0 && atomicAdd(&status[2], 1u) >= (mask >> 1)) status[0] = 1;
            }
            if (k == h) {
                // the value's representative is its FIRST row; rows arrive roughly in order, so after the first few
                // inserts this is a plain (cached) read and no atomic
                if (ld_volatile_u32(&slots[s].min_row) > (u32)i) atomicMin(&slots[s].min_row, (u32)i);
                slot_of_row[i] = (int32_t)s;
                break;
            }
            s = (s + 1) & mask;
            if (++probes > mask || ld_volatile_u32(&status[0]) != 0) {
                status[0] = 1;
                slot_of_row[i] = 0;
                break;
            }
        }
    }
}

// every row compares its bytes with its representative's (two values with one 64-bit hash must not share a code) and
// the representatives raise their first-occurrence flag; one bitmask word per warp
__global__ void __launch_bounds__(256) dict_verify_kernel(const u32* offsets, const uint8_t* bytes, int64_t n, const DictSlot* slots,
                                                         const int32_t* slot_of_row, u32* first_bits, u32* status) {
    if (ld_volatile_u32(&status[0]) != 0) return;  // the insert pass gave up (table too small): slot_of_row is not meaningful
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_pad = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
        bool first = false;
        if (i < n) {
            const u32 rep = slots[slot_of_row[i]].min_row;
            first = rep == (u32)i;
            if (!first) {
                const u32 a0 = offsets[i], la = offsets[i + 1] - a0;
                const u32 b0 = offsets[rep], lb = offsets[rep + 1] - b0;
                bool same = la == lb;
                if (same && la > 0) {
                    StrWords wa(bytes, a0), wb(bytes, b0);
                    for (u32 k = 0; same && k < la; k += 4) same = wa.next(la - k) == wb.next(la - k);
                }
                if (!same) status[1] = 1;
            }
        }
        const u32 word = __ballot_sync(FULL_MASK, first);
        if ((threadIdx.x & 31) == 0) first_bits[i >> 5] = word;
    }
}

// representative k (ascending row order = first-appearance order) gets code k
__global__ void __launch_bounds__(256) dict_assign_kernel(const int32_t* first_rows, int64_t n_dict, const int32_t* slot_of_row, DictSlot* slots) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_dict; k += stride) slots[slot_of_row[first_rows[k]]].code = (u32)k;
}

// in place: slot index -> dictionary code
__global__ void __launch_bounds__(256) dict_codes_kernel(int32_t* slot_then_code, int64_t n, const DictSlot* slots) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) slot_then_code[i] = (int32_t)slots[slot_then_code[i]].code;
}

__global__ void __launch_bounds__(256) narrow_offsets_kernel(const u64* src, u32* dst, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (u32)src[i];
}

// ---- association ingest -------------------------------------------------------------------------------------------

struct AssocStats {
    unsigned long long max_degree;   // largest number of targets of one row
    unsigned long long bad_offsets;  // rows whose offsets decrease
    int32_t min_target, max_target;  // over all edges (INT32_MAX / INT32_MIN when there are none)
};

__global__ void __launch_bounds__(256) assoc_stats_kernel(const int64_t* offsets, const int32_t* targets, int64_t n, int64_t nnz, AssocStats* out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long deg = 0, bad = 0;
    for (int64_t i = t0; i < n; i += stride) {
        const int64_t d = offsets[i + 1] - offsets[i];
        if (d < 0) ++bad;
        else if ((unsigned long long)d > deg) deg = (unsigned long long)d;
    }
    int32_t lo = INT32_MAX, hi = INT32_MIN;
    for (int64_t e = t0; e < nnz; e += stride) {
        const int32_t v = targets[e];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long od = __shfl_down_sync(FULL_MASK, deg, d), ob = __shfl_down_sync(FULL_MASK, bad, d);
        const int32_t a = __shfl_down_sync(FULL_MASK, lo, d), b = __shfl_down_sync(FULL_MASK, hi, d);
        deg = od > deg ? od : deg;
        bad += ob;
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        if (deg) atomicMax(&out->max_degree, deg);
        if (bad) atomicAdd(&out->bad_offsets, bad);
        atomicMin(&out->min_target, lo);
        atomicMax(&out->max_target, hi);
    }
}

// every row is None or One: the dense to-one form (Association.None -> -1, DS/Association.java:27-43)
__global__ void __launch_bounds__(256) csr_to_fk_kernel(const int64_t* offsets, const int32_t* targets, int64_t n, int32_t* fk) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t e0 = offsets[i];
        fk[i] = offsets[i + 1] > e0 ? targets[e0] : -1;
    }
}

}  // namespace colq
