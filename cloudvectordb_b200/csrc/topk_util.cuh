// Candidate keys and the warp-level bitonic network used by the fused top-k
// epilogue and by the merge kernels.
//
// A candidate is one 64-bit key: (orderable(score) << 32) | ~row.  Larger key
// == better candidate: higher score first, then LOWER row id (the oracle's tie
// rule).  Key 0 is "no candidate" (it would need row == 0xFFFFFFFF).
#pragma once
#include <stdint.h>

namespace cvdb {

__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    if (u == 0x80000000u) u = 0;  // -0.0 == +0.0: one code for both, or equal scores would order by sign
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
    return (static_cast<uint64_t>(float_to_ordered(score)) << 32) | static_cast<uint32_t>(~row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return ordered_to_float(static_cast<uint32_t>(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return ~static_cast<uint32_t>(k); }

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int mask) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v), mask);
    uint32_t hi = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), mask);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), src);
    uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), src);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Sort 32*E keys held by one warp into DESCENDING order.  Element p lives in
// register key[p / 32] of lane p % 32.  Fully unrolled: every register index
// is a compile-time constant.
template <int E>
__device__ __forceinline__ void warp_bitonic_sort_desc(uint64_t (&key)[E]) {
    const uint32_t lane = threadIdx.x & 31;
    constexpr int C = 32 * E;
#pragma unroll
    for (int size = 2; size <= C; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride < 32) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const uint64_t other = shfl_xor_u64(key[e], stride);
                    const uint32_t p = e * 32 + lane;
                    const bool desc_block = (size == C) || ((p & size) == 0);
                    const bool first = (lane & stride) == 0;
                    const bool take_max = (desc_block == first);
                    const uint64_t mx = key[e] > other ? key[e] : other;
                    const uint64_t mn = key[e] > other ? other : key[e];
                    key[e] = take_max ? mx : mn;
                }
            } else {
                constexpr int dummy = 0;
                (void)dummy;
                const int es = stride >> 5;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    if ((e & es) == 0) {
                        const int e2 = e | es;
                        const bool desc_block = (size == C) || (((e * 32) & size) == 0);
                        const uint64_t a = key[e], b = key[e2];
                        const uint64_t mx = a > b ? a : b;
                        const uint64_t mn = a > b ? b : a;
                        key[e] = desc_block ? mx : mn;
                        key[e2] = desc_block ? mn : mx;
                    }
                }
            }
        }
    }
}

// value of element `pos` (0-based, warp-uniform) after the sort, broadcast to all lanes
template <int E>
__device__ __forceinline__ uint64_t warp_sorted_at(const uint64_t (&key)[E], int pos) {
    uint64_t sel = key[0];
#pragma unroll
    for (int e = 1; e < E; ++e)
        if ((pos >> 5) == e) sel = key[e];
    return shfl_u64(sel, pos & 31);
}
// The select itself, on keys held in registers (slot e*32+lane in key[e], 0 = empty): returns T such that the
// survivors are exactly the non-empty keys >= T (the best k, or all of them when there are at most k), and sets
// kth to a key whose score word is the k-th best score (0 when fewer than k candidates exist).
template <int E>
__device__ __forceinline__ uint64_t warp_select_threshold(const uint64_t (&key)[E], int k, uint64_t& kth) {
    constexpr unsigned kFull = 0xffffffffu;
    uint32_t hi[E];
    int my_valid = 0;
    uint32_t mx = 0, mn = 0xFFFFFFFFu;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        hi[e] = static_cast<uint32_t>(key[e] >> 32);  // 0 <=> empty slot (the score word of a real candidate is never 0)
        my_valid += hi[e] != 0;
        mx = max(mx, hi[e]);
        mn = hi[e] != 0 ? min(mn, hi[e]) : mn;
    }
    const int n_valid = __reduce_add_sync(kFull, my_valid);
    mx = __reduce_max_sync(kFull, mx);
    mn = __reduce_min_sync(kFull, mn);
    kth = 0;
    if (n_valid <= k) {  // warp-uniform
        if (n_valid == k) kth = (static_cast<uint64_t>(mn) << 32) | 1u;
        return 1;  // keep every candidate
    }
    uint32_t prefix = mx;
    const uint32_t diff = mx ^ mn;
    if (diff != 0) {
        int bit = 31 - __clz(diff);
        prefix = mx & ~((2u << bit) - 1u);  // the bits every candidate shares
        for (; bit >= 0; --bit) {
            const uint32_t trial = prefix | (1u << bit);
            int c = 0;
#pragma unroll
            for (int e = 0; e < E; ++e) c += hi[e] >= trial;
            if (__reduce_add_sync(kFull, c) >= k) prefix = trial;
        }
    }
    // prefix == the k-th largest score word
    int c_gt = 0, c_eq = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        c_gt += hi[e] > prefix;
        c_eq += hi[e] == prefix;
    }
    c_gt = __reduce_add_sync(kFull, c_gt);
    c_eq = __reduce_add_sync(kFull, c_eq);
    const int need = k - c_gt;  // how many of the rows tied at the k-th score stay: the lowest ids
    uint32_t low = 0;
    if (need < c_eq) {
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t trial = low | (1u << bit);
            int c = 0;
#pragma unroll
            for (int e = 0; e < E; ++e) c += (hi[e] == prefix) && (static_cast<uint32_t>(key[e]) >= trial);
            if (__reduce_add_sync(kFull, c) >= need) low = trial;
        }
    }
    const uint64_t T = (static_cast<uint64_t>(prefix) << 32) | low;
    kth = T | 1u;
    return T;
}

// Move the survivors (non-empty keys >= T) to the front of `b`, in any order; returns how many there are.
template <int E>
__device__ __forceinline__ int warp_store_survivors(uint64_t* b, const uint64_t (&key)[E], uint64_t T) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    int base = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const bool keep = (key[e] >> 32) != 0 && key[e] >= T;
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (keep) b[base + __popc(m & lt_mask)] = key[e];
        base += __popc(m);
    }
    return base;
}

#endif  // __CUDACC__

}  // namespace cvdb
