// Fused distance GEMM + per-query top-k for sm_100a.
//
//   scores[q, r] = sum_k Q[q, k] * X[r, k]        (bf16 x bf16 -> fp32, tcgen05.mma)
//   per query keep the k best (score desc, row asc) -- the score matrix lives
//   only in TMEM, never in HBM.
//
// This file holds the kernel families (see DESIGN.md 4.1 for when each one runs):
//   gemm_topk_ss_kernel       single CTA, both operands streamed          (variant 1)
//   gemm_topk_ts2_kernel      CTA pair, queries resident in TMEM / smem   (variants 2 and 4)
//   gemm_topk_ss2_kernel      CTA pair, both operands streamed            (variant 3)
//   gemm_topk_grouped_kernel  single CTA, work items from a table (IVF list scan)
//
// Roles inside one CTA (256 threads, one CTA per SM, persistent):
//   warp 0      TMA producer: streams 64-wide K blocks of the query tile (A)
//               and of the database tile (B) into a STAGES-deep smem ring
//   warp 1      MMA issuer: one thread issues tcgen05.mma into one of two TMEM
//               accumulators (128 lanes x BLOCK_N fp32 columns each)
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: thread t owns query row t of the tile (= TMEM lane t),
//               reads its BLOCK_N scores with tcgen05.ld, and runs a threshold
//               filter against the query's current k-th best score.  Survivors
//               are appended to a per-query candidate buffer (32*E slots, L2
//               resident); when a buffer is nearly full the warp sorts it with
//               a register bitonic network and keeps the best k.
//
// Work item = (query tile, database slice).  Items are ordered slice-major so
// CTAs that run concurrently stream the SAME database rows (they hit in L2 and
// HBM sees each database byte about once per batch).  Every item ends by
// writing its sorted top-k to part[query][slice][k]; merge_partials_kernel
// (aux_kernels.cuh) does the final k-way merge.
#pragma once
#include <cuda_bf16.h>

#include "gemm_topk_params.h"
#include "ptx_sm100.cuh"
#include "topk_util.cuh"

// Buffers of at least 32 * CVDB_SELECT_MIN_E slots are compacted with a select instead of a sort (A/B: -D...=99).
#ifndef CVDB_SELECT_MIN_E
#define CVDB_SELECT_MIN_E 4
#endif

namespace cvdb {

// ---------------------------------------------------------------------------
// Soft wave alignment.  Work items are ordered slice-major so that CTAs running
// at the same time stream the same database rows and share them through L2,
// but with a static schedule CTAs drift apart over a launch (a 1 % speed
// difference is more than the L2 can bridge) and every one ends up fetching
// the database from HBM on its own.  Before starting the loads of its next
// item a producer therefore waits until every producer has finished issuing
// the loads of the current one.  It is only a performance hint: the wait gives
// up after a bounded time, so it cannot deadlock even if some CTAs are not
// resident.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void wave_arrive(uint32_t* cnt) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(cnt) : "memory");
}
__device__ __forceinline__ void wave_wait(const uint32_t* cnt, uint32_t expected) {
    const uint64_t t0 = globaltimer_ns();
    for (;;) {
        uint32_t v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
        if (v >= expected || globaltimer_ns() - t0 > 400000ull) break;
        __nanosleep(256);
    }
}
// producers (CTAs) that take part in wave `it`
__device__ __forceinline__ uint32_t wave_members(int it, int n_items, int n_workers, int ctas_per_worker) {
    const int left = n_items - it * n_workers;
    return static_cast<uint32_t>((left < n_workers ? left : n_workers) * ctas_per_worker);
}

// ---------------------------------------------------------------------------
// Per-thread (= per-query) selection state used by the epilogue warps.
// ---------------------------------------------------------------------------
template <int E>
struct LaneTopk {
    float thr;       // a score must be strictly greater than this to be a candidate
    int cnt;         // filled slots in buf
    uint64_t* buf;   // 32*E slots, slot p of every query of the warp is zero when p >= cnt
    uint32_t* gq;    // &gthr[query] (null for padding rows)
    int grp = -1;    // group id of the query when same-group rows are excluded (checked when the buffer is sorted)
};
template <>
struct LaneTopk<0> {
    float thr;
    uint64_t best;
    uint32_t* gq;
};

// Threshold sharing between database slices.  gthr[q] holds (as an ordered
// uint) the k-th best score some slice has already found for query q, i.e. at
// least k rows score >= it, so any row scoring strictly less can be dropped by
// every other slice too.  Equal scores must still pass (the tie rule is decided
// by row id at the merge), hence the "- 1" when adopting a foreign threshold.
// The next code below +0.0 (0x80000000) would be -0.0, and `s > -0.0f` is false for s == +0.0: step to a
// negative number instead (any value below the shared score is a valid filter bound; keys decide ties).
__device__ __forceinline__ float thr_from_shared(uint32_t g) {
    if (g == 0x80000000u) return -1.17549435e-38f;
    return g > 1u ? ordered_to_float(g - 1u) : -INFINITY;
}
// strict_own: rows still to come in this item have larger ids than everything collected so far (true for the
// flat kernels, which walk rows in id order), so a score equal to the own k-th cannot displace it.
template <bool strict_own = true>
__device__ __forceinline__ float publish_and_refresh(uint32_t* gq, uint64_t kth, float thr) {
    const uint32_t ord = static_cast<uint32_t>(kth >> 32);  // 0 when fewer than k candidates exist
    float t = ord ? (strict_own ? ordered_to_float(ord) : thr_from_shared(ord)) : -INFINITY;
    if (gq != nullptr) {
        const uint32_t old = atomicMax(gq, ord);
        if (old > ord) t = fmaxf(t, thr_from_shared(old));
    }
    return fmaxf(thr, t);
}

// POOLED THRESHOLDS across database slices.  gthr[q] only ever holds the best k-th score of a SINGLE slice, i.e. the
// k-th best of ~rows/slices rows: with S slices every item still collects about k candidates per query (S*k in all,
// where one running threshold over the whole database would collect ~k*ln(N/k)), and at k = 200 three of four 32x32
// chunks took the slow path for that reason.  But finished slices say more than their k-th score: if m DIFFERENT
// slices each hold ceil(k/m) rows scoring >= t, then k rows score >= t.  So every item also publishes, for
// m = 2, 4, 8, 16 (kPoolLevels), the score of rank ceil(k/m) of its result list into gpool[q] -- per level a list of the m largest
// such values seen so far (kept by an atomicMax chain: each step keeps the larger value in the slot and carries the
// smaller one down, so the slots always hold values of m distinct slices) -- and an item starts from the largest of
// {gthr[q], min over the slots of every complete level}.  With 16 finished slices the bound is the rank-k/16 score
// of a slice: about 16x fewer candidates per item.  Exact: only rows that provably cannot reach the global top-k
// are filtered, equal scores still pass (thr_from_shared).  Off with slice inheritance (lists of different slices
// then share rows) and for k = 1 (gthr is already exact).
__device__ __forceinline__ float pooled_threshold(const uint32_t* __restrict__ g) {
    uint32_t w[kPoolSlots];
#pragma unroll
    for (int i = 0; i < kPoolSlots / 4; ++i) {
        const uint4 v = __ldcg(reinterpret_cast<const uint4*>(g) + i);
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
    uint32_t best = 0;
#pragma unroll
    for (int m = 2; m <= (1 << kPoolLevels); m <<= 1) {
        uint32_t mn = 0xFFFFFFFFu;  // an empty slot (0) makes the level incomplete: min = 0
#pragma unroll
        for (int j = 0; j < m; ++j) mn = min(mn, w[m - 2 + j]);
        best = max(best, mn);
    }
    return best != 0 ? thr_from_shared(best) : -INFINITY;
}
// lv[i] = score word at rank ceil(k / (2 << i)) of a sorted list held in registers (0: the list is shorter)
template <int ES>
__device__ __forceinline__ void pool_stats(const uint64_t (&key)[ES], int k, uint32_t (&lv)[kPoolLevels]) {
#pragma unroll
    for (int i = 0; i < kPoolLevels; ++i) {
        const int m = 2 << i;
        const int pos = (k + m - 1) / m - 1;
        lv[i] = pos < 32 * ES ? static_cast<uint32_t>(warp_sorted_at<ES>(key, pos) >> 32) : 0u;
    }
}
// every lane: publish the order statistics `lv` of query `q` (q < 0: nothing to publish); the four levels' chains
// are independent, so their atomics overlap
__device__ __forceinline__ void pool_publish(uint32_t* __restrict__ gpool, int q, const uint32_t (&lv)[kPoolLevels]) {
    if (q < 0) return;
    uint32_t* g = gpool + static_cast<size_t>(q) * kPoolSlots;
    uint32_t v[kPoolLevels];
#pragma unroll
    for (int i = 0; i < kPoolLevels; ++i) v[i] = lv[i];
#pragma unroll
    for (int j = 0; j < (1 << kPoolLevels); ++j) {
#pragma unroll
        for (int i = 0; i < kPoolLevels; ++i) {
            const int m = 2 << i;
            if (j < m && v[i] != 0) {
                const uint32_t old = atomicMax(g + (m - 2) + j, v[i]);
                v[i] = min(old, v[i]);
            }
        }
    }
}

// Group exclusion (hard-negative mining: rows of the query's own group are not candidates).  Looking the group of
// a row up while scanning would put an L2 round trip into the per-candidate path, so rows are collected
// unchecked and the check runs here, on the whole buffer at once (the loads overlap), right before every sort;
// thresholds only ever come out of a sort, so an unchecked row never influences one.  grp < 0: nothing to do.
template <int E>
__device__ __forceinline__ void drop_same_group(uint64_t (&key)[E], int grp, const int32_t* __restrict__ group_db) {
    if (grp < 0 || group_db == nullptr) return;  // warp-uniform
    int g[E];
#pragma unroll
    for (int e = 0; e < E; ++e) g[e] = key[e] != 0 ? __ldg(group_db + key_row(key[e])) : -1;
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (g[e] == grp) key[e] = 0;
}

// Sort the buffer of lane `l` (warp-cooperative), keep its best k, zero the rest.
// Returns the k-th best key (0 when fewer than k candidates exist).
template <int E>
__device__ __forceinline__ uint64_t warp_compact(uint64_t* b, int k, uint64_t (&key)[E], int grp = -1,
                                                 const int32_t* __restrict__ group_db = nullptr) {
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < E; ++e) key[e] = __ldcg(reinterpret_cast<const unsigned long long*>(b + e * 32 + lane));
    drop_same_group<E>(key, grp, group_db);
    warp_bitonic_sort_desc<E>(key);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int pos = e * 32 + lane;
        if (pos >= k) key[e] = 0;
        b[pos] = key[e];
    }
    return warp_sorted_at<E>(key, k - 1);
}

// Compaction WITHOUT sorting (buffers of 128..512 slots, k = 29..248).  Between compactions a buffer only has
// to hold a superset of the query's best k, in any order -- it is sorted once, when the work item ends.  So a
// full buffer is reduced with a SELECT: find the k-th largest key by bisection on the score word (one
// per-lane count over the E registers + one warp reduction per bit, starting below the bits all candidates
// share), resolve a tie at the k-th score on the row word (rare), then move the k survivors to the front with
// ballot prefix sums.  ~(2E + 2) instructions per bit against ~12E per stage of the E*log^2 bitonic network
// (512 slots: ~1k instructions instead of ~9k), and a dependent chain of ~20 reductions instead of 45 shuffle
// stages -- a sort used to hold its epilogue warp for more than two MMA tiles, i.e. it stalled the tensor pipe.
// Returns a key whose score word is the k-th best score (0 when fewer than k candidates exist); `kept` = number of
// survivors now at the front of the buffer (slots behind them are zero).
template <int E>
__device__ __forceinline__ uint64_t warp_select_compact(uint64_t* b, int k, int& kept, int grp = -1,
                                                        const int32_t* __restrict__ group_db = nullptr) {
    const uint32_t lane = threadIdx.x & 31;
    uint64_t key[E];
    uint32_t was = 0;  // bit e: slot e*32+lane held a candidate when it was loaded
#pragma unroll
    for (int e = 0; e < E; ++e) {
        key[e] = __ldcg(reinterpret_cast<const unsigned long long*>(b + e * 32 + lane));
        was |= static_cast<uint32_t>(key[e] != 0) << e;
    }
    drop_same_group<E>(key, grp, group_db);
    uint64_t kth;
    const uint64_t T = warp_select_threshold<E>(key, k, kth);
    kept = warp_store_survivors<E>(b, key, T);
#pragma unroll
    for (int e = 0; e < E; ++e) {  // the flat kernels rely on "slots behind the candidates are zero"
        const int pos = e * 32 + static_cast<int>(lane);
        if (pos >= kept && ((was >> e) & 1u)) b[pos] = 0;
    }
    return kth;
}

// Large buffers (E > 16, i.e. k > 248): the same bitonic network, but run as loops over the buffer where it
// lives (L2-resident global memory) instead of unrolled over registers -- a 1024..4096-key register network would
// not fit the register file (and takes minutes to compile).  Slower per sort, but these buffers are sized >= 2k so
// sorts are rare.  Returns the k-th best key; leaves the best k sorted at the front and zeros behind them.
template <int E>
__device__ __noinline__ uint64_t warp_compact_mem(uint64_t* b, int k, int grp = -1,
                                                  const int32_t* __restrict__ group_db = nullptr) {
    constexpr int C = 32 * E;
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (grp >= 0 && group_db != nullptr) {  // see drop_same_group
        for (int t = lane; t < C; t += 32) {
            const uint64_t a = __ldcg(reinterpret_cast<const unsigned long long*>(b + t));
            if (a != 0 && __ldg(group_db + key_row(a)) == grp) b[t] = 0;
        }
        __syncwarp();
    }
    for (int size = 2; size <= C; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int sh = __ffs(stride) - 1;
            for (int t = lane; t < C / 2; t += 32) {
                const int i = ((t >> sh) << (sh + 1)) | (t & (stride - 1));
                const int j = i + stride;
                const bool desc = (size == C) || ((i & size) == 0);
                const uint64_t a = __ldcg(reinterpret_cast<const unsigned long long*>(b + i));
                const uint64_t c = __ldcg(reinterpret_cast<const unsigned long long*>(b + j));
                if (desc ? (a < c) : (a > c)) {
                    b[i] = c;
                    b[j] = a;
                }
            }
            __syncwarp();
        }
    }
    for (int p = k + lane; p < C; p += 32) b[p] = 0;
    __syncwarp();
    return __ldcg(reinterpret_cast<const unsigned long long*>(b + k - 1));
}

// `room`: compact once a buffer holds more than this many candidates (set by the host: 32*E - 8 for E = 1,
// 32*E - 32 otherwise, see scan_chunk).  Between
// compactions the threshold is stale (it admits rows that the next sort throws away again); compacting earlier
// was measured and is slower, because every sort stalls the warp on L2 round trips (cvdb_api.cu).
template <int E, bool strict_own = true>
__device__ __forceinline__ void make_room(LaneTopk<E>& st, int k, int room, const int32_t* __restrict__ group_db = nullptr) {
    unsigned mask = __ballot_sync(0xffffffffu, st.cnt > room);
    if (mask == 0) return;
    __syncwarp();
    const uint32_t lane = threadIdx.x & 31;
    while (mask) {
        const int l = __ffs(mask) - 1;
        mask &= mask - 1;
        uint64_t* b = reinterpret_cast<uint64_t*>(shfl_u64(reinterpret_cast<uint64_t>(st.buf), l));
        const int grp_l = __shfl_sync(0xffffffffu, st.grp, l);
        uint64_t kth;
        int kept = k;
        if constexpr (E >= CVDB_SELECT_MIN_E && E <= 16) {
            kth = warp_select_compact<E>(b, k, kept, grp_l, group_db);
        } else if constexpr (E <= 16) {
            uint64_t key[E];
            kth = warp_compact<E>(b, k, key, grp_l, group_db);
        } else {
            kth = warp_compact_mem<E>(b, k, grp_l, group_db);
        }
        if (static_cast<int>(lane) == l) {
            st.thr = publish_and_refresh<strict_own>(st.gq, kth, st.thr);
            st.cnt = st.cnt < kept ? st.cnt : kept;
        }
    }
    __syncwarp();
}

// Process 32 scores (database rows row0 .. row0+31) of this thread's query.
// Fast path: four independent max-of-8 chains, one vote.  Slow path: only the
// groups of 8 in which some lane has a candidate are walked element by element.
template <int E>
__device__ __forceinline__ void scan_chunk(LaneTopk<E>& st, const uint32_t (&v)[32], uint32_t row0, uint32_t row_end,
                                           uint32_t self, int grp, const int32_t* __restrict__ group_db, int k, int room,
                                           float* m8_out = nullptr) {
    float m8[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float m = fmaxf(fmaxf(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])), __uint_as_float(v[8 * g + 2]));
        m = fmaxf(fmaxf(m, __uint_as_float(v[8 * g + 3])), __uint_as_float(v[8 * g + 4]));
        m = fmaxf(fmaxf(m, __uint_as_float(v[8 * g + 5])), __uint_as_float(v[8 * g + 6]));
        m8[g] = fmaxf(m, __uint_as_float(v[8 * g + 7]));
    }
    const float m = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
    if (m8_out != nullptr) {  // the column direction of the symmetric join tests the same maxima
        m8_out[0] = m8[0]; m8_out[1] = m8[1]; m8_out[2] = m8[2]; m8_out[3] = m8[3];
    }
    if (!__any_sync(0xffffffffu, m > st.thr)) return;  // common case once the threshold has settled
    if constexpr (E >= 2) {
        // Buffers of 64+ slots: make room for a whole chunk (32 candidates) once, then let only the lanes that
        // hold a candidate walk their groups -- typically a single lane with a single row, so the divergent
        // walk costs one pass instead of a warp-wide pass per group of eight.
        make_room<E>(st, k, room, group_db);
        if (m > st.thr) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (!(m8[g] > st.thr)) continue;
#pragma unroll
                for (int j = 8 * g; j < 8 * g + 8; ++j) {
                    const float s = __uint_as_float(v[j]);
                    if (s > st.thr) {
                        const uint32_t row = row0 + j;
                        if (row < row_end && row != self) st.buf[st.cnt++] = make_key(s, row);  // group: see drop_same_group
                    }
                }
            }
        }
    } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (!__any_sync(0xffffffffu, m8[g] > st.thr)) continue;
            if constexpr (E > 0) make_room<E>(st, k, room, group_db);
#pragma unroll
            for (int j = 8 * g; j < 8 * g + 8; ++j) {
                const float s = __uint_as_float(v[j]);
                if (s > st.thr) {
                    const uint32_t row = row0 + j;
                    bool ok = row < row_end && row != self;
                    // k = 1 keeps its best row in a register, so it has to check the group right here
                    if (E == 0 && ok && grp >= 0 && group_db != nullptr) ok = __ldg(group_db + row) != grp;
                    if (ok) {
                        if constexpr (E > 0) {
                            st.buf[st.cnt++] = make_key(s, row);
                        } else {
                            st.best = make_key(s, row);
                            st.thr = s;
                        }
                    }
                }
            }
        }
    }
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Start of a work item: reset the selection state of this thread's query.
//
// INHERITANCE (optional, p.done != null; off by default: see cvdb_api.cu for the A/B that decided it).
// A slice that starts from an empty buffer only ever learns the k-th best score of ITS OWN rows,
// and the shared threshold gthr[q] is the best such single-slice value: with S slices every slice still collects
// about k candidates per query, S*k in all, where a scan of the whole database with one running threshold would
// collect ~k*ln(N/k).  At k = 200 on the headline shape that made every second 32x32 chunk take the slow
// (candidate) path.  So an item starts from the sorted result list of the latest FINISHED earlier slice of the
// same query tile (part[q][s'], found through the done[] counters): the buffer begins with those k keys, the
// threshold with their k-th score, and what the item writes at its end is the top-k of everything the chain has
// seen.  The lists of different slices then overlap; the final merge drops the copies (a key present in several
// lists reaches the head of all of them in the same round).  Results do not depend on which slice, if any, was
// inherited.
template <int E>
__device__ __forceinline__ void item_begin(LaneTopk<E>& st, const GemmTopkParams& p, int q_row, bool q_valid,
                                           uint64_t* warp_buf, int qt = 0, int slice = 0) {
    constexpr int C = 32 * (E > 0 ? E : 1);
    const uint32_t lane = threadIdx.x & 31;
    st.gq = (q_valid && p.gthr != nullptr) ? p.gthr + q_row : nullptr;
    // padding lanes (and, in a quartered small-batch tile, lanes that see another warp's query) never collect
    st.thr = !q_valid ? INFINITY : (st.gq ? thr_from_shared(__ldcg(st.gq)) : -INFINITY);
    if constexpr (E > 0) {
        if (p.gpool != nullptr && q_valid)
            st.thr = fmaxf(st.thr, pooled_threshold(p.gpool + static_cast<size_t>(q_row) * kPoolSlots));
        st.cnt = 0;
        int src = -1;
        if (E <= 16 && p.done != nullptr) {  // warp-uniform: every lane reads the same counters
            const uint32_t* d = p.done + static_cast<size_t>(qt) * p.n_slices;
            for (int s = slice - 1; s >= 0 && s >= slice - 4; --s)
                if (ld_acquire_u32(d + s) >= static_cast<uint32_t>(p.done_full)) { src = s; break; }
        }
        if (src < 0) {
            for (int i = lane; i < 32 * C; i += 32) warp_buf[i] = 0;  // the warp's 32 buffers are contiguous
            __syncwarp();
        } else {
            const int k = p.k;
            constexpr int EE = E > 0 ? E : 1;
#pragma unroll 2
            for (int l = 0; l < 32; ++l) {
                const int ql = __shfl_sync(0xffffffffu, q_row, l);
                const bool vl = __shfl_sync(0xffffffffu, static_cast<int>(q_valid), l) != 0;
                const uint64_t* from = p.part + (static_cast<size_t>(vl ? ql : 0) * p.n_slices + src) * k;
                uint64_t* b = warp_buf + static_cast<size_t>(l) * C;
                uint64_t key[EE];
#pragma unroll
                for (int e = 0; e < EE; ++e) {  // all loads of the list in flight at once: one round trip per list
                    const int pos = e * 32 + static_cast<int>(lane);
                    key[e] = (vl && pos < k) ? __ldcg(reinterpret_cast<const unsigned long long*>(from + pos)) : 0ull;
                }
                int n = 0;
#pragma unroll
                for (int e = 0; e < EE; ++e) {
                    b[e * 32 + lane] = key[e];
                    n += __popc(__ballot_sync(0xffffffffu, key[e] != 0));
                }
                if (static_cast<int>(lane) == l) st.cnt = n;
            }
            __syncwarp();
            if (st.cnt >= k) {  // a full list: its k-th score bounds what can still enter (equal scores stay candidates)
                const uint64_t kth = __ldcg(reinterpret_cast<const unsigned long long*>(st.buf + k - 1));
                st.thr = fmaxf(st.thr, thr_from_shared(static_cast<uint32_t>(kth >> 32)));
            }
        }
    } else {
        st.best = 0;
    }
}

// A buffer whose candidates all sit in its first 32*ES slots: sort only those (warp-cooperative) and write the k
// output entries.  Returns the k-th best key (0 when fewer than k candidates exist).
template <int ES>
__device__ __forceinline__ uint64_t flush_prefix(const uint64_t* b, uint64_t* out, int k, int grp,
                                                 const int32_t* __restrict__ group_db, uint32_t (&lv)[kPoolLevels]) {
    const int lane = threadIdx.x & 31;
    uint64_t key[ES];
#pragma unroll
    for (int e = 0; e < ES; ++e) key[e] = __ldcg(reinterpret_cast<const unsigned long long*>(b + e * 32 + lane));
    drop_same_group<ES>(key, grp, group_db);
    warp_bitonic_sort_desc<ES>(key);
#pragma unroll
    for (int e = 0; e < ES; ++e) {
        const int pos = e * 32 + lane;
        if (pos < k) out[pos] = key[e];
    }
    for (int pos = 32 * ES + lane; pos < k; pos += 32) out[pos] = 0;
    pool_stats<ES>(key, k, lv);
    return k <= 32 * ES ? warp_sorted_at<ES>(key, k - 1) : 0;
}

// End of a work item: sorted top-k of every query of the warp -> part[q][slice][0..k).
template <int E>
__device__ __forceinline__ void item_flush(LaneTopk<E>& st, const GemmTopkParams& p, int q_base, int q_row, bool q_valid,
                                           int slice, int q_lanes = 32, int qt = 0) {
    const uint32_t lane = threadIdx.x & 31;
    if constexpr (E > 0) {
        __syncwarp();
        int pool_q = -1;                       // this lane's query and its order statistics for gpool
        uint32_t pool_lv[kPoolLevels] = {};
        for (int l = 0; l < q_lanes; ++l) {
            const int qr = q_base + l;
            if (qr >= p.nq) break;
            uint32_t lv[kPoolLevels] = {};
            uint64_t* b = reinterpret_cast<uint64_t*>(shfl_u64(reinterpret_cast<uint64_t>(st.buf), l));
            uint64_t* out = p.part + (static_cast<size_t>(qr) * p.n_slices + slice) * p.k;
            const int grp_l = __shfl_sync(0xffffffffu, st.grp, l);
            uint64_t kth;
            if constexpr (E <= 16) {
                // Later slices start from a good shared threshold and often collect only a handful of rows: sort
                // just the filled prefix then (32 slots cost a fifth of the 128-slot network of k = 50).
                const int cnt_l = (p.dbg & 32) ? 32 * E : __shfl_sync(0xffffffffu, st.cnt, l);
                if (E > 1 && cnt_l <= 32) {
                    kth = flush_prefix<1>(b, out, p.k, grp_l, p.group_db, lv);
                } else if (E > 2 && cnt_l <= 64) {
                    kth = flush_prefix<2>(b, out, p.k, grp_l, p.group_db, lv);
                } else if (E > 4 && cnt_l <= 128) {
                    kth = flush_prefix<4>(b, out, p.k, grp_l, p.group_db, lv);
                } else if (E > 8 && cnt_l <= 256) {
                    kth = flush_prefix<8>(b, out, p.k, grp_l, p.group_db, lv);
                } else {
                    uint64_t key[E];
                    kth = warp_compact<E>(b, p.k, key, grp_l, p.group_db);
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int pos = e * 32 + lane;
                        if (pos < p.k) out[pos] = key[e];
                    }
                    pool_stats<E>(key, p.k, lv);
                }
            } else {
                kth = warp_compact_mem<E>(b, p.k, grp_l, p.group_db);
                for (int pos = lane; pos < p.k; pos += 32)
                    out[pos] = __ldcg(reinterpret_cast<const unsigned long long*>(b + pos));
            }
            if (static_cast<int>(lane) == l && kth != 0 && st.gq != nullptr) atomicMax(st.gq, static_cast<uint32_t>(kth >> 32));
            if (static_cast<int>(lane) == l && st.gq != nullptr) {
                pool_q = qr;
#pragma unroll
                for (int i = 0; i < kPoolLevels; ++i) pool_lv[i] = lv[i];
            }
        }
        if (p.gpool != nullptr) pool_publish(p.gpool, pool_q, pool_lv);   // all lanes at once: the chains overlap
        if (p.done != nullptr) {  // this warp's lists of the item are in part[]: later slices may start from them
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(p.done + static_cast<size_t>(qt) * p.n_slices + slice, 1u);
        }
        __syncwarp();
    } else {
        if (q_valid) p.part[static_cast<size_t>(q_row) * p.n_slices + slice] = st.best;
    }
}

template <int BLOCK_N, int STAGES, int E>
__global__ void __launch_bounds__(256, 1)
gemm_topk_ss_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                    const GemmTopkParams p) {
    constexpr int BLOCK_M = 128;
    constexpr int BLOCK_K = 64;
    constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
    constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
    static_assert(2 * BLOCK_N <= 512, "two accumulators must fit TMEM");
    constexpr int C = 32 * (E > 0 ? E : 1);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_x);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, TMEM_COLS);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_items = p.q_tiles * p.n_slices;
    const int ksteps = p.nkb * p.n_combo;

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer (warp-uniform loop)
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
            const int slice = w / p.q_tiles, qt = w - slice * p.q_tiles;
            const int t0 = p.tile0 + slice * p.tiles_per_slice;
            const int t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
            if (p.wave_cnt != nullptr && it > 0) {
                if (elect_one_sync()) wave_wait(p.wave_cnt + it - 1, wave_members(it - 1, n_items, gridDim.x, 1));
                __syncwarp();
            }
            for (int t = t0; t < t1; ++t) {
                for (int c = 0; c < p.n_combo; ++c) {
                    const int a_col = static_cast<int>((p.a_planes >> (4 * c)) & 0xF) * p.plane_cols;
                    const int b_col = static_cast<int>((p.b_planes >> (4 * c)) & 0xF) * p.plane_cols;
                    for (int kb = 0; kb < p.nkb; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (elect_one_sync()) {
                            mbar_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
                            if (p.a_quarter > 0) {
                                // a small batch: spread its queries over the four epilogue warps (tmap_q has 32-row
                                // boxes then; the rows a box holds beyond its quarter stay idle lanes)
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    tma_load_2d(&tmap_q, &full_bar[stage], smem_a + stage * A_BYTES + j * (A_BYTES / 4),
                                                a_col + kb * BLOCK_K, j * p.a_quarter, kEvictLast);
                            } else {
                                tma_load_2d(&tmap_q, &full_bar[stage], smem_a + stage * A_BYTES, a_col + kb * BLOCK_K,
                                            qt * BLOCK_M, kEvictLast);
                            }
                            tma_load_2d(&tmap_x, &full_bar[stage], smem_b + stage * B_BYTES, b_col + kb * BLOCK_K,
                                        t * BLOCK_N, kEvictNormal);
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
            if (p.wave_cnt != nullptr && elect_one_sync()) wave_arrive(p.wave_cnt + it);
            __syncwarp();
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer (warp-uniform loop)
        constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const int slice = w / p.q_tiles;
            const int t0 = p.tile0 + slice * p.tiles_per_slice;
            const int t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int s = 0; s < ksteps; ++s) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * A_BYTES));
                        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * B_BYTES));
                        const int kb = s % p.nkb;
                        const int nk = min(4, p.k16 - 4 * kb);  // the last K block of a plane may be partly padding
                        if (nk == 4) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)  // +32 bytes along K inside the 128-byte swizzled row
                                umma_ss<1>(tmem_d, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (s | kk) != 0);
                        } else {
                            for (int kk = 0; kk < nk; ++kk)
                                umma_ss<1>(tmem_d, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (s | kk) != 0);
                        }
                        umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
                        if (s == ksteps - 1) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ----------------------------------------------------------- epilogue
        const int ewarp = warp - 4;  // == warp % 4: TMEM lanes [32*ewarp, 32*ewarp+32)
        const uint32_t lane_base = static_cast<uint32_t>(ewarp * 32) << 16;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int dbg = p.dbg;
        LaneTopk<E> st;
        if constexpr (E > 0)
            st.buf = p.cand + (static_cast<size_t>(blockIdx.x) * BLOCK_M + ewarp * 32 + lane) * C;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const int slice = w / p.q_tiles, qt = w - slice * p.q_tiles;
            const int t0 = p.tile0 + slice * p.tiles_per_slice;
            const int t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
            const int q_lanes = p.a_quarter > 0 ? p.a_quarter : 32;  // query rows owned by this warp
            const int q_base = qt * BLOCK_M + ewarp * q_lanes;
            const int q_row = q_base + static_cast<int>(lane);
            const bool q_valid = static_cast<int>(lane) < q_lanes && q_row < p.nq;
            const uint32_t self = (p.self_ids && q_valid) ? static_cast<uint32_t>(__ldg(p.self_ids + q_row)) : 0xFFFFFFFFu;
            const int grp = (p.group_q && q_valid) ? __ldg(p.group_q + q_row) : -1;
            item_begin<E>(st, p, q_row, q_valid, p.cand + (static_cast<size_t>(blockIdx.x) * BLOCK_M + ewarp * 32) * C, qt, slice);
            if constexpr (E > 0) st.grp = p.group_db != nullptr ? grp : -1;
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t row0 = static_cast<uint32_t>(t) * BLOCK_N;
                const uint32_t taddr = tmem_base + lane_base + acc * BLOCK_N;
#pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 32) {
                    uint32_t v[32];
                    if (!(dbg & 2)) {
                        tmem_ld32(taddr + c, v);
                        tmem_ld_wait();
                    }
                    if (c + 32 == BLOCK_N) {
                        // all of this thread's scores are in registers: hand the accumulator back
                        tc_fence_before();
                        mbar_arrive(&tmem_empty[acc]);
                    }
                    if (!(dbg & 3))
                        scan_chunk<E>(st, v, row0 + c, static_cast<uint32_t>(p.n_rows), self, grp, p.group_db, p.k, p.room);
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            item_flush<E>(st, p, q_base, q_row, q_valid, slice, q_lanes, qt);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem_base, TMEM_COLS);
}

template <int BLOCK_N, int STAGES>
constexpr size_t gemm_topk_ss_smem_bytes() {
    return 1024 /*align slack*/ + size_t(STAGES) * (128 * 64 * 2 + BLOCK_N * 64 * 2) + (2 * STAGES + 4) * 8 + 16;
}

// ===========================================================================
// Grouped variant for inverted-list (IVF) search.  The database rows are stored
// list-major; a work item is (one inverted list, up to 128 of the (query, probe)
// pairs that probe it).  The pairs' query rows have been gathered next to each
// other (qg), so the A tile is a plain TMA box again; B tiles walk the rows of
// the list.  Same pipeline and epilogue as the single-CTA streaming kernel; the
// differences are the item table, per-lane destinations (pair -> part[query][probe])
// and candidate ids translated to the caller's row ids (row_ids) when a
// candidate is appended.  Thresholds are shared per QUERY across its probes.
// ===========================================================================
template <int E>
__device__ __forceinline__ void scan_chunk_grouped(LaneTopk<E>& st, const uint32_t (&v)[32], uint32_t row0,
                                                   uint32_t row_end, uint32_t lane_id, int k, int room) {
    // lane_id: the caller's row id of the chunk's column `lane` (the ids of a chunk are the same for every query
    // row, so the warp loads them once, ahead of time, instead of one dependent lookup per candidate)
    float m8[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float m = fmaxf(fmaxf(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])), __uint_as_float(v[8 * g + 2]));
        m = fmaxf(fmaxf(m, __uint_as_float(v[8 * g + 3])), __uint_as_float(v[8 * g + 4]));
        m = fmaxf(fmaxf(m, __uint_as_float(v[8 * g + 5])), __uint_as_float(v[8 * g + 6]));
        m8[g] = fmaxf(m, __uint_as_float(v[8 * g + 7]));
    }
    const float m = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
    if (!__any_sync(0xffffffffu, m > st.thr)) return;
    uint32_t id[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) id[j] = __shfl_sync(0xffffffffu, lane_id, j);
    if constexpr (E >= 2) {
        // as in scan_chunk: room for a whole chunk once, then only the lanes with a candidate walk their groups
        make_room<E, false>(st, k, room);
        if (m > st.thr) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (!(m8[g] > st.thr)) continue;
#pragma unroll
                for (int j = 8 * g; j < 8 * g + 8; ++j) {
                    const float s = __uint_as_float(v[j]);
                    const uint32_t row = row0 + j;
                    if (s > st.thr && row < row_end) st.buf[st.cnt++] = make_key(s, id[j]);
                }
            }
        }
    } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (!__any_sync(0xffffffffu, m8[g] > st.thr)) continue;
            // rows of a list are stored in no particular id order: equal scores must stay candidates (keys decide)
            if constexpr (E > 0) make_room<E, false>(st, k, room);
#pragma unroll
            for (int j = 8 * g; j < 8 * g + 8; ++j) {
                const float s = __uint_as_float(v[j]);
                const uint32_t row = row0 + j;
                if (s > st.thr && row < row_end) {
                    if constexpr (E > 0) {
                        st.buf[st.cnt++] = make_key(s, id[j]);
                    } else {
                        // top-1: ids are not visited in increasing order here, so ties go through the key
                        const uint64_t key = make_key(s, id[j]);
                        if (key > st.best) st.best = key;
                        st.thr = thr_from_shared(static_cast<uint32_t>(st.best >> 32));  // equal scores still pass
                    }
                }
            }
        }
    }
}

template <int BLOCK_N, int STAGES, int E>
__global__ void __launch_bounds__(256, 1)
gemm_topk_grouped_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                         const GroupedParams p) {
    constexpr int BLOCK_M = 128;
    constexpr int BLOCK_K = 64;
    constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
    constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 256) ? 256 : 512;
    static_assert(2 * BLOCK_N <= 512 && BLOCK_N % 32 == 0, "two accumulators must fit TMEM");
    constexpr int C = 32 * (E > 0 ? E : 1);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_x);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, TMEM_COLS);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_items = __ldg(p.n_items_ptr);

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const GroupItem it = p.items[w];
            const int n_tiles = (it.x_rows + BLOCK_N - 1) / BLOCK_N;
            // An item usually has far fewer than 128 query rows.  Packed into consecutive tile rows they would all
            // sit in TMEM lanes 0.. and be scanned by a single epilogue warp, so the A tile is loaded as four
            // 32-row boxes, box j starting at the item's j-th quarter: every epilogue warp gets a quarter of the
            // pairs (the other rows of a box are somebody else's pairs; their lanes stay idle).
            const int q4 = (it.a_rows + 3) >> 2;
            for (int t = 0; t < n_tiles; ++t) {
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one_sync()) {
                        mbar_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            tma_load_2d(&tmap_q, &full_bar[stage], smem_a + stage * A_BYTES + j * (A_BYTES / 4), kb * BLOCK_K,
                                        it.a_row0 + j * q4, kEvictNormal);
                        tma_load_2d(&tmap_x, &full_bar[stage], smem_b + stage * B_BYTES, kb * BLOCK_K,
                                    it.x_row0 + t * BLOCK_N, kEvictNormal);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const GroupItem it = p.items[w];
            const int n_tiles = (it.x_rows + BLOCK_N - 1) / BLOCK_N;
            for (int t = 0; t < n_tiles; ++t) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * A_BYTES));
                        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * B_BYTES));
                        const int nk = min(4, p.k16 - 4 * kb);
                        for (int kk = 0; kk < nk; ++kk)
                            umma_ss<1>(tmem_d, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (kb | kk) != 0);
                        umma_commit(&empty_bar[stage]);
                        if (kb == p.nkb - 1) umma_commit(&tmem_full[acc]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ----------------------------------------------------------- epilogue
        const int ewarp = warp - 4;
        const uint32_t lane_base = static_cast<uint32_t>(ewarp * 32) << 16;
        int acc = 0;
        uint32_t acc_phase = 0;
        LaneTopk<E> st;
        uint64_t* warp_buf = p.cand + (static_cast<size_t>(blockIdx.x) * BLOCK_M + ewarp * 32) * C;
        if constexpr (E > 0) st.buf = warp_buf + static_cast<size_t>(lane) * C;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const GroupItem it = p.items[w];
            const int n_tiles = (it.x_rows + BLOCK_N - 1) / BLOCK_N;
            const int q4 = (it.a_rows + 3) >> 2;  // query rows per epilogue warp (see the producer)
            const int a_local = ewarp * q4 + static_cast<int>(lane);
            const bool valid = static_cast<int>(lane) < q4 && a_local < it.a_rows;
            const int a_row = it.a_row0 + a_local;
            const int query = valid ? __ldg(p.pair_query + a_row) : 0;
            const int dst = valid ? __ldg(p.pair_dst + a_row) : -1;
            // ---- reset the selection state; the threshold starts from what other probes of the query found
            st.gq = valid ? p.gthr + query : nullptr;
            st.thr = valid ? thr_from_shared(__ldcg(st.gq)) : INFINITY;  // invalid lanes never collect
            if constexpr (E > 0) {
                st.cnt = 0;
                for (int i = lane; i < 32 * C; i += 32) warp_buf[i] = 0;
                __syncwarp();
            } else {
                st.best = 0;
            }
            const uint32_t row_end = static_cast<uint32_t>(it.x_row0 + it.x_rows);
            for (int t = 0; t < n_tiles; ++t) {
                const uint32_t row0 = static_cast<uint32_t>(it.x_row0 + t * BLOCK_N);
                // row ids of the tile's columns, one chunk of 32 per register, fetched while the MMA still runs
                uint32_t tile_id[BLOCK_N / 32];
#pragma unroll
                for (int q = 0; q < BLOCK_N / 32; ++q) {
                    const uint32_t r = row0 + q * 32 + lane;
                    tile_id[q] = (p.row_ids != nullptr && r < row_end) ? static_cast<uint32_t>(__ldg(p.row_ids + r)) : r;
                }
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_base + acc * BLOCK_N;
#pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c, v);
                    tmem_ld_wait();
                    if (c + 32 == BLOCK_N) {
                        tc_fence_before();
                        mbar_arrive(&tmem_empty[acc]);
                    }
                    uint32_t lane_id = tile_id[0];
#pragma unroll
                    for (int q = 1; q < BLOCK_N / 32; ++q)
                        if (c == q * 32) lane_id = tile_id[q];
                    scan_chunk_grouped<E>(st, v, row0 + c, row_end, lane_id, p.k, p.room);
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            // ---- flush: part[dst][0..k)
            if constexpr (E > 0) {
                __syncwarp();
                for (int l = 0; l < 32; ++l) {
                    const int dst_l = __shfl_sync(0xffffffffu, dst, l);
                    if (dst_l < 0) continue;  // warp-uniform
                    uint64_t* b = warp_buf + static_cast<size_t>(l) * C;
                    uint64_t* out = p.part + static_cast<size_t>(dst_l) * p.k;
                    uint64_t kth;
                    uint32_t lv_unused[kPoolLevels];  // (order statistics for the flat kernels' pooled thresholds)
                    if constexpr (E <= 16) {
                        // a list is a few hundred rows: most buffers hold a few dozen candidates
                        const int cnt_l = __shfl_sync(0xffffffffu, st.cnt, l);
                        if (E > 1 && cnt_l <= 32) {
                            kth = flush_prefix<1>(b, out, p.k, -1, nullptr, lv_unused);
                        } else if (E > 2 && cnt_l <= 64) {
                            kth = flush_prefix<2>(b, out, p.k, -1, nullptr, lv_unused);
                        } else if (E > 4 && cnt_l <= 128) {
                            kth = flush_prefix<4>(b, out, p.k, -1, nullptr, lv_unused);
                        } else if (E > 8 && cnt_l <= 256) {
                            kth = flush_prefix<8>(b, out, p.k, -1, nullptr, lv_unused);
                        } else {
                            uint64_t key[E];
                            kth = warp_compact<E>(b, p.k, key);
#pragma unroll
                            for (int e = 0; e < E; ++e) {
                                const int pos = e * 32 + lane;
                                if (pos < p.k) out[pos] = key[e];
                            }
                        }
                    } else {
                        kth = warp_compact_mem<E>(b, p.k);
                        for (int pos = lane; pos < p.k; pos += 32)
                            out[pos] = __ldcg(reinterpret_cast<const unsigned long long*>(b + pos));
                    }
                    if (static_cast<int>(lane) == l && kth != 0) atomicMax(st.gq, static_cast<uint32_t>(kth >> 32));
                }
                __syncwarp();
            } else {
                if (valid) {
                    p.part[dst] = st.best;
                    if (st.best != 0) atomicMax(st.gq, static_cast<uint32_t>(st.best >> 32));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem_base, TMEM_COLS);
}

// ===========================================================================
// COLUMN DIRECTION (symmetric self-join).  When the queries are rows of the database itself, the score of
// (query i, row j) is also the score of (query j, row i): a tile computed once can feed the top-k of its query
// rows (the usual row-wise filter) AND the top-k of its database rows, with the query as the candidate.  The
// column side keeps its state in global memory, one threshold / counter / key buffer per database row: thresholds
// (GemmTopkParams::col_thr) are read once per tile, one tile ahead (lane l holds the columns l, l+32, ...); a chunk
// of 32 columns is skipped when no score of the warp beats the smallest of its 32 thresholds, a group of 8 when
// none beats the smallest of its 8; a candidate is appended to a log that col_scatter_kernel distributes to the
// rows' buffers after the launch.  Buffers are compacted between launches (col_compact_kernel): a row that ran over
// its buffer is flagged and recomputed exactly by the caller.
// ===========================================================================
// anchor_id: the query's id in the id space of the column lists (a global id when the database is one shard of
// many); self_row: the database row that IS the query (none: 0xFFFFFFFF); grp: the query's group (< 0: none),
// recorded with the candidate.
// A candidate is not written into its row's buffer here (claiming a slot needs the atomic's return value: one
// global round trip per candidate, ~2.5 us per slow chunk in the first version) but appended to a LOG: every
// thread owns segments of 8 records in a global array, reserved 8 at a time with one atomicAdd, and just stores
// (row, key) at its cursor.  col_scatter_kernel distributes the log to the rows after the launch, with all the
// parallelism of a plain kernel to hide the atomics.
// (Tried and dropped: a WARP-cooperative log -- one ballot per column, the lanes holding a candidate take consecutive
// records of a 256-record segment that lane 0 reserves one segment ahead, so no atomic is ever waited for.  Same-box
// A/B on the candidate-rich early chunks of a join: 404 ms against 328 ms for this per-thread version.  The epilogue
// is issue-bound there, and eight ballot/popc/branch sequences per group of columns cost more than the few
// divergent pushes they replace.)
struct ColLogCursor {
    uint32_t cur = 0, left = 0;  // next record of this thread's segment, records left in it
};
constexpr int kColLogSeg = 8;
__device__ __forceinline__ void col_log_push(ColLogCursor& lc, const GemmTopkParams& p, uint32_t row, int grp, uint64_t key) {
    if (lc.left == 0) {
        // the counter saturates far below 2^32 (the host caps the log at 2^31 records and every launch starts at 0)
        lc.cur = atomicAdd(p.col_log_cnt, static_cast<uint32_t>(kColLogSeg));
        lc.left = kColLogSeg;
    }
    if (lc.cur < p.col_log_cap)  // past the end: dropped; the host sees the counter and falls back
        p.col_log[lc.cur] = make_uint4(row, static_cast<uint32_t>(grp), static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32));
    ++lc.cur;
    --lc.left;
}

__device__ __forceinline__ void scan_chunk_col(const uint32_t (&v)[32], const float (&m8)[4], float cthr_lane, uint32_t row0,
                                               uint32_t anchor_id, uint32_t self_row, int grp, bool q_valid,
                                               const GemmTopkParams& p, ColLogCursor& lc) {
    // gmin[g]: the smallest threshold among the chunk's columns 8g .. 8g+7 (lane l holds column l's threshold)
    float gm = cthr_lane;
    gm = fminf(gm, __shfl_xor_sync(0xffffffffu, gm, 1));
    gm = fminf(gm, __shfl_xor_sync(0xffffffffu, gm, 2));
    gm = fminf(gm, __shfl_xor_sync(0xffffffffu, gm, 4));
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float gmin = __shfl_sync(0xffffffffu, gm, 8 * g);
        if (!__any_sync(0xffffffffu, q_valid && m8[g] > gmin)) continue;  // nobody beats any column of the group
#pragma unroll
        for (int j = 8 * g; j < 8 * g + 8; ++j) {
            const float tj = __shfl_sync(0xffffffffu, cthr_lane, j);  // +inf for rows that do not collect
            const float s = __uint_as_float(v[j]);
            if (q_valid && s > tj) {
                const uint32_t row = row0 + j;
                // (a row of the query's own group is refused by col_scatter_kernel: the group lookup is an L2 round
                // trip that would stall the warp here)
                if (row != self_row) col_log_push(lc, p, row, grp, make_key(s, anchor_id));
            }
        }
    }
}

// ===========================================================================
// CTA pair (cta_group::2, M = 256) with the QUERIES RESIDENT ON CHIP.
//
//   * a cluster of two CTAs (one SM pair) owns 256 queries; each CTA keeps its
//     128 query rows for the whole work item: the first KB_T 64-wide K blocks in
//     TENSOR MEMORY (bf16 pairs, 32 columns per block, written with tcgen05.st)
//     and, when K is larger, KB_S more blocks in shared memory ("tail", loaded
//     once per item by TMA) -- the A operand is never re-streamed from L2,
//   * only database rows stream: each CTA TMA-loads half of the BLOCK_N rows of
//     a tile, KB_STAGE K blocks per pipeline stage; the MMA
//     (tcgen05.mma.cta_group::2) reads both halves, with A from TMEM for the
//     first KB_T blocks and from the shared-memory tail afterwards,
//   * accumulators: two buffers of BLOCK_N fp32 columns after the A region
//     (KB_T*32 + 2*BLOCK_N == 512 TMEM columns),
//   * the epilogue is the same threshold filter, one thread per query row.
//
// Configurations: <128, 8, 0, 4, 6> for K <= 512, <128, 8, 4, 4, 5> for
// K <= 768 (hybrid), <64, 12, 0, 12, 4> (all of K <= 768 in TMEM, N = 64).
// L2 -> SM traffic per SM is a third of the streaming kernel's.
// ===========================================================================
template <int BLOCK_N, int KB_T, int KB_S, int KB_STAGE, int STAGES, int E, bool COL = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
gemm_topk_ts2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_q,
                     const __nv_bfloat16* __restrict__ q_pack, const int q_row_elems, const GemmTopkParams p) {
    constexpr int HALF_N = BLOCK_N / 2;                    // database rows this CTA loads per tile
    constexpr uint32_t KB_BYTES = HALF_N * 128;            // one 64-wide K block of those rows
    constexpr uint32_t STAGE_BYTES = KB_STAGE * KB_BYTES;  // one pipeline stage
    constexpr uint32_t A_COLS = KB_T * 32;                 // TMEM columns holding the query tile
    constexpr uint32_t TAIL_KB_BYTES = 128 * 128;          // one K block of this CTA's 128 query rows
    constexpr uint32_t TAIL_BYTES = KB_S * TAIL_KB_BYTES;
    static_assert(A_COLS + 2 * BLOCK_N == 512, "TMEM budget: queries + two accumulators = 512 columns");
    constexpr int C = 32 * (E > 0 ? E : 1);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_tail = smem;                     // [KB_S][128 rows][128 B], swizzled
    uint8_t* smem_b = smem + TAIL_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TAIL_BYTES + STAGES * STAGE_BYTES);
    uint64_t* full_bar = bars;                     // leader CTA's copy is the live one
    uint64_t* empty_bar = bars + STAGES;           // per CTA
    uint64_t* tmem_full = bars + 2 * STAGES;       // per CTA
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // leader's copy
    uint64_t* a_ready = bars + 2 * STAGES + 4;     // leader's copy: query rows are in TMEM (+ tail in smem)
    uint64_t* a_free = bars + 2 * STAGES + 5;      // per CTA: all MMAs of the item have retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();       // 0 = leader
    const int pair = blockIdx.x >> 1;
    const int n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        if constexpr (KB_S > 0) prefetch_tmap(&tmap_q);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);  // 4 epilogue warps x 2 CTAs
        }
        mbar_init(a_ready, KB_S > 0 ? 9 : 8);  // + the leader producer's expect_tx for the tails
        mbar_init(a_free, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<2>(tmem_slot, 512);
        tmem_relinquish<2>();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_items = p.q_tiles * p.n_slices;  // q_tiles counts 256-query tiles here
    const int nkb = p.nkb;                       // <= KB_T + KB_S
    const int nkb_t = min(nkb, KB_T);            // K blocks served from TMEM
    const int nkb_s = nkb - nkb_t;               // K blocks served from the smem tail
    const int n_groups = (nkb + KB_STAGE - 1) / KB_STAGE;

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer (both CTAs, warp-uniform loop)
        int stage = 0;
        uint32_t phase = 0, item_phase = 0;
        int it = 0;
        for (int w = pair; w < n_items; w += n_pairs, ++it) {
            const int slice = w / p.q_tiles, qt = w - slice * p.q_tiles;
            const int t0 = p.tile0 + slice * p.tiles_per_slice;
            const int t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
            if (p.wave_cnt != nullptr && it > 0) {
                if (elect_one_sync()) wave_wait(p.wave_cnt + it - 1, wave_members(it - 1, n_items, n_pairs, 2));
                __syncwarp();
            }
            if constexpr (KB_S > 0) {
                // the query tail of this item: wait until the previous item's MMAs are done with the old one
                mbar_wait(a_free, item_phase ^ 1);
                item_phase ^= 1;
                if (elect_one_sync()) {
                    if (rank == 0) mbar_expect_tx(a_ready, 2u * nkb_s * TAIL_KB_BYTES);
                    for (int kb = 0; kb < nkb_s; ++kb)
                        tma_load_2d_2sm(&tmap_q, a_ready, smem_tail + kb * TAIL_KB_BYTES, (KB_T + kb) * 64,
                                        qt * 256 + static_cast<int>(rank) * 128, kEvictLast);
                }
                __syncwarp();
            }
            for (int t = t0; t < t1; ++t) {
                const int row = t * BLOCK_N + static_cast<int>(rank) * HALF_N;
                for (int g = 0; g < n_groups; ++g) {
                    const int kb0 = g * KB_STAGE, kb1 = min(kb0 + KB_STAGE, nkb);
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one_sync()) {
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * (kb1 - kb0) * KB_BYTES);
                        uint8_t* dst = smem_b + stage * STAGE_BYTES;
                        for (int kb = kb0; kb < kb1; ++kb)
                            tma_load_2d_2sm(&tmap_x, &full_bar[stage], dst + (kb - kb0) * KB_BYTES, kb * 64, row,
                                            kEvictNormal);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (p.wave_cnt != nullptr && elect_one_sync()) wave_arrive(p.wave_cnt + it);
            __syncwarp();
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer (leader only, warp-uniform loop)
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0, item_phase = 0;
            for (int w = pair; w < n_items; w += n_pairs) {
                const int slice = w / p.q_tiles;
                const int t0 = p.tile0 + slice * p.tiles_per_slice;
                const int t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
                mbar_wait(a_ready, item_phase);  // both CTAs hold their query rows (TMEM + tail)
                item_phase ^= 1;
                tc_fence_after();
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                    const uint32_t tmem_d = tmem_base + A_COLS + acc * BLOCK_N;
                    for (int g = 0; g < n_groups; ++g) {
                        const int kb0 = g * KB_STAGE, kb1 = min(kb0 + KB_STAGE, nkb);
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(smem_b + stage * STAGE_BYTES));
                            // Fast path: a whole stage of full K blocks, all served from TMEM or all from the
                            // tail -> 4*KB_STAGE MMAs with compile-time operand offsets (the issue loop, not the
                            // tensor pipe, bounded short-K tiles: ncu showed the pipe 81 % active at K = 384).
                            const bool whole = (kb1 - kb0 == KB_STAGE) && (p.k16 >= 4 * kb1);
                            if (whole && (KB_S == 0 || kb1 <= KB_T)) {
                                const uint32_t a0 = tmem_base + kb0 * 32;
#pragma unroll
                                for (int i = 0; i < KB_STAGE; ++i) {
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk)
                                        umma_ts<2>(tmem_d, a0 + i * 32 + kk * 8,
                                                   b_desc0 + static_cast<uint64_t>(((i * KB_BYTES) >> 4) + 2 * kk), idesc,
                                                   (i | kk) != 0 ? 1u : static_cast<uint32_t>(g != 0));
                                }
                            } else if (whole && KB_S > 0 && kb0 >= KB_T) {
                                const uint64_t a_desc0 =
                                    make_smem_desc_sw128(smem_u32(smem_tail + (kb0 - KB_T) * TAIL_KB_BYTES));
#pragma unroll
                                for (int i = 0; i < KB_STAGE; ++i) {
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk)
                                        umma_ss<2>(tmem_d, a_desc0 + static_cast<uint64_t>(((i * TAIL_KB_BYTES) >> 4) + 2 * kk),
                                                   b_desc0 + static_cast<uint64_t>(((i * KB_BYTES) >> 4) + 2 * kk), idesc, 1u);
                                }
                            } else
                            for (int kb = kb0; kb < kb1; ++kb) {
                                const uint64_t b_desc = b_desc0 + static_cast<uint64_t>(((kb - kb0) * KB_BYTES) >> 4);
                                // the last K block may be partly padding: issue only the K=16 steps that hold data
                                const int nk = (kb == nkb - 1) ? p.k16 - 4 * kb : 4;
                                if (KB_S == 0 || kb < KB_T) {
                                    if (nk == 4) {
#pragma unroll
                                        for (int kk = 0; kk < 4; ++kk)
                                            umma_ts<2>(tmem_d, tmem_base + kb * 32 + kk * 8, b_desc + 2 * kk, idesc,
                                                       (kb | kk) != 0);
                                    } else {
                                        for (int kk = 0; kk < nk; ++kk)
                                            umma_ts<2>(tmem_d, tmem_base + kb * 32 + kk * 8, b_desc + 2 * kk, idesc,
                                                       (kb | kk) != 0);
                                    }
                                } else {
                                    const uint64_t a_desc =
                                        make_smem_desc_sw128(smem_u32(smem_tail + (kb - KB_T) * TAIL_KB_BYTES));
                                    if (nk == 4) {
#pragma unroll
                                        for (int kk = 0; kk < 4; ++kk)
                                            umma_ss<2>(tmem_d, a_desc + 2 * kk, b_desc + 2 * kk, idesc, 1u);
                                    } else {
                                        for (int kk = 0; kk < nk; ++kk)
                                            umma_ss<2>(tmem_d, a_desc + 2 * kk, b_desc + 2 * kk, idesc, 1u);
                                    }
                                }
                            }
                            umma_commit_2sm(&empty_bar[stage], 3);                        // frees the stage in both CTAs
                            if (g == n_groups - 1) umma_commit_2sm(&tmem_full[acc], 3);   // accumulator ready in both CTAs
                            if (KB_S > 0 && g == n_groups - 1 && t == t1 - 1) umma_commit_2sm(a_free, 3);
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ----------------------------------------------------------- epilogue (both CTAs)
        const int ewarp = warp - 4;
        const uint32_t lane_base = static_cast<uint32_t>(ewarp * 32) << 16;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int dbg = p.dbg;
        LaneTopk<E> st;
        ColLogCursor col_log;  // (column direction only)
        if constexpr (E > 0)
            st.buf = p.cand + (static_cast<size_t>(blockIdx.x) * 128 + ewarp * 32 + lane) * C;
        for (int w = pair; w < n_items; w += n_pairs) {
            const int slice = w / p.q_tiles, qt = w - slice * p.q_tiles;
            const int t0 = p.tile0 + slice * p.tiles_per_slice;
            const int t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
            const int q_base = qt * 256 + static_cast<int>(rank) * 128 + ewarp * 32;
            const int q_row = q_base + lane;
            const bool q_valid = q_row < p.nq;
            // ---- (1) this thread's query row -> TMEM lane (bf16 pairs, 16 columns per store).
            // Safe to overwrite: the previous item's last accumulator was consumed, so all
            // MMAs that read the old rows have retired.
            {
                const uint4* src = reinterpret_cast<const uint4*>(q_pack + static_cast<size_t>(q_valid ? q_row : 0) * q_row_elems);
                const int n_vec = q_valid ? (min(q_row_elems, nkb_t * 64) >> 3) : 0;  // 16-byte granules with data
                for (int c16 = 0; c16 < nkb_t * 2; ++c16) {
                    uint32_t r[16];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int vi = c16 * 4 + v;
                        uint4 x = make_uint4(0, 0, 0, 0);
                        if (vi < n_vec) x = __ldg(src + vi);
                        r[4 * v + 0] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
                    }
                    tmem_st16(tmem_base + lane_base + c16 * 16, r);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(a_ready, 0);
            }
            const uint32_t self = (p.self_ids && q_valid) ? static_cast<uint32_t>(__ldg(p.self_ids + q_row)) : 0xFFFFFFFFu;
            const int grp = (p.group_q && q_valid) ? __ldg(p.group_q + q_row) : -1;
            item_begin<E>(st, p, q_row, q_valid, p.cand + (static_cast<size_t>(blockIdx.x) * 128 + ewarp * 32) * C, qt, slice);
            if constexpr (E > 0) st.grp = p.group_db != nullptr ? grp : -1;
            uint32_t anchor_id = self;
            if constexpr (COL) {
                if (p.q_ids != nullptr && q_valid) anchor_id = static_cast<uint32_t>(__ldg(p.q_ids + q_row));
            }
            // column direction: the thresholds of a tile's database rows (lane l: columns l, l+32, ...), fetched one
            // tile ahead so the L2 round trip never sits between an accumulator becoming ready and its scan.
            // Tried on the 6.25M-row join and dropped (same-box A/B against this version, 25.2-25.4 s):
            //  * a precomputed per-tile minimum of the thresholds (one broadcast load per tile, row thresholds only
            //    fetched when a chunk beats it): 25.3-25.5 s, no gain;
            //  * deciding once per TILE (accumulate the chunk maxima, one vote, re-read the accumulator from tensor
            //    memory for the rare tile that fires): the accumulator has to be kept until the decision, and that
            //    later hand-off to the MMA warp cost more than the three votes saved: 26.4 s;
            //  * issuing the tensor-memory load of the next 32 columns before scanning the current 32 (two register
            //    sets, +35 registers): 29.1 s, and k = 200 on the headline shape 91.5k -> 88.8k qps.
            float cthr_next[COL ? BLOCK_N / 32 : 1];
            auto load_cthr = [&](int t, float (&out)[COL ? BLOCK_N / 32 : 1]) {
#pragma unroll
                for (int c = 0; c < (COL ? BLOCK_N / 32 : 1); ++c) {
                    const uint32_t r = static_cast<uint32_t>(t) * BLOCK_N + c * 32 + lane;
                    out[c] = (r >= static_cast<uint32_t>(p.col_row_min) && r < static_cast<uint32_t>(p.n_rows))
                                 ? __ldcg(p.col_thr + r) : INFINITY;
                }
            };
            if constexpr (COL) load_cthr(t0, cthr_next);
            for (int t = t0; t < t1; ++t) {
                const uint32_t row0 = static_cast<uint32_t>(t) * BLOCK_N;
                float cthr[COL ? BLOCK_N / 32 : 1];
                float cmin_tile = INFINITY;  // the smallest threshold among the tile's columns (+inf: nobody collects)
                if constexpr (COL) {
#pragma unroll
                    for (int c = 0; c < BLOCK_N / 32; ++c) {
                        cthr[c] = cthr_next[c];
                        cmin_tile = fminf(cmin_tile, cthr[c]);
                    }
                    if (t + 1 < t1) load_cthr(t + 1, cthr_next);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) cmin_tile = fminf(cmin_tile, __shfl_xor_sync(0xffffffffu, cmin_tile, o));
                    if (!q_valid) cmin_tile = INFINITY;  // padding query rows never offer themselves
                }
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_base + A_COLS + acc * BLOCK_N;
#pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 32) {
                    uint32_t v[32];
                    if (!(dbg & 2)) {
                        tmem_ld32(taddr + c, v);
                        tmem_ld_wait();
                    }
                    if (c + 32 == BLOCK_N) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
                    }
                    float m8c[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    if (!(dbg & 3))
                        scan_chunk<E>(st, v, row0 + c, static_cast<uint32_t>(p.n_rows), self, grp, p.group_db, p.k, p.room,
                                      COL ? m8c : nullptr);
                    if constexpr (COL) {
                        // fast path: one compare against the tile's smallest threshold and one vote per chunk
                        const float m_chunk = fmaxf(fmaxf(m8c[0], m8c[1]), fmaxf(m8c[2], m8c[3]));
                        if (__any_sync(0xffffffffu, m_chunk > cmin_tile)) {
                            float ct = cthr[0];  // (the loop is not unrolled: pick the chunk's register)
#pragma unroll
                            for (int q = 1; q < BLOCK_N / 32; ++q)
                                if (c == q * 32) ct = cthr[q];
                            scan_chunk_col(v, m8c, ct, row0 + c, anchor_id, self, grp, q_valid, p, col_log);
                        }
                    }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            item_flush<E>(st, p, q_base, q_row, q_valid, slice, 32, qt);
        }
    }

    tc_fence_before();
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still signal it
    if (warp == 2) tmem_dealloc<2>(tmem_base, 512);
}

// ===========================================================================
// CTA-pair streaming variant (cta_group::2, M = 256, N = BLOCK_N): both
// operands stream through shared memory in 64-wide K blocks like the
// single-CTA kernel, but one MMA covers 256 queries, so every database tile is
// fetched from L2 once per 256 queries and each CTA stages (and the tensor core
// reads from its shared memory) only half of the B tile.  Any K, any number of
// plane combos (exact split storage included).
// ===========================================================================
template <int BLOCK_N, int STAGES, int E>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
gemm_topk_ss2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                     const GemmTopkParams p) {
    constexpr int BLOCK_K = 64;
    constexpr int HALF_N = BLOCK_N / 2;
    constexpr uint32_t A_BYTES = 128 * BLOCK_K * 2;     // this CTA's 128 query rows
    constexpr uint32_t B_BYTES = HALF_N * BLOCK_K * 2;  // this CTA's half of the database tile
    static_assert(2 * BLOCK_N <= 512, "two accumulators must fit TMEM");
    constexpr uint32_t TMEM_COLS = 512;
    constexpr int C = 32 * (E > 0 ? E : 1);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
    uint64_t* full_bar = bars;                     // leader's copy is live
    uint64_t* empty_bar = bars + STAGES;           // per CTA
    uint64_t* tmem_full = bars + 2 * STAGES;       // per CTA
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // leader's copy
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_x);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);  // 4 epilogue warps x 2 CTAs
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<2>(tmem_slot, TMEM_COLS);
        tmem_relinquish<2>();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_items = p.q_tiles * p.n_slices;  // q_tiles counts 256-query tiles
    const int ksteps = p.nkb * p.n_combo;

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer (both CTAs)
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int w = pair; w < n_items; w += n_pairs, ++it) {
            const int slice = w / p.q_tiles, qt = w - slice * p.q_tiles;
            const int t0 = p.tile0 + slice * p.tiles_per_slice;
            const int t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
            const int q_row0 = qt * 256 + static_cast<int>(rank) * 128;
            if (p.wave_cnt != nullptr && it > 0) {
                if (elect_one_sync()) wave_wait(p.wave_cnt + it - 1, wave_members(it - 1, n_items, n_pairs, 2));
                __syncwarp();
            }
            for (int t = t0; t < t1; ++t) {
                const int x_row0 = t * BLOCK_N + static_cast<int>(rank) * HALF_N;
                for (int c = 0; c < p.n_combo; ++c) {
                    const int a_col = static_cast<int>((p.a_planes >> (4 * c)) & 0xF) * p.plane_cols;
                    const int b_col = static_cast<int>((p.b_planes >> (4 * c)) & 0xF) * p.plane_cols;
                    for (int kb = 0; kb < p.nkb; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (elect_one_sync()) {
                            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * (A_BYTES + B_BYTES));
                            tma_load_2d_2sm(&tmap_q, &full_bar[stage], smem_a + stage * A_BYTES, a_col + kb * BLOCK_K,
                                            q_row0, kEvictLast);
                            tma_load_2d_2sm(&tmap_x, &full_bar[stage], smem_b + stage * B_BYTES, b_col + kb * BLOCK_K,
                                            x_row0, kEvictNormal);
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
            if (p.wave_cnt != nullptr && elect_one_sync()) wave_arrive(p.wave_cnt + it);
            __syncwarp();
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer (leader only)
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int w = pair; w < n_items; w += n_pairs) {
                const int slice = w / p.q_tiles;
                const int t0 = p.tile0 + slice * p.tiles_per_slice;
                const int t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                    for (int s = 0; s < ksteps; ++s) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * A_BYTES));
                            const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * B_BYTES));
                            const int nk = min(4, p.k16 - 4 * (s % p.nkb));
                            if (nk == 4) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    umma_ss<2>(tmem_d, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (s | kk) != 0);
                            } else {
                                for (int kk = 0; kk < nk; ++kk)
                                    umma_ss<2>(tmem_d, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (s | kk) != 0);
                            }
                            umma_commit_2sm(&empty_bar[stage], 3);
                            if (s == ksteps - 1) umma_commit_2sm(&tmem_full[acc], 3);
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ----------------------------------------------------------- epilogue (both CTAs)
        const int ewarp = warp - 4;
        const uint32_t lane_base = static_cast<uint32_t>(ewarp * 32) << 16;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int dbg = p.dbg;
        LaneTopk<E> st;
        if constexpr (E > 0)
            st.buf = p.cand + (static_cast<size_t>(blockIdx.x) * 128 + ewarp * 32 + lane) * C;
        for (int w = pair; w < n_items; w += n_pairs) {
            const int slice = w / p.q_tiles, qt = w - slice * p.q_tiles;
            const int t0 = p.tile0 + slice * p.tiles_per_slice;
            const int t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
            const int q_base = qt * 256 + static_cast<int>(rank) * 128 + ewarp * 32;
            const int q_row = q_base + lane;
            const bool q_valid = q_row < p.nq;
            const uint32_t self = (p.self_ids && q_valid) ? static_cast<uint32_t>(__ldg(p.self_ids + q_row)) : 0xFFFFFFFFu;
            const int grp = (p.group_q && q_valid) ? __ldg(p.group_q + q_row) : -1;
            item_begin<E>(st, p, q_row, q_valid, p.cand + (static_cast<size_t>(blockIdx.x) * 128 + ewarp * 32) * C, qt, slice);
            if constexpr (E > 0) st.grp = p.group_db != nullptr ? grp : -1;
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t row0 = static_cast<uint32_t>(t) * BLOCK_N;
                const uint32_t taddr = tmem_base + lane_base + acc * BLOCK_N;
#pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 32) {
                    uint32_t v[32];
                    if (!(dbg & 2)) {
                        tmem_ld32(taddr + c, v);
                        tmem_ld_wait();
                    }
                    if (c + 32 == BLOCK_N) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
                    }
                    if (!(dbg & 3))
                        scan_chunk<E>(st, v, row0 + c, static_cast<uint32_t>(p.n_rows), self, grp, p.group_db, p.k, p.room);
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            item_flush<E>(st, p, q_base, q_row, q_valid, slice, 32, qt);
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc<2>(tmem_base, TMEM_COLS);
}

template <int BLOCK_N, int STAGES>
constexpr size_t gemm_topk_ss2_smem_bytes() {
    return 1024 + size_t(STAGES) * (128 * 64 * 2 + (BLOCK_N / 2) * 64 * 2) + (2 * STAGES + 5) * 8 + 16;
}

template <int BLOCK_N, int KB_S, int KB_STAGE, int STAGES>
constexpr size_t gemm_topk_ts2_smem_bytes() {
    return 1024 + size_t(KB_S) * 128 * 128 + size_t(STAGES) * KB_STAGE * (BLOCK_N / 2) * 128 + (2 * STAGES + 7) * 8 + 16;
}

}  // namespace cvdb
