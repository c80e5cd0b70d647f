// Fused distance GEMM + per-query top-k for sm_100a.
//
//   scores[q, r] = sum_k Q[q, k] * X[r, k]        (bf16 x bf16 -> fp32, tcgen05.mma)
//   per query keep the k best (score desc, row asc) -- the score matrix lives
//   only in TMEM, never in HBM.
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
// (select_kernels.cuh) does the final k-way select.
#pragma once
#include "ptx_sm100.cuh"
#include "topk_util.cuh"

namespace cvdb {

constexpr int kMaxCombos = 6;

struct GemmTopkParams {
    int nq;               // queries in this launch
    int n_rows;           // database rows in this launch
    int k;                // results kept per (query, slice); k <= 32*E - 8 (E>0) or 1 (E==0)
    int q_tiles;          // ceil(nq / 128)
    int n_tiles;          // ceil(n_rows / BLOCK_N)
    int n_slices;         // database slices
    int tiles_per_slice;  // in BLOCK_N units
    int nkb;              // 64-element K blocks per plane
    int n_combo;          // (A plane, B plane) pairs accumulated per tile: 1 (bf16) or 6 (exact split)
    int plane_cols;       // columns per plane (Kp)
    uint32_t a_planes;    // 4 bits per combo: plane of the A (query) operand
    uint32_t b_planes;    // 4 bits per combo: plane of the B (database) operand
    const int32_t* self_ids;  // [nq] database row to drop for each query, or null
    const int32_t* group_q;   // [nq] group id per query (<0: none), or null
    const int32_t* group_db;  // [n_rows] group id per database row, or null
    uint64_t* cand;  // [gridDim.x][128][32*E] candidate scratch (E>0)
    uint64_t* part;  // [nq][n_slices][k] per-slice results (keys)
};

// ---------------------------------------------------------------------------
// Per-thread (= per-query) selection state used by the epilogue warps.
// ---------------------------------------------------------------------------
template <int E>
struct LaneTopk {
    float thr;       // score of the current k-th best; -inf until k candidates were seen
    int cnt;         // filled slots in buf
    uint64_t* buf;   // 32*E slots, slot p of every query of the warp is zero when p >= cnt
};
template <>
struct LaneTopk<0> {
    float thr;
    uint64_t best;
};

// Sort the buffer of lane `l` (warp-cooperative), keep its best k, zero the rest.
// Returns the k-th best key (0 when fewer than k candidates exist).
template <int E>
__device__ __forceinline__ uint64_t warp_compact(uint64_t* b, int k, uint64_t (&key)[E]) {
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < E; ++e) key[e] = __ldcg(reinterpret_cast<const unsigned long long*>(b + e * 32 + lane));
    warp_bitonic_sort_desc<E>(key);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int pos = e * 32 + lane;
        if (pos >= k) key[e] = 0;
        b[pos] = key[e];
    }
    return warp_sorted_at<E>(key, k - 1);
}

template <int E>
__device__ __forceinline__ void make_room(LaneTopk<E>& st, int k) {
    constexpr int C = 32 * E;
    unsigned mask = __ballot_sync(0xffffffffu, st.cnt > C - 8);
    if (mask == 0) return;
    __syncwarp();
    const uint32_t lane = threadIdx.x & 31;
    while (mask) {
        const int l = __ffs(mask) - 1;
        mask &= mask - 1;
        uint64_t* b = reinterpret_cast<uint64_t*>(shfl_u64(reinterpret_cast<uint64_t>(st.buf), l));
        uint64_t key[E];
        const uint64_t kth = warp_compact<E>(b, k, key);
        if (static_cast<int>(lane) == l) {
            st.thr = kth ? key_score(kth) : -INFINITY;
            st.cnt = st.cnt < k ? st.cnt : k;
        }
    }
    __syncwarp();
}

// Process 32 scores (database rows row0 .. row0+31) of this thread's query.
template <int E>
__device__ __forceinline__ void scan_chunk(LaneTopk<E>& st, const uint32_t (&v)[32], uint32_t row0, uint32_t row_end,
                                           uint32_t self, int grp, const int32_t* __restrict__ group_db, int k) {
    float m = __uint_as_float(v[0]);
#pragma unroll
    for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
    if (!__any_sync(0xffffffffu, m > st.thr)) return;  // common case once the threshold has settled
#pragma unroll
    for (int g = 0; g < 32; g += 8) {
        if constexpr (E > 0) make_room<E>(st, k);
#pragma unroll
        for (int j = g; j < g + 8; ++j) {
            const float s = __uint_as_float(v[j]);
            if (s > st.thr) {
                const uint32_t row = row0 + j;
                bool ok = row < row_end && row != self;
                if (ok && grp >= 0 && group_db != nullptr) ok = __ldg(group_db + row) != grp;
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
        // ------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
                const int slice = w / p.q_tiles, qt = w - slice * p.q_tiles;
                const int t0 = slice * p.tiles_per_slice;
                const int t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
                for (int t = t0; t < t1; ++t) {
                    for (int c = 0; c < p.n_combo; ++c) {
                        const int a_col = static_cast<int>((p.a_planes >> (4 * c)) & 0xF) * p.plane_cols;
                        const int b_col = static_cast<int>((p.b_planes >> (4 * c)) & 0xF) * p.plane_cols;
                        for (int kb = 0; kb < p.nkb; ++kb) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            mbar_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
                            tma_load_2d(&tmap_q, &full_bar[stage], smem_a + stage * A_BYTES, a_col + kb * BLOCK_K,
                                        qt * BLOCK_M, kEvictLast);
                            tma_load_2d(&tmap_x, &full_bar[stage], smem_b + stage * B_BYTES, b_col + kb * BLOCK_K,
                                        t * BLOCK_N, kEvictNormal);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
                const int slice = w / p.q_tiles;
                const int t0 = slice * p.tiles_per_slice;
                const int t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                    for (int s = 0; s < ksteps; ++s) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * A_BYTES));
                        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + stage * B_BYTES));
#pragma unroll
                        for (int kk = 0; kk < BLOCK_K / 16; ++kk) {
                            // +32 bytes along K inside the 128-byte swizzled row
                            umma_ss<1>(tmem_d, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (s | kk) != 0);
                        }
                        umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ----------------------------------------------------------- epilogue
        const int ewarp = warp - 4;  // == warp % 4: TMEM lanes [32*ewarp, 32*ewarp+32)
        const uint32_t lane_base = static_cast<uint32_t>(ewarp * 32) << 16;
        int acc = 0;
        uint32_t acc_phase = 0;
        LaneTopk<E> st;
        if constexpr (E > 0)
            st.buf = p.cand + (static_cast<size_t>(blockIdx.x) * BLOCK_M + ewarp * 32 + lane) * C;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const int slice = w / p.q_tiles, qt = w - slice * p.q_tiles;
            const int t0 = slice * p.tiles_per_slice;
            const int t1 = min(t0 + p.tiles_per_slice, p.n_tiles);
            const int q_row = qt * BLOCK_M + ewarp * 32 + lane;
            const bool q_valid = q_row < p.nq;
            const uint32_t self = (p.self_ids && q_valid) ? static_cast<uint32_t>(__ldg(p.self_ids + q_row)) : 0xFFFFFFFFu;
            const int grp = (p.group_q && q_valid) ? __ldg(p.group_q + q_row) : -1;
            st.thr = -INFINITY;
            if constexpr (E > 0) {
                st.cnt = 0;
                // zero the 32 buffers of this warp (contiguous: 32*C slots)
                uint64_t* wb = p.cand + (static_cast<size_t>(blockIdx.x) * BLOCK_M + ewarp * 32) * C;
                for (int i = lane; i < 32 * C; i += 32) wb[i] = 0;
                __syncwarp();
            } else {
                st.best = 0;
            }
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t row0 = static_cast<uint32_t>(t) * BLOCK_N;
                const uint32_t taddr = tmem_base + lane_base + acc * BLOCK_N;
#pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c, v);
                    tmem_ld_wait();
                    if (c + 32 == BLOCK_N) {
                        // all of this thread's scores are in registers: hand the accumulator back
                        tc_fence_before();
                        mbar_arrive(&tmem_empty[acc]);
                    }
                    scan_chunk<E>(st, v, row0 + c, static_cast<uint32_t>(p.n_rows), self, grp, p.group_db, p.k);
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
            // ---- flush this item's result: part[q][slice][0..k)
            if constexpr (E > 0) {
                __syncwarp();
                for (int l = 0; l < 32; ++l) {
                    const int qr = qt * BLOCK_M + ewarp * 32 + l;
                    if (qr >= p.nq) break;
                    uint64_t* b = reinterpret_cast<uint64_t*>(shfl_u64(reinterpret_cast<uint64_t>(st.buf), l));
                    uint64_t key[E];
                    warp_compact<E>(b, p.k, key);
                    uint64_t* out = p.part + (static_cast<size_t>(qr) * p.n_slices + slice) * p.k;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int pos = e * 32 + lane;
                        if (pos < p.k) out[pos] = key[e];
                    }
                }
                __syncwarp();
            } else {
                if (q_valid) p.part[static_cast<size_t>(q_row) * p.n_slices + slice] = st.best;
            }
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

}  // namespace cvdb
