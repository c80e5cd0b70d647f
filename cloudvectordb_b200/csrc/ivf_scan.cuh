// Inverted-list (IVF) scan with the LIST ROWS on the M side of the MMA.
//
// A probed list is a few hundred rows and is probed by a handful of queries per
// batch (10M rows, nlist 16 384, nprobe 8, 10k queries: ~610 rows, ~5 queries).
// The grouped kernel in gemm_topk.cuh puts the queries on the M = 128 side, so
// ~92 % of every MMA multiplies padding and the few valid query rows leave most
// epilogue lanes idle.  Here the roles are swapped:
//
//   D[128 list rows][NQ queries] = X_tile[128][K] * Qg[NQ][K]^T      (tcgen05.mma, M = 128, N = NQ = 16)
//
//   * A = 128 rows of the list, streamed by TMA (K-major, 128-byte swizzle, 64-wide K blocks,
//     STAGES-deep ring): the kernel is a pure HBM stream of the list rows,
//   * B = the item's <= NQ gathered query rows, loaded ONCE per work item and resident in shared
//     memory (double-buffered across items),
//   * accumulators: four buffers of NQ fp32 TMEM columns (64 columns in all),
//   * epilogue thread t owns list row t of the tile: one tcgen05.ld.32x32b.x16 gives it the row's
//     scores against all NQ queries; a score above the query's threshold (thr[NQ], shared memory) is
//     appended to that query's candidate buffer in shared memory (slot from an atomicAdd on cnt[q]),
//   * after every tile the four epilogue warps meet at a named barrier; queries whose buffer passed
//     the trigger are compacted (select of the best k, warp w takes queries w, w+4, ...), which also
//     tightens thr[q] and publishes it to the other probes of the same query (gthr),
//   * a buffer holds CAP = 256 keys and is compacted down to <= k <= trigger <= CAP - 128 before the
//     next tile, so the <= 128 candidates a tile can add always fit: no overflow path,
//   * work items (list, group of <= NQ pairs) are handed out dynamically (atomic counter) because
//     list lengths vary; the tail tile of a list is fetched in 32-row boxes.
//
// Results are bit-identical to the grouped kernel's: same keys, same tie rule (equal scores stay
// candidates and the key decides, because rows of a list are not stored in id order).
#pragma once
#include "gemm_topk.cuh"

namespace cvdb {

struct IvfScanParams {
    const int* n_items_ptr;     // device scalar: number of work items
    unsigned int* work_counter; // device scalar, zero at launch: next item to hand out
    int k;                      // results per (query, probe), <= 128
    int trigger;                // compact a query's buffer once it holds more than this (k <= trigger <= CAP - 128)
    int nkb, k16;               // 64-wide K blocks (<= kIvfMaxKb), 16-wide MMA K steps
    const GroupItem* items;
    const int32_t* pair_query;  // [pairs] query id of each gathered row
    const int32_t* pair_dst;    // [pairs] output slot (query * nprobe + probe) of each gathered row
    const int32_t* row_ids;     // [n_rows] caller-visible id of each stored row (or null: position)
    uint64_t* part;             // [nq * nprobe][k]
    uint32_t* gthr;             // [nq]
};

constexpr int kIvfMaxKb = 13;   // K <= 832
constexpr int kIvfCap = 256;    // candidate slots per query
constexpr int kIvfSched = 4;    // depth of the work-item ring

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// the same barrier, returning the OR of `pred` over the 128 epilogue threads (identical in every thread)
__device__ __forceinline__ bool epi_barrier_or(bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %1, 0;\n\t"
        "bar.red.or.pred q, 1, 128, p;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"(static_cast<uint32_t>(pred))
        : "memory");
    return r != 0;
}

// Sort the first n (<= 32*ES) keys of a shared-memory buffer (warp-cooperative) and write the k output entries
// (missing ones as 0).  Returns the k-th best key (0 when fewer than k candidates exist).
template <int ES>
__device__ __forceinline__ uint64_t flush_sorted(const uint64_t* b, int n, uint64_t* out, int k) {
    const int lane = threadIdx.x & 31;
    uint64_t key[ES];
#pragma unroll
    for (int e = 0; e < ES; ++e) {
        const int pos = e * 32 + lane;
        key[e] = pos < n ? b[pos] : 0;
    }
    warp_bitonic_sort_desc<ES>(key);
#pragma unroll
    for (int e = 0; e < ES; ++e) {
        const int pos = e * 32 + lane;
        if (pos < k) out[pos] = key[e];
    }
    for (int pos = 32 * ES + lane; pos < k; pos += 32) out[pos] = 0;
    return k <= 32 * ES ? warp_sorted_at<ES>(key, k - 1) : 0;
}

template <int NQ, int STAGES>
constexpr size_t ivf_scan_smem_bytes() {
    return 1024 + size_t(STAGES) * 16384 + 2 * size_t(kIvfMaxKb) * NQ * 128 + size_t(NQ) * kIvfCap * 8 +
           (2 * STAGES + 4 + 8 + 2 * kIvfSched) * 8 + kIvfSched * 32 + NQ * 16 + 64;
}

template <int NQ, int STAGES>
__global__ void __launch_bounds__(256, 1)
ivf_scan_kernel(const __grid_constant__ CUtensorMap tmap_x128, const __grid_constant__ CUtensorMap tmap_x32,
                const __grid_constant__ CUtensorMap tmap_q, const IvfScanParams p) {
    static_assert(NQ == 16, "the epilogue reads one 16-column TMEM chunk per tile");
    constexpr uint32_t A_BYTES = 128 * 128;       // one K block of 128 list rows
    constexpr uint32_t QB_BYTES = NQ * 128;       // one K block of the item's queries
    constexpr int N_ACC = 4;
    constexpr uint32_t TMEM_COLS = 64;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_q = smem_a + STAGES * A_BYTES;                       // [2][kIvfMaxKb][NQ rows][128 B]
    uint64_t* cand = reinterpret_cast<uint64_t*>(smem_q + 2 * kIvfMaxKb * QB_BYTES);  // [NQ][kIvfCap]
    uint64_t* bars = cand + NQ * kIvfCap;
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* q_full = empty_bar + STAGES;        // [2]
    uint64_t* q_empty = q_full + 2;               // [2]
    uint64_t* tmem_full = q_empty + 2;            // [4]
    uint64_t* tmem_empty = tmem_full + N_ACC;     // [4]
    uint64_t* sched_full = tmem_empty + N_ACC;    // [kIvfSched]
    uint64_t* sched_empty = sched_full + kIvfSched;
    int* sched_item = reinterpret_cast<int*>(sched_empty + kIvfSched);  // [kIvfSched][8]
    float* thr_s = reinterpret_cast<float*>(sched_item + kIvfSched * 8);  // [NQ]
    int* cnt_s = reinterpret_cast<int*>(thr_s + NQ);                      // [NQ]
    int* query_s = cnt_s + NQ;                                            // [NQ]
    int* dst_s = query_s + NQ;                                            // [NQ]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dst_s + NQ);

    const int warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x128);
        prefetch_tmap(&tmap_x32);
        prefetch_tmap(&tmap_q);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&q_full[b], 1);
            mbar_init(&q_empty[b], 1);
        }
        for (int a = 0; a < N_ACC; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);  // one arrival per epilogue warp
        }
        for (int s = 0; s < kIvfSched; ++s) {
            mbar_init(&sched_full[s], 1);
            mbar_init(&sched_empty[s], 5);  // the MMA warp + four epilogue warps
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
        // ------------------------------------------------ scheduler + TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int n = 0;; ++n) {
            const int slot = n & (kIvfSched - 1);
            const uint32_t sph = (n / kIvfSched) & 1;
            mbar_wait(&sched_empty[slot], sph ^ 1);
            int a_row0 = 0, a_rows = -1, x_row0 = 0, x_rows = 0;
            if (lane == 0) {
                const unsigned w = atomicAdd(p.work_counter, 1u);
                if (w < static_cast<unsigned>(n_items)) {
                    const GroupItem it = p.items[w];
                    a_row0 = it.a_row0; a_rows = it.a_rows; x_row0 = it.x_row0; x_rows = it.x_rows;
                }
                int* si = sched_item + slot * 8;
                si[0] = a_row0; si[1] = a_rows; si[2] = x_row0; si[3] = x_rows;
                mbar_arrive(&sched_full[slot]);  // release: the item is visible to whoever sees the phase flip
            }
            a_row0 = __shfl_sync(0xffffffffu, a_row0, 0);
            a_rows = __shfl_sync(0xffffffffu, a_rows, 0);
            x_row0 = __shfl_sync(0xffffffffu, x_row0, 0);
            x_rows = __shfl_sync(0xffffffffu, x_rows, 0);
            if (a_rows < 0) break;
            // the item's queries -> buffer n & 1 (free once the MMAs of item n-2 have retired)
            const int qb = n & 1;
            mbar_wait(&q_empty[qb], ((n >> 1) & 1) ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(&q_full[qb], p.nkb * QB_BYTES);
                for (int kb = 0; kb < p.nkb; ++kb)
                    tma_load_2d(&tmap_q, &q_full[qb], smem_q + (qb * kIvfMaxKb + kb) * QB_BYTES, kb * 64, a_row0, kEvictNormal);
            }
            __syncwarp();
            const int n_tiles = (x_rows + 127) >> 7;
            for (int t = 0; t < n_tiles; ++t) {
                const int row = x_row0 + t * 128;
                const int left = x_rows - t * 128;
                const int n_box = left >= 128 ? 4 : (left + 31) >> 5;  // 32-row boxes that hold rows of the list
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one_sync()) {
                        uint8_t* dst = smem_a + stage * A_BYTES;
                        if (n_box == 4) {
                            mbar_expect_tx(&full_bar[stage], A_BYTES);
                            tma_load_2d(&tmap_x128, &full_bar[stage], dst, kb * 64, row, kEvictFirst);
                        } else {
                            mbar_expect_tx(&full_bar[stage], n_box * (A_BYTES / 4));
                            for (int j = 0; j < n_box; ++j)
                                tma_load_2d(&tmap_x32, &full_bar[stage], dst + j * (A_BYTES / 4), kb * 64, row + j * 32,
                                            kEvictFirst);
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = make_idesc_bf16(128, NQ);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int n = 0;; ++n) {
            const int slot = n & (kIvfSched - 1);
            mbar_wait(&sched_full[slot], (n / kIvfSched) & 1);
            const int a_rows = sched_item[slot * 8 + 1];
            const int x_rows = sched_item[slot * 8 + 3];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sched_empty[slot]);
            if (a_rows < 0) break;
            const int qb = n & 1;
            mbar_wait(&q_full[qb], (n >> 1) & 1);
            tc_fence_after();
            const int n_tiles = (x_rows + 127) >> 7;
            for (int t = 0; t < n_tiles; ++t) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * NQ;
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + stage * A_BYTES));
                        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_q + (qb * kIvfMaxKb + kb) * QB_BYTES));
                        const int nk = min(4, p.k16 - 4 * kb);
                        for (int kk = 0; kk < nk; ++kk)
                            umma_ss<1>(tmem_d, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (kb | kk) != 0);
                        umma_commit(&empty_bar[stage]);
                        if (kb == p.nkb - 1) {
                            umma_commit(&tmem_full[acc]);
                            if (t == n_tiles - 1) umma_commit(&q_empty[qb]);  // the item's queries are no longer read
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == N_ACC) { acc = 0; acc_phase ^= 1; }
            }
            if (n_tiles == 0 && elect_one_sync()) mbar_arrive(&q_empty[qb]);  // (items always have rows; defensive)
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue: thread t <-> list row t of the tile
        const int ewarp = warp - 4;
        const int tid = ewarp * 32 + static_cast<int>(lane);
        const uint32_t lane_base = static_cast<uint32_t>(ewarp * 32) << 16;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int k = p.k;

        // compact query q's buffer to its best k (warp-cooperative), tighten and publish its threshold
        auto compact = [&](int q) {
            uint64_t* b = cand + q * kIvfCap;
            const int n = cnt_s[q];
            uint64_t key[kIvfCap / 32];
#pragma unroll
            for (int e = 0; e < kIvfCap / 32; ++e) {
                const int pos = e * 32 + static_cast<int>(lane);
                key[e] = pos < n ? b[pos] : 0;
            }
            uint64_t kth;
            const uint64_t T = warp_select_threshold<kIvfCap / 32>(key, k, kth);
            const int kept = warp_store_survivors<kIvfCap / 32>(b, key, T);
            if (lane == 0) {
                cnt_s[q] = kept;
                const uint32_t ord = static_cast<uint32_t>(kth >> 32);
                if (ord != 0) {
                    // rows of a list are stored in no particular id order: equal scores must stay candidates
                    float t = thr_from_shared(ord);
                    const uint32_t old = atomicMax(p.gthr + query_s[q], ord);
                    if (old > ord) t = fmaxf(t, thr_from_shared(old));
                    thr_s[q] = fmaxf(thr_s[q], t);
                }
            }
            __syncwarp();
        };

        for (int n = 0;; ++n) {
            const int slot = n & (kIvfSched - 1);
            mbar_wait(&sched_full[slot], (n / kIvfSched) & 1);
            const int a_row0 = sched_item[slot * 8 + 0];
            const int a_rows = sched_item[slot * 8 + 1];
            const int x_row0 = sched_item[slot * 8 + 2];
            const int x_rows = sched_item[slot * 8 + 3];
            __syncwarp();
            if (lane == 0) mbar_arrive(&sched_empty[slot]);
            if (a_rows < 0) break;
            // ---- item setup (the previous item's flush ended with a barrier)
            if (tid < NQ) {
                const bool valid = tid < a_rows;
                const int query = valid ? __ldg(p.pair_query + a_row0 + tid) : 0;
                query_s[tid] = query;
                dst_s[tid] = valid ? __ldg(p.pair_dst + a_row0 + tid) : -1;
                cnt_s[tid] = 0;
                thr_s[tid] = valid ? thr_from_shared(__ldcg(p.gthr + query)) : INFINITY;  // padding columns never collect
            }
            epi_barrier();
            const int n_tiles = (x_rows + 127) >> 7;
            const int row_end = x_row0 + x_rows;
            for (int t = 0; t < n_tiles; ++t) {
                const int r = x_row0 + t * 128 + tid;
                const bool row_ok = r < row_end;
                const uint32_t id = row_ok ? (p.row_ids != nullptr ? static_cast<uint32_t>(__ldg(p.row_ids + r))
                                                                   : static_cast<uint32_t>(r))
                                           : 0u;
                float thr[NQ];
#pragma unroll
                for (int q4 = 0; q4 < NQ / 4; ++q4) {
                    const float4 v4 = reinterpret_cast<const float4*>(thr_s)[q4];
                    thr[4 * q4] = v4.x; thr[4 * q4 + 1] = v4.y; thr[4 * q4 + 2] = v4.z; thr[4 * q4 + 3] = v4.w;
                }
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                uint32_t v[NQ];
                tmem_ld16(tmem_base + lane_base + acc * NQ, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                if (++acc == N_ACC) { acc = 0; acc_phase ^= 1; }
                bool filled = false;  // one of my candidates took a buffer past the trigger
                if (row_ok) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const float s = __uint_as_float(v[q]);
                        if (s > thr[q]) {
                            const int pos = atomicAdd(&cnt_s[q], 1);
                            cand[q * kIvfCap + pos] = make_key(s, id);
                            filled |= pos >= p.trigger;
                        }
                    }
                }
                // The barrier itself carries the decision (OR over the 128 threads), so every thread takes the same
                // branch even though fast threads may already be pushing candidates of the next tile.
                if (epi_barrier_or(filled)) {
                    const int c = cnt_s[lane & (NQ - 1)];
                    const uint32_t over = __ballot_sync(0xffffffffu, c > p.trigger) & ((1u << NQ) - 1u);
#pragma unroll
                    for (int j = 0; j < NQ / 4; ++j) {
                        const int q = ewarp + 4 * j;
                        if ((over >> q) & 1u) compact(q);
                    }
                    epi_barrier();
                }
            }
            // ---- flush: sorted top-k of every pair of the item -> part[dst][0..k)
#pragma unroll 1
            for (int j = 0; j < NQ / 4; ++j) {
                const int q = ewarp + 4 * j;
                const int dst = dst_s[q];
                if (dst < 0) continue;  // warp-uniform
                const uint64_t* b = cand + q * kIvfCap;
                const int cn = cnt_s[q];
                uint64_t* out = p.part + static_cast<size_t>(dst) * k;
                uint64_t kth;
                // a buffer usually holds a few dozen candidates when its list ends: sort only what is there
                if (cn <= 32) {
                    kth = flush_sorted<1>(b, cn, out, k);
                } else if (cn <= 64) {
                    kth = flush_sorted<2>(b, cn, out, k);
                } else {
                    kth = flush_sorted<kIvfCap / 32>(b, cn, out, k);
                }
                if (lane == 0 && kth != 0) atomicMax(p.gthr + query_s[q], static_cast<uint32_t>(kth >> 32));
            }
            epi_barrier();  // nobody resets the per-query state while another warp still flushes
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<1>(tmem_base, TMEM_COLS);
}

}  // namespace cvdb
