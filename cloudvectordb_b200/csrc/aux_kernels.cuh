// CUDA-core kernels around the fused GEMM+top-k: row packing (fp32/bf16 ->
// padded bf16 planes, norms), the k-way merges, and the k-means update.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "topk_util.cuh"

namespace cvdb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint64_t other = shfl_xor_u64(v, o);
        v = other > v ? other : v;
    }
    return v;
}
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }

// ---------------------------------------------------------------------------
// pack_rows: one warp per row.
//   planes == 1 : out[r][0:d] = bf16(x)                       (row width Kp)
//   planes == 3 : out[r][p*Kp + c] = p-th bf16 piece of x     (x == hi+mid+lo to 24 bits)
//   L2 metric   : three extra columns right after d in plane 0 carry the norm
//                 term so that the GEMM itself yields q.x - |x|^2/2:
//                   database row: bf16 split of -|x|^2/2 ; query row: 1, 1, 1
//   norms[r] = |x|^2 of the values actually stored (optional).
// ---------------------------------------------------------------------------
// Rows n .. n_pad-1 of `out` are written as zeros (query matrices are padded to whole tiles so that the
// A-operand TMA boxes never run out of bounds).
template <typename Tin>
__global__ void pack_rows_kernel(const Tin* __restrict__ in, int64_t n, int64_t n_pad, int d, int64_t in_stride,
                                 __nv_bfloat16* __restrict__ out, int Kp, int planes, int l2, int is_query,
                                 float* __restrict__ norms, unsigned long long* __restrict__ bad_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int row_elems = planes * Kp;
    for (int64_t r = warp0; r < n_pad; r += nwarps) {
        __nv_bfloat16* dst = out + r * row_elems;
        if (r >= n) {
            for (int c = lane; c < row_elems; c += 32) dst[c] = __float2bfloat16_rn(0.f);
            continue;
        }
        const Tin* src = in + r * in_stride;
        float nrm = 0.f;
        for (int c = lane; c < Kp; c += 32) {
            float x = c < d ? to_f32(src[c]) : 0.f;
            const __nv_bfloat16 hi = __float2bfloat16_rn(x);
            // the three norm columns of plane 0 are written by lane 0 below
            if (!(l2 && c >= d && c < d + 3)) dst[c] = hi;
            if (planes == 3) {
                const float r1 = x - __bfloat162float(hi);
                const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
                const float r2 = r1 - __bfloat162float(mid);
                const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
                dst[Kp + c] = mid;
                dst[2 * Kp + c] = lo;
                const float stored = __bfloat162float(hi) + __bfloat162float(mid) + __bfloat162float(lo);
                nrm = fmaf(stored, stored, nrm);
            } else {
                const float stored = __bfloat162float(hi);
                nrm = fmaf(stored, stored, nrm);
            }
        }
        nrm = warp_sum(nrm);
        if (lane == 0) {
            if (norms) norms[r] = nrm;
            // a NaN or infinite element makes the squared norm non-finite: count such rows (cvdb_index_nonfinite_rows)
            if (bad_rows && !isfinite(nrm)) atomicAdd(bad_rows, 1ull);
            if (l2) {
                if (is_query) {
                    const __nv_bfloat16 one = __float2bfloat16_rn(1.f);
                    dst[d] = one; dst[d + 1] = one; dst[d + 2] = one;
                } else {
                    const float t = -0.5f * nrm;
                    const __nv_bfloat16 a = __float2bfloat16_rn(t);
                    const float r1 = t - __bfloat162float(a);
                    const __nv_bfloat16 b = __float2bfloat16_rn(r1);
                    const float r2 = r1 - __bfloat162float(b);
                    dst[d] = a; dst[d + 1] = b; dst[d + 2] = __float2bfloat16_rn(r2);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// merge_partials: final k-way merge of the per-slice (or per-probe, or per-rank)
// lists of one query, one warp per query.  List l of query q starts at
// part[q * q_stride + l * l_stride] and holds k_in keys.
//   KEYS_OUT = false: keys -> (D, I).  IP: D = score.  L2: D = |q|^2 - 2*score
//                     (squared distance).  Missing results: I = -1, D = -inf (IP) / +inf (L2).
//   KEYS_OUT = true : keys -> keys_out[q][k] with the id rewritten to the caller's id
//                     (row_ids translation, + id_base); missing results are key 0.  This is what
//                     ranks exchange in a sharded search: 8 bytes per candidate, still sorted, still
//                     mergeable by this same kernel.
// row_ids (optional): caller-visible id of stored row r for r < row_ids_n (an index whose rows were
// re-stored list-major keeps returning the ids the rows were added under).
// Every list is sorted descending with its empty slots (key 0) at the end, so
// this is a head-pointer merge: lane l owns lists l, l+32, ...; per output rank
// each lane offers the best head among its lists, the warp picks the largest
// and the owning lane advances that head.  Cost per query: k * n_lists / 32
// loads per lane.
// Dynamic shared memory: blockDim.x/32 * n_lists uint16 head positions.
// ---------------------------------------------------------------------------
template <typename Tidx, bool KEYS_OUT>
__global__ void merge_partials_kernel(const uint64_t* __restrict__ part, int64_t nq, int n_lists, int k_in, int k, int l2,
                                      const float* __restrict__ qnorm, int64_t id_base, int64_t q_stride,
                                      int64_t l_stride, const int32_t* __restrict__ row_ids, int64_t row_ids_n,
                                      float* __restrict__ D, Tidx* __restrict__ I, uint64_t* __restrict__ keys_out) {
    extern __shared__ uint16_t s_heads[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t q = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    uint16_t* head = s_heads + static_cast<size_t>(wib) * n_lists;
    for (int l = lane; l < n_lists; l += 32) head[l] = 0;
    __syncwarp();
    const uint64_t* c = part + q * q_stride;
    for (int r = 0; r < k; ++r) {
        uint64_t best = 0;
        int bl = -1;
        for (int l = lane; l < n_lists; l += 32) {
            const int h = head[l];
            if (h < k_in) {
                const uint64_t key = c[static_cast<int64_t>(l) * l_stride + h];
                if (key > best) { best = key; bl = l; }
            }
        }
        const uint64_t win = warp_max_u64(best);
        if (win == 0) {  // every list is exhausted: pad the rest
            for (int rr = r + lane; rr < k; rr += 32) {
                if constexpr (KEYS_OUT) {
                    keys_out[q * k + rr] = 0;
                } else {
                    D[q * k + rr] = l2 ? INFINITY : -INFINITY;
                    I[q * k + rr] = static_cast<Tidx>(-1);
                }
            }
            break;
        }
        // A key can sit in several lists (a slice that started from an earlier slice's result carries those keys on):
        // every copy is at the head of its list right now, so drop them all; the owners write the same output.
        if (best == win) {
            for (int l = lane; l < n_lists; l += 32) {
                const int h = head[l];
                if (h < k_in && c[static_cast<int64_t>(l) * l_stride + h] == win) head[l] = static_cast<uint16_t>(h + 1);
            }
            const uint32_t row = key_row(win);
            const int64_t id = ((row_ids != nullptr && row < row_ids_n) ? static_cast<int64_t>(row_ids[row])
                                                                         : static_cast<int64_t>(row)) + id_base;
            if constexpr (KEYS_OUT) {
                keys_out[q * k + r] = (win & 0xFFFFFFFF00000000ull) | static_cast<uint32_t>(~static_cast<uint32_t>(id));
            } else {
                const float s = key_score(win);
                D[q * k + r] = l2 ? (qnorm[q] - 2.f * s) : s;
                I[q * k + r] = static_cast<Tidx>(id);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// merge_lists: k-way select over `nlists` (D, I) result lists per query, as
// produced by different shards (layout [nlists][nq][k_in]).  Order: better
// distance first, then lower id; ids < 0 are padding.  One warp per query.
// ---------------------------------------------------------------------------
__global__ void merge_lists_kernel(const float* __restrict__ Dc, const int64_t* __restrict__ Ic, int64_t nq, int nlists,
                                   int k_in, int k, int l2, float* __restrict__ D, int64_t* __restrict__ I) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    const int n_cand = nlists * k_in;
    // previous pick as (ordered score, id); start above everything
    uint32_t prev_s = 0xFFFFFFFFu;
    int64_t prev_id = -1;
    bool first = true;
    for (int r = 0; r < k; ++r) {
        uint32_t bs = 0;
        int64_t bid = INT64_MAX;
        bool found = false;
        for (int i = lane; i < n_cand; i += 32) {
            const int l = i / k_in, j = i - l * k_in;
            const int64_t off = (static_cast<int64_t>(l) * nq + q) * k_in + j;
            const int64_t id = Ic[off];
            if (id < 0) continue;
            const float dv = Dc[off];
            const uint32_t s = float_to_ordered(l2 ? -dv : dv);
            // strictly after the previous pick in (score desc, id asc) order
            const bool after = first || s < prev_s || (s == prev_s && id > prev_id);
            if (!after) continue;
            if (!found || s > bs || (s == bs && id < bid)) { bs = s; bid = id; found = true; }
        }
        // warp argmax on (bs desc, bid asc)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t os = __shfl_xor_sync(0xffffffffu, bs, o);
            const int64_t oid = static_cast<int64_t>(shfl_xor_u64(static_cast<uint64_t>(bid), o));
            const int of = __shfl_xor_sync(0xffffffffu, static_cast<int>(found), o);
            if (of && (!found || os > bs || (os == bs && oid < bid))) { bs = os; bid = oid; found = true; }
        }
        if (lane == 0) {
            if (!found) {
                D[q * k + r] = l2 ? INFINITY : -INFINITY;
                I[q * k + r] = -1;
            } else {
                const float s = ordered_to_float(bs);
                D[q * k + r] = l2 ? -s : s;
                I[q * k + r] = bid;
            }
        }
        if (!found) {
            for (int rr = r + 1 + lane; rr < k; rr += 32) {
                D[q * k + rr] = l2 ? INFINITY : -INFINITY;
                I[q * k + rr] = -1;
            }
            break;
        }
        prev_s = bs; prev_id = bid; first = false;
    }
}

// ---------------------------------------------------------------------------
// k-means update: sums[assign[i]] += x[i], counts[assign[i]] += 1.
// One warp per point; fp32 vector reductions at L2 (red.global.add.v4.f32).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename Tin>
__global__ void kmeans_update_kernel(const Tin* __restrict__ x, int64_t n, int d, const int32_t* __restrict__ assign,
                                     float* __restrict__ sums, int32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t i = warp0; i < n; i += nwarps) {
        const int a = assign[i];
        if (a < 0) continue;
        const Tin* src = x + i * d;
        float* dst = sums + static_cast<int64_t>(a) * d;
        if ((d & 3) == 0) {
            for (int c = lane * 4; c < d; c += 128)
                red_add_v4(dst + c, to_f32(src[c]), to_f32(src[c + 1]), to_f32(src[c + 2]), to_f32(src[c + 3]));
        } else {
            for (int c = lane; c < d; c += 32) atomicAdd(dst + c, to_f32(src[c]));
        }
        if (lane == 0) atomicAdd(counts + a, 1);
    }
}

// new centroid = sums / counts where counts > 0, else keep the old centroid
__global__ void kmeans_finalize_kernel(const float* __restrict__ sums, const int32_t* __restrict__ counts, int K, int d,
                                       float* __restrict__ centroids) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<int64_t>(K) * d) return;
    const int c = counts[i / d];
    if (c > 0) centroids[i] = sums[i] / static_cast<float>(c);
}

// Empty clusters after an update (FAISS convention: an empty cluster takes half of a big one).  One block walks the
// empty clusters in ascending order; each takes the currently largest cluster j (ties -> lower id, at least two
// points), copies its centroid with a symmetric relative perturbation of eps (even dimensions up / odd down, the
// donor the other way round) and half of its count.  Deterministic, so every rank of a sharded run -- which sees
// the same all-reduced counts -- does the same thing.  n_split (optional) receives the number of clusters re-seeded.
__global__ void __launch_bounds__(1024, 1)
kmeans_split_empty_kernel(float* __restrict__ centroids, int32_t* __restrict__ counts, int K, int d, float eps,
                          int32_t* __restrict__ n_split) {
    __shared__ unsigned long long s_key[32];
    __shared__ unsigned char s_flag[1024];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int done = 0, split = 0;
    for (int base = 0; base < K && !done; base += 1024) {
        const int idx = base + tid;
        const int empty = idx < K && counts[idx] == 0;
        if (!__syncthreads_or(empty)) continue;
        s_flag[tid] = static_cast<unsigned char>(empty);
        __syncthreads();
        for (int i = 0; i < 1024 && !done; ++i) {
            if (!s_flag[i]) continue;  // uniform
            const int e = base + i;
            // block-wide argmax of counts, ties -> lower id
            unsigned long long best = 0;
            for (int j = tid; j < K; j += 1024) {
                const unsigned long long key =
                    (static_cast<unsigned long long>(static_cast<uint32_t>(counts[j])) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(j));
                best = key > best ? key : best;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other > best ? other : best;
            }
            if (lane == 0) s_key[warp] = best;
            __syncthreads();
            best = s_key[0];
            for (int w = 1; w < 32; ++w) best = s_key[w] > best ? s_key[w] : best;
            const int cj = static_cast<int>(best >> 32);
            const int j = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(best & 0xFFFFFFFFu));
            if (cj < 2) {
                done = 1;  // nothing left to split (uniform: every thread computed the same key)
            } else {
                for (int c = tid; c < d; c += 1024) {
                    const float v = centroids[static_cast<size_t>(j) * d + c];
                    const float up = v * (1.f + eps), down = v * (1.f - eps);
                    centroids[static_cast<size_t>(e) * d + c] = (c & 1) ? down : up;
                    centroids[static_cast<size_t>(j) * d + c] = (c & 1) ? up : down;
                }
                if (tid == 0) {
                    counts[e] = cj / 2;
                    counts[j] = cj - cj / 2;
                }
                ++split;
            }
            __syncthreads();  // counts / centroids / s_key settled before the next argmax
        }
        __syncthreads();
    }
    if (tid == 0 && n_split != nullptr) *n_split = split;
}

// ---------------------------------------------------------------------------
// Inverted-list (IVF) plumbing.
// ---------------------------------------------------------------------------
// out[p] = in[perm[p]] for whole packed rows (row_bytes is a multiple of 16)
__global__ void permute_rows_kernel(const uint4* __restrict__ in, const int32_t* __restrict__ perm, int64_t n,
                                    int row_vec16, uint4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n; r += nwarps) {
        const uint4* src = in + static_cast<int64_t>(perm[r]) * row_vec16;
        uint4* dst = out + r * row_vec16;
        for (int c = lane; c < row_vec16; c += 32) dst[c] = src[c];
    }
}

// rows per list, for the rows as currently stored: list of stored row p = list_of_id[row_ids[p]]
__global__ void ivf_count_rows_kernel(const int32_t* __restrict__ list_of_id, const int32_t* __restrict__ row_ids,
                                      int64_t n, int nlist, int32_t* __restrict__ cnt, int32_t* __restrict__ bad) {
    const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int l = list_of_id[row_ids[p]];
    if (l < 0 || l >= nlist) { atomicAdd(bad, 1); return; }
    atomicAdd(cnt + l, 1);
}

// one warp per stored row: claim a slot in its list (order inside a list is arbitrary; candidate keys carry
// the row id, so results do not depend on it) and move the packed row and its id there
__global__ void ivf_scatter_rows_kernel(const int32_t* __restrict__ list_of_id, const int32_t* __restrict__ row_ids,
                                        int64_t n, int nlist, const int32_t* __restrict__ list_off,
                                        int32_t* __restrict__ cursor, const uint4* __restrict__ x_old, int row_vec16,
                                        uint4* __restrict__ x_new, int32_t* __restrict__ row_ids_new) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (p >= n) return;
    const int id = row_ids[p];
    const int l = list_of_id[id];
    if (l < 0 || l >= nlist) return;
    int pos = 0;
    if (lane == 0) pos = list_off[l] + atomicAdd(cursor + l, 1);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (lane == 0) row_ids_new[pos] = id;
    const uint4* src = x_old + p * row_vec16;
    uint4* dst = x_new + static_cast<int64_t>(pos) * row_vec16;
    for (int c = lane; c < row_vec16; c += 32) dst[c] = src[c];
}

__global__ void fill_f32_kernel(float* __restrict__ out, int64_t n, float v) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

__global__ void iota_kernel(int32_t* __restrict__ out, int64_t first, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = static_cast<int32_t>(first + i);
}

// list_off[l] = exclusive prefix (from the scan), list_off[nlist] = total
__global__ void ivf_close_offsets_kernel(int32_t* __restrict__ list_off, int nlist, const int32_t* __restrict__ scal) {
    if (threadIdx.x == 0 && blockIdx.x == 0) list_off[nlist] = scal[1];
}

// (query, probe) pairs per list; pairs naming an invalid or empty list are dropped
__global__ void ivf_count_pairs_kernel(const int32_t* __restrict__ probes, int64_t n_pairs, int nlist,
                                       const int32_t* __restrict__ list_off, int32_t* __restrict__ cnt) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const int l = probes[i];
    if (l < 0 || l >= nlist || list_off[l + 1] == list_off[l]) return;
    atomicAdd(cnt + l, 1);
}

// single block: exclusive scans over the lists -> pair_off[l] (first gathered row of list l),
// item_off[l] (first work item of list l; an item takes up to item_cap pairs of one list);
// scal[0] = total items, scal[1] = total gathered rows, scal[3] = 0 (the scan kernel's work counter)
__global__ void ivf_scan_lists_kernel(const int32_t* __restrict__ cnt, int nlist, int32_t* __restrict__ pair_off,
                                      int32_t* __restrict__ item_off, int32_t* __restrict__ scal, int item_cap = 128) {
    __shared__ int s_pairs[1024], s_items[1024];
    __shared__ int base_pairs, base_items;
    const int tid = threadIdx.x;
    if (tid == 0) { base_pairs = 0; base_items = 0; }
    __syncthreads();
    for (int l0 = 0; l0 < nlist; l0 += 1024) {
        const int l = l0 + tid;
        const int c = l < nlist ? cnt[l] : 0;
        const int it = (c + item_cap - 1) / item_cap;
        s_pairs[tid] = c;
        s_items[tid] = it;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
            const int a = tid >= o ? s_pairs[tid - o] : 0, b = tid >= o ? s_items[tid - o] : 0;
            __syncthreads();
            s_pairs[tid] += a;
            s_items[tid] += b;
            __syncthreads();
        }
        if (l < nlist) {
            pair_off[l] = base_pairs + s_pairs[tid] - c;
            item_off[l] = base_items + s_items[tid] - it;
        }
        __syncthreads();
        if (tid == 1023) { base_pairs += s_pairs[1023]; base_items += s_items[1023]; }
        __syncthreads();
    }
    if (tid == 0) { scal[0] = base_items; scal[1] = base_pairs; scal[3] = 0; }
}

// work items of every probed list: up to item_cap gathered query rows x the rows of the list
struct IvfItem { int a_row0, a_rows, x_row0, x_rows; };
__global__ void ivf_make_items_kernel(const int32_t* __restrict__ cnt, const int32_t* __restrict__ pair_off,
                                      const int32_t* __restrict__ item_off, const int32_t* __restrict__ list_off,
                                      int nlist, IvfItem* __restrict__ items, int item_cap = 128) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlist) return;
    const int c = cnt[l];
    for (int i = 0; i * item_cap < c; ++i) {
        IvfItem it;
        it.a_row0 = pair_off[l] + i * item_cap;
        it.a_rows = min(item_cap, c - i * item_cap);
        it.x_row0 = list_off[l];
        it.x_rows = list_off[l + 1] - list_off[l];
        items[item_off[l] + i] = it;
    }
}

// one warp per pair: claim a slot next to the other pairs of the same list, record where the
// pair's result goes, and copy the packed query row there
__global__ void ivf_scatter_pairs_kernel(const int32_t* __restrict__ probes, int64_t n_pairs, int nprobe, int nlist,
                                         const int32_t* __restrict__ list_off, const int32_t* __restrict__ pair_off,
                                         int32_t* __restrict__ cursor, const uint4* __restrict__ q_pack, int row_vec16,
                                         int32_t* __restrict__ pair_query, int32_t* __restrict__ pair_dst,
                                         uint4* __restrict__ qg) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_pairs) return;
    const int l = probes[i];
    if (l < 0 || l >= nlist || list_off[l + 1] == list_off[l]) return;
    int pos = 0;
    if (lane == 0) pos = pair_off[l] + atomicAdd(cursor + l, 1);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    const int q = static_cast<int>(i / nprobe);
    if (lane == 0) {
        pair_query[pos] = q;
        pair_dst[pos] = static_cast<int32_t>(i);
    }
    const uint4* src = q_pack + static_cast<int64_t>(q) * row_vec16;
    uint4* dst = qg + static_cast<int64_t>(pos) * row_vec16;
    for (int c = lane; c < row_vec16; c += 32) dst[c] = src[c];
}

// ---------------------------------------------------------------------------
// Triplet assembly from mined neighbours (README.md:2 "dataset of triplets"):
// anchor i, its positive pos[i], and up to `per_anchor` hard negatives taken
// from ranks [skip_top, ...) of the anchor's mined list, keeping only rows whose
// score is on the "not a false negative" side of `limit`
// (IP: score <= limit, L2: distance >= limit).  Fixed-stride output
// [n][per_anchor][3] int64, unused slots are -1: no compaction, deterministic.
// ---------------------------------------------------------------------------
__global__ void build_triplets_kernel(const int64_t* __restrict__ I, const float* __restrict__ D, int64_t n, int k,
                                      const int64_t* __restrict__ pos, int64_t anchor_base, int skip_top,
                                      int per_anchor, int l2, float limit, int use_limit,
                                      int64_t* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t* o = out + i * per_anchor * 3;
    int w = 0;
    const int64_t p = pos[i];
    if (p >= 0) {
        for (int r = skip_top; r < k && w < per_anchor; ++r) {
            const int64_t id = I[i * k + r];
            if (id < 0) break;
            const float s = D[i * k + r];
            if (use_limit && (l2 ? (s < limit) : (s > limit))) continue;
            if (id == p) continue;
            o[3 * w] = anchor_base + i;
            o[3 * w + 1] = p;
            o[3 * w + 2] = id;
            ++w;
        }
    }
    for (; w < per_anchor; ++w) { o[3 * w] = -1; o[3 * w + 1] = -1; o[3 * w + 2] = -1; }
}

// ---------------------------------------------------------------------------
// Symmetric self-join, column direction (gemm_topk.cuh, scan_chunk_col): per database row a buffer of
// kColCap keys, the first col_base[r] of which are the survivors of the previous compaction.
// col_compact: one warp per row that received candidates since then: keep the best k (select, not sort), publish
// the row's new threshold (same-group candidates were refused when they were offered).  A row whose counter ran past the buffer lost
// candidates: it is flagged dirty (the caller recomputes it exactly) and stops collecting.
// ---------------------------------------------------------------------------
constexpr int kColCap = 256;

__global__ void col_compact_kernel(uint64_t* __restrict__ col_buf, uint32_t* __restrict__ col_cnt,
                                   uint32_t* __restrict__ col_base, float* __restrict__ col_thr,
                                   uint8_t* __restrict__ dirty, int k, int64_t row_min, int64_t n_rows) {
    constexpr int E = kColCap / 32;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    // A warp takes 32 consecutive rows at a time: every lane looks at one row's counters (coalesced), then the rows
    // that did receive candidates are compacted one after the other by the whole warp.  (Measured on a steady-state
    // chunk of the 6.25M-row join: 2.2 ms per chunk either way -- about two thirds of the rows receive a candidate per
    // chunk, so the selects themselves are the cost: 0.9 % of the chunk.)
    for (int64_t r0 = row_min + warp0 * 32; r0 < n_rows; r0 += nwarps * 32) {
        const int64_t rl = r0 + lane;
        uint32_t cnt_l = 0, base_l = 0;
        if (rl < n_rows) {
            cnt_l = col_cnt[rl];
            base_l = col_base[rl];
        }
        unsigned todo = __ballot_sync(0xffffffffu, cnt_l != base_l);
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t r = r0 + l;
            const uint32_t cnt = __shfl_sync(0xffffffffu, cnt_l, l);
            if (cnt > static_cast<uint32_t>(kColCap)) {
                if (lane == 0) {
                    dirty[r] = 1;
                    col_thr[r] = INFINITY;
                    col_base[r] = cnt;
                }
                continue;
            }
            uint64_t* b = col_buf + r * kColCap;
            uint64_t key[E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const uint32_t pos = e * 32 + lane;
                key[e] = pos < cnt ? b[pos] : 0;
            }
            uint64_t kth;
            const uint64_t T = warp_select_threshold<E>(key, k, kth);
            const int kept = warp_store_survivors<E>(b, key, T);
            if (lane == 0) {
                col_cnt[r] = kept;
                col_base[r] = kept;
                const uint32_t ord = static_cast<uint32_t>(kth >> 32);
                // equal scores stay candidates (the key decides): one step below the k-th score
                if (ord != 0) col_thr[r] = ord == 0x80000000u ? -1.17549435e-38f : ordered_to_float(ord - 1u);
            }
        }
    }
}

// Distribute the log of column candidates (gemm_topk.cuh, col_log_push) to the rows' buffers: one thread per record,
// same-group candidates dropped, slot from an atomicAdd on the row's counter (a counter past kColCap marks the row as overflowed); consumed
// records are zeroed so that the unused tail of a thread's last segment reads as empty next time.
__global__ void col_scatter_kernel(uint4* __restrict__ log, const uint32_t* __restrict__ log_cnt,
                                   uint32_t log_cap, uint64_t* __restrict__ col_buf,
                                   uint32_t* __restrict__ col_cnt, const int32_t* __restrict__ group_db) {
    const size_t n = min(*log_cnt, log_cap);
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint4 r = log[i];
        const uint64_t key = (static_cast<uint64_t>(r.w) << 32) | r.z;
        if (key == 0) continue;
        log[i] = make_uint4(0u, 0u, 0u, 0u);
        // record = {row, the candidate's group, key}: a row does not take a candidate of its own group
        if (group_db != nullptr && static_cast<int>(r.y) >= 0 && group_db[r.x] == static_cast<int>(r.y)) continue;
        const uint32_t pos = atomicAdd(col_cnt + r.x, 1u);
        if (pos < static_cast<uint32_t>(kColCap)) col_buf[static_cast<size_t>(r.x) * kColCap + pos] = key;
    }
}
// after the scatter: a log that ran over its capacity lost candidates (overflow[0] = 1); the counter restarts
__global__ void col_log_reset_kernel(uint32_t* __restrict__ log_cnt, uint32_t log_cap,
                                     unsigned long long* __restrict__ overflow) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (*log_cnt > log_cap) *overflow = 1ull;
        *log_cnt = 0u;
    }
}

// Seeding: the sorted top-k keys of row r among the seed anchors (a plain row-direction search) become the first
// entries of its column list; a full list sets the row's threshold.  One warp per row.
__global__ void col_seed_kernel(const uint64_t* __restrict__ keys, int64_t row0, int64_t n, int k,
                                uint64_t* __restrict__ col_buf, uint32_t* __restrict__ col_cnt,
                                uint32_t* __restrict__ col_base, float* __restrict__ col_thr) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const int64_t r = row0 + i;
    int cnt = 0;
    for (int pos = lane; pos < k; pos += 32) {
        const uint64_t key = keys[i * k + pos];
        col_buf[r * kColCap + pos] = key;
        cnt += key != 0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) {
        col_cnt[r] = cnt;
        col_base[r] = cnt;
        if (cnt >= k) {
            const uint32_t ord = static_cast<uint32_t>(keys[i * k + k - 1] >> 32);
            col_thr[r] = ord == 0x80000000u ? -1.17549435e-38f : ordered_to_float(ord - 1u);
        }
    }
}

// Final answer of anchor i = top-k of (its row-direction keys, sorted, k of them) and (its column list, <= k keys,
// unsorted).  The two sets are disjoint (rows at or after the anchor's chunk / anchors of earlier chunks).
// One warp per anchor; k <= 124.
__global__ void selfjoin_finalize_kernel(const uint64_t* __restrict__ row_keys, const uint64_t* __restrict__ col_buf,
                                         const uint32_t* __restrict__ col_cnt, int64_t row0, int64_t n, int k,
                                         float* __restrict__ D, int64_t* __restrict__ I) {
    constexpr int E = 8;
    const int lane = threadIdx.x & 31;
    const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const uint32_t nc = min(col_cnt[row0 + i], static_cast<uint32_t>(k));
    uint64_t key[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int pos = e * 32 + lane;
        uint64_t v = 0;
        if (pos < k) v = row_keys[i * k + pos];
        else if (pos - k < static_cast<int>(nc)) v = col_buf[(row0 + i) * kColCap + (pos - k)];
        key[e] = v;
    }
    warp_bitonic_sort_desc<E>(key);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int pos = e * 32 + lane;
        if (pos < k) {
            const uint64_t v = key[e];
            D[i * k + pos] = v != 0 ? key_score(v) : -INFINITY;
            I[i * k + pos] = v != 0 ? static_cast<int64_t>(key_row(v)) : -1;
        }
    }
}

// rows flagged dirty -> compact list (order unspecified)
__global__ void collect_flagged_kernel(const uint8_t* __restrict__ flag, int64_t n, int32_t* __restrict__ out,
                                       int64_t max_out, unsigned long long* __restrict__ count) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const unsigned long long pos = atomicAdd(count, 1ull);
    if (static_cast<int64_t>(pos) < max_out) out[pos] = static_cast<int32_t>(i);
}

// Exact-mode rescoring: recompute the score of each returned row from the
// three bf16 planes (their sum reproduces the fp32 value) with fp32 FMAs, so
// the reported distance does not carry tensor-core accumulation order effects.
// One warp per (query, rank).
template <typename Tidx>
__global__ void rescore_exact_kernel(const __nv_bfloat16* __restrict__ xq, const __nv_bfloat16* __restrict__ xb,
                                     int64_t nq, int k, int d, int Kp, int l2, int64_t id_base,
                                     const Tidx* __restrict__ I, float* __restrict__ D) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (wid >= nq * k) return;
    const int64_t q = wid / k;
    const int64_t id = static_cast<int64_t>(I[wid]);
    if (id < 0) return;
    const __nv_bfloat16* a = xq + q * 3 * Kp;
    const __nv_bfloat16* b = xb + (id - id_base) * 3 * Kp;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float av = __bfloat162float(a[c]) + __bfloat162float(a[Kp + c]) + __bfloat162float(a[2 * Kp + c]);
        const float bv = __bfloat162float(b[c]) + __bfloat162float(b[Kp + c]) + __bfloat162float(b[2 * Kp + c]);
        if (l2) {
            const float df = av - bv;
            acc = fmaf(df, df, acc);
        } else {
            acc = fmaf(av, bv, acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) D[wid] = acc;
}

}  // namespace cvdb
