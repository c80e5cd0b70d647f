// C ABI of the engine (include/cvdb_b200.h): index storage in HBM, query
// packing, kernel selection and launch, k-way merge.  Host logic only - every
// arithmetic step runs in the CUDA kernels included below.
#include "../../include/cvdb_b200.h"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#ifdef __linux__
#include <sched.h>
#endif
#include <vector>

#include "aux_kernels.cuh"
#include "ivf_scan.cuh"
#include "launchers.h"

namespace {

using namespace cvdb;

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(e_ == cudaErrorMemoryAllocation ? CVDB_ENOMEM : CVDB_ECUDA, "%s failed: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                   \
    } while (0)

#define TRY(expr)                  \
    do {                           \
        int r_ = (expr);           \
        if (r_ != CVDB_OK) return r_; \
    } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return CVDB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) return fail(CVDB_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        cap = want;
        return CVDB_OK;
    }
    // grow to `bytes`, preserving the first `keep` bytes
    int grow_keep(size_t bytes, size_t keep, cudaStream_t st) {
        if (bytes <= cap) return CVDB_OK;
        void* np_ = nullptr;
        const size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&np_, want);
        if (e != cudaSuccess) return fail(CVDB_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        if (p && keep) {
            e = cudaMemcpyAsync(np_, p, keep, cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) {
                cudaFree(np_);
                return fail(CVDB_ECUDA, "device copy failed: %s", cudaGetErrorString(e));
            }
        }
        if (p) cudaFree(p);
        p = np_;
        cap = want;
        return CVDB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() { return static_cast<T*>(p); }
};

// ---------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// 2-D bf16 row-major matrix [rows][row_elems]; box = {64 columns, box_rows}, 128-byte swizzle.
int make_tmap_2d(CUtensorMap* m, const void* base, int64_t rows, int row_elems, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(CVDB_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(row_elems), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstride[1] = {static_cast<cuuint64_t>(row_elems) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CVDB_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    return CVDB_OK;
}

// ------------------------------------------------------------------- index
struct Index {
    int d = 0, metric = 0, storage = 0, device = 0;
    int Kd = 0;         // data columns per plane: d (+3 norm columns for L2)
    int Kp = 0;         // stored columns per plane (Kd padded)
    int planes = 1;     // 1 (bf16) or 3 (exact split)
    int row_elems = 0;  // planes * Kp
    int64_t ntotal = 0, capacity = 0;
    __nv_bfloat16* x = nullptr;
    int num_sms = 0;

    // inverted-list (IVF) layout: rows stored list-major after cvdb_index_group_rows
    bool grouped = false;
    int nlist = 0;
    int64_t row_ids_n = 0;  // leading stored rows whose row_ids entry is valid (later rows: id == position)
    DevBuf row_ids, list_off, ivf_cnt, ivf_pair_off, ivf_item_off, ivf_cursor, ivf_scal, ivf_items, ivf_pair_query,
        ivf_pair_dst, ivf_qg, ivf_probes;
    DevBuf gthr, waves, stage_in, q_pack, q_norm, cand, part, out_d, out_i, ids_a, ids_b, groups;
    // symmetric self-join state (cvdb_selfjoin_*): per database row a threshold, a counter, the survivors' count
    // and a buffer of kColCap keys for the column direction; sj_k > 0 while a join is open
    DevBuf col_thr, col_cnt, col_base, col_buf, col_dirty, col_scal, col_log;
    uint32_t col_log_cap = 0;  // records
    int sj_k = 0;
    DevBuf bad_rows;  // one uint64: rows (added or queried) whose squared norm was not finite
    bool has_groups = false;
    // Rows re-stored list-major keep their caller-visible ids in row_ids: every id leaving the library is
    // translated through it, and everything indexed by STORED POSITION (groups, self exclusion, export,
    // truncate) is refused while the index is in this state.
    bool permuted() const { return row_ids_n > 0; }

    // Tensor maps of the last launch, keyed by what they describe (encoding one costs microseconds of host time
    // on the small-batch path).
    struct TmapSlot {
        const void* base = nullptr;
        int64_t rows = -1;
        int row_elems = 0, box_rows = 0;
        CUtensorMap map;
    };
    TmapSlot tm_q, tm_x, tm_q_ivf, tm_x_ivf, tm_x32_ivf;

    // The scratch buffers above are per handle: a call on another stream than the previous one first waits for
    // the previous call's work (event recorded at the end of every call).
    cudaEvent_t done_ev = nullptr;
    cudaStream_t last_stream = nullptr;
    bool has_last = false;

    // host->device ingest: two pinned staging buffers + two device staging buffers, so the copy of chunk
    // i+1 overlaps the packing kernel of chunk i
    void* pin[2] = {nullptr, nullptr};
    size_t pin_cap = 0;
    cudaEvent_t pin_free[2] = {nullptr, nullptr};  // the H2D copy out of pin[b] has finished
    DevBuf stage2[2];
    cudaEvent_t stage_free[2] = {nullptr, nullptr};  // the packing kernel reading stage2[b] has finished
    cudaStream_t copy_stream = nullptr;

    // ring of event pairs bracketing profiled GEMM+top-k launches (opts.profile)
    static constexpr int kProfSlots = 64;
    cudaEvent_t ev0[kProfSlots] = {}, ev1[kProfSlots] = {};
    int prof_head = 0;   // next slot to use
    int prof_count = 0;  // recorded and not yet read (<= kProfSlots)
    double last_flops = 0, last_bytes = 0;
    int last_slices = 0, last_grid = 0, last_variant = 0;
};

struct cvdb_guard {
    int prev = -1;
    explicit cvdb_guard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~cvdb_guard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Orders a call on `st` after the previous call on this handle when that one ran on another stream.
struct StreamOrder {
    Index* ix;
    cudaStream_t st;
    StreamOrder(Index* ix_, cudaStream_t st_) : ix(ix_), st(st_) {
        if (ix->has_last && ix->last_stream != st && ix->done_ev) cudaStreamWaitEvent(st, ix->done_ev, 0);
    }
    ~StreamOrder() {
        if (!ix->done_ev && cudaEventCreateWithFlags(&ix->done_ev, cudaEventDisableTiming) != cudaSuccess) {
            ix->done_ev = nullptr;
            return;
        }
        if (cudaEventRecord(ix->done_ev, st) == cudaSuccess) {
            ix->last_stream = st;
            ix->has_last = true;
        }
    }
};

// The handle-less entry points (k-means update, merges, triplets) take raw device pointers: run them on the
// device that owns the data, whatever the caller's current device is.
int device_of(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}
struct ptr_guard {
    int prev = -1;
    explicit ptr_guard(const void* p) {
        const int dev = device_of(p);
        if (dev < 0) return;
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~ptr_guard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int grow(Index* ix, int64_t need, cudaStream_t st) {
    if (need <= ix->capacity) return CVDB_OK;
    int64_t cap = std::max<int64_t>(need, ix->capacity + ix->capacity / 2);
    cap = std::max<int64_t>(cap, 1024);
    __nv_bfloat16* nx = nullptr;
    size_t bytes = static_cast<size_t>(cap) * ix->row_elems * 2;
    cudaError_t e = cudaMalloc(&nx, bytes);
    if (e != cudaSuccess && cap > need) {
        cap = need;
        bytes = static_cast<size_t>(cap) * ix->row_elems * 2;
        e = cudaMalloc(&nx, bytes);
    }
    if (e != cudaSuccess) return fail(CVDB_ENOMEM, "cudaMalloc(%zu bytes) for %lld rows failed: %s", bytes,
                                      static_cast<long long>(cap), cudaGetErrorString(e));
    if (ix->ntotal > 0) {
        CU_TRY(cudaMemcpyAsync(nx, ix->x, static_cast<size_t>(ix->ntotal) * ix->row_elems * 2, cudaMemcpyDeviceToDevice, st));
        CU_TRY(cudaStreamSynchronize(st));
    }
    if (ix->x) cudaFree(ix->x);
    ix->x = nx;
    ix->capacity = cap;
    return CVDB_OK;
}

template <typename Tin>
void launch_pack(const Tin* in, int64_t n, int64_t n_pad, int d, __nv_bfloat16* out, const Index* ix, int is_query,
                 float* norms, cudaStream_t st) {
    const int threads = 256;
    const int64_t blocks = std::min<int64_t>(ceil_div(n_pad, threads / 32), 148 * 32);
    pack_rows_kernel<Tin><<<static_cast<unsigned>(blocks), threads, 0, st>>>(in, n, n_pad, d, d, out, ix->Kp, ix->planes,
                                                                                ix->metric == CVDB_METRIC_L2, is_query, norms,
                                                                                static_cast<unsigned long long*>(ix->bad_rows.p));
    ++g_launches;
}

// n rows of `in` -> packed rows of `out`; rows n .. n_pad-1 of `out` become zeros
int pack_dispatch(const void* in, int dtype, int64_t n, __nv_bfloat16* out, const Index* ix, int is_query, float* norms,
                  cudaStream_t st, int64_t n_pad = 0) {
    n_pad = std::max(n_pad, n);
    if (n_pad == 0) return CVDB_OK;
    if (dtype == CVDB_DTYPE_F32)
        launch_pack<float>(static_cast<const float*>(in), n, n_pad, ix->d, out, ix, is_query, norms, st);
    else if (dtype == CVDB_DTYPE_F16)
        launch_pack<__half>(static_cast<const __half*>(in), n, n_pad, ix->d, out, ix, is_query, norms, st);
    else
        launch_pack<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(in), n, n_pad, ix->d, out, ix, is_query, norms, st);
    CU_TRY(cudaGetLastError());
    return CVDB_OK;
}

int get_tmap(Index::TmapSlot& s, CUtensorMap* out, const void* base, int64_t rows, int row_elems, int box_rows) {
    if (s.base != base || s.rows != rows || s.row_elems != row_elems || s.box_rows != box_rows) {
        TRY(make_tmap_2d(&s.map, base, rows, row_elems, box_rows));
        s.base = base;
        s.rows = rows;
        s.row_elems = row_elems;
        s.box_rows = box_rows;
    }
    *out = s.map;
    return CVDB_OK;
}

// final merge of `n_lists` sorted key lists per query -> (D, I) or rewritten keys (see merge_partials_kernel)
int launch_merge(const uint64_t* part, int64_t nq, int n_lists, int k_in, int k, int l2, const float* qnorm,
                 int64_t id_base, int64_t q_stride, int64_t l_stride, const int32_t* row_ids, int64_t row_ids_n, float* D,
                 int64_t* I64, int32_t* I32, uint64_t* keys_out, cudaStream_t st) {
    const unsigned blocks = static_cast<unsigned>(ceil_div(nq, 8));
    const size_t smem = 8 * static_cast<size_t>(n_lists) * sizeof(uint16_t);
    if (keys_out)
        merge_partials_kernel<int64_t, true><<<blocks, 256, smem, st>>>(part, nq, n_lists, k_in, k, l2, qnorm, id_base, q_stride,
                                                                         l_stride, row_ids, row_ids_n, nullptr, nullptr, keys_out);
    else if (I64)
        merge_partials_kernel<int64_t, false><<<blocks, 256, smem, st>>>(part, nq, n_lists, k_in, k, l2, qnorm, id_base, q_stride,
                                                                          l_stride, row_ids, row_ids_n, D, I64, nullptr);
    else
        merge_partials_kernel<int32_t, false><<<blocks, 256, smem, st>>>(part, nq, n_lists, k_in, k, l2, qnorm, id_base, q_stride,
                                                                          l_stride, row_ids, row_ids_n, D, I32, nullptr);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return CVDB_OK;
}

// ------------------------------------------------------------ host ingest
constexpr size_t kPinChunk = size_t(64) << 20;  // bytes per pinned staging buffer
constexpr size_t kSmallAdd = size_t(8) << 20;   // below this an add() is one plain copy

int host_threads() {
    int n = 0;
#ifdef __linux__
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
#endif
    if (n <= 0) n = static_cast<int>(std::thread::hardware_concurrency());
    return std::max(1, n);
}

// memcpy split over a few host threads: one core moves ~10 GB/s, PCIe Gen5 x16 takes ~50
void parallel_copy(char* dst, const char* src, size_t bytes) {
    static const int kThreads = std::max(1, std::min(8, host_threads() / 2));
    const int nt = bytes < (size_t(4) << 20) ? 1 : kThreads;
    if (nt == 1) {
        memcpy(dst, src, bytes);
        return;
    }
    const size_t per = ((bytes + nt - 1) / nt + 4095) & ~size_t(4095);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) {
        const size_t off = static_cast<size_t>(t) * per;
        if (off >= bytes) break;
        th.emplace_back([=] { memcpy(dst + off, src + off, std::min(per, bytes - off)); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (auto& t : th) t.join();
}

// n host rows -> packed index rows at `dst`.  Large inputs run as a three-stage pipeline over 64 MB chunks
// (buffer b = chunk & 1):
//   host threads   user rows -> pin[b]             once the H2D copy of chunk c-2 has left pin[b]
//   copy stream    pin[b] -> stage2[b]             once the packing kernel of chunk c-2 is done with stage2[b]
//   caller stream  stage2[b] -> index rows (pack)  once the copy has landed
// so the PCIe copy of a chunk overlaps the packing of the previous one and the host memcpy of the next.
// Memory the caller already pinned (cudaHostAlloc / cudaHostRegister) is copied from directly.
// Returns after the last row has left host memory (the caller may reuse x) and has been packed.
int ingest_host_rows(Index* ix, const char* x, int64_t n, int dtype, size_t esz, __nv_bfloat16* dst, cudaStream_t st) {
    const size_t row_b = static_cast<size_t>(ix->d) * esz;
    const size_t total = static_cast<size_t>(n) * row_b;
    if (total <= kSmallAdd) {
        TRY(ix->stage_in.ensure(total));
        CU_TRY(cudaMemcpyAsync(ix->stage_in.p, x, total, cudaMemcpyHostToDevice, st));
        TRY(pack_dispatch(ix->stage_in.p, dtype, n, dst, ix, 0, nullptr, st));
        CU_TRY(cudaStreamSynchronize(st));
        return CVDB_OK;
    }
    cudaPointerAttributes attr{};
    const bool src_pinned = cudaPointerGetAttributes(&attr, x) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();  // unregistered host memory is reported as an error by older runtimes
    const int64_t chunk_rows = std::max<int64_t>(1, static_cast<int64_t>(kPinChunk / row_b));
    const size_t chunk_b = static_cast<size_t>(chunk_rows) * row_b;
    if (!ix->copy_stream) CU_TRY(cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        if (!ix->pin_free[b]) CU_TRY(cudaEventCreateWithFlags(&ix->pin_free[b], cudaEventDisableTiming));
        if (!ix->stage_free[b]) CU_TRY(cudaEventCreateWithFlags(&ix->stage_free[b], cudaEventDisableTiming));
        TRY(ix->stage2[b].ensure(chunk_b));
    }
    if (!src_pinned && ix->pin_cap < chunk_b) {
        for (int b = 0; b < 2; ++b) {
            if (ix->pin[b]) cudaFreeHost(ix->pin[b]);
            ix->pin[b] = nullptr;
        }
        ix->pin_cap = 0;
        for (int b = 0; b < 2; ++b) CU_TRY(cudaHostAlloc(&ix->pin[b], chunk_b, cudaHostAllocDefault));
        ix->pin_cap = chunk_b;
    }
    int64_t c = 0;
    for (int64_t r0 = 0; r0 < n; r0 += chunk_rows, ++c) {
        const int b = static_cast<int>(c & 1);
        const int64_t m = std::min(chunk_rows, n - r0);
        const size_t bytes = static_cast<size_t>(m) * row_b;
        const char* src = x + static_cast<size_t>(r0) * row_b;
        if (!src_pinned) {
            if (c >= 2) CU_TRY(cudaEventSynchronize(ix->pin_free[b]));
            parallel_copy(static_cast<char*>(ix->pin[b]), src, bytes);
            src = static_cast<const char*>(ix->pin[b]);
        }
        if (c >= 2) CU_TRY(cudaStreamWaitEvent(ix->copy_stream, ix->stage_free[b], 0));
        CU_TRY(cudaMemcpyAsync(ix->stage2[b].p, src, bytes, cudaMemcpyHostToDevice, ix->copy_stream));
        CU_TRY(cudaEventRecord(ix->pin_free[b], ix->copy_stream));
        CU_TRY(cudaStreamWaitEvent(st, ix->pin_free[b], 0));
        TRY(pack_dispatch(ix->stage2[b].p, dtype, m, dst + r0 * ix->row_elems, ix, 0, nullptr, st));
        CU_TRY(cudaEventRecord(ix->stage_free[b], st));
    }
    CU_TRY(cudaStreamSynchronize(ix->copy_stream));
    CU_TRY(cudaStreamSynchronize(st));
    return CVDB_OK;
}

// ------------------------------------------------------- kernel selection
constexpr int kBlockN = 256;

// Candidate-buffer size (32*E slots per query): about twice k, so that a compaction (a bitonic sort of the whole
// buffer) is needed only once per ~k accepted rows.  Up to 512 slots the sort runs in registers; larger buffers
// (k > 248) are sorted in place in L2-resident memory.
int pick_E(int k) {
    if (const char* env = getenv("CVDB_E_MULT")) {  // experiments: size the buffer as if k were larger
        const int m = atoi(env);
        if (m > 1 && k > 1) k = std::min(k * m, 248);
    }
    if (k == 1) return 0;
    if (k <= 12) return 1;
    if (k <= 28) return 2;
    if (k <= 60) return 4;
    if (k <= 124) return 8;
    if (k <= 248) return 16;
    if (k <= 504) return 32;
    if (k <= 1016) return 64;
    return 128;
}

// Candidates a buffer of C slots may hold before it is compacted.  The scan makes room once per group of eight
// scores (32-slot buffers, k <= 12) or once per chunk of 32 scores (larger buffers), so that many slots must stay
// free.  CVDB_ROOM_EXTRA = x lowers the trigger to k + x (experiments only).
int compaction_trigger(int k, int C) {
    // measured (mining shape, k = 50): compacting earlier than necessary is slower (500 / 524 / 550 ms for
    // C-8 / k+25 / k+12) -- a compaction stalls the warp on L2 round trips -- so fill the buffer.
    const int full = C == 32 ? C - 8 : C - 32;
    int room = full;
    if (const char* env = getenv("CVDB_ROOM_EXTRA")) room = std::min(full, k + std::max(1, atoi(env)));
    return room;
}

// Split the database into slices so that (query tiles x slices) fills the grid
// in whole waves.  Cost model: waves * (tiles per slice + fixed per-item overhead).
void choose_slices(int q_tiles, int n_tiles, int grid, int64_t max_slices, int& n_slices, int& tps, int ovh = 4) {
    // cost of a plan in tile times: waves * (tiles per slice + per-item overhead), plus a penalty for LONG slices.  The
    // CTAs of a wave stream the same slice and meet again only at the item boundary; over thousands of tiles they
    // drift further apart than the L2 bridges and re-read the slice from HBM (which costs clock under the power cap).
    // Measured on the mining chunk (65 536 anchors x 10M rows, k = 50, tools/mining_slices_probe.py): 6010 tiles per
    // slice 755 ms, 3907: 680, 3005: 665, 2112: 659, 1954: 662 -- i.e. +0.9 / +3 / +14 % at 3000 / 3900 / 6000 tiles.
    double best = 1e300;
    n_slices = 1;
    tps = n_tiles;
    const int64_t s_max = std::min<int64_t>(std::min<int64_t>(n_tiles, 8LL * grid), std::max<int64_t>(max_slices, 1));
    for (int64_t s = 1; s <= s_max; ++s) {
        const int64_t t = ceil_div(n_tiles, s);
        const int64_t se = ceil_div(n_tiles, t);
        const int64_t items = se * q_tiles;
        const int64_t waves = ceil_div(items, grid);
        const double over = t > 2000 ? (static_cast<double>(t) - 2000.0) / 1000.0 : 0.0;
        const double cost = static_cast<double>(waves * (t + ovh)) * (1.0 + 0.009 * over * over);
        if (cost < best) {
            best = cost;
            n_slices = static_cast<int>(se);
            tps = static_cast<int>(t);
        }
    }
}

// translate a launcher's cudaError_t
#define LAUNCH(expr)                                                                                        \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        ++g_launches;                                                                                       \
        if (e_ != cudaSuccess) return fail(CVDB_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));     \
    } while (0)

// Core: search `nq` packed-on-the-fly queries against the whole index.
// Outputs are device pointers: D [nq][k] f32 and either I64 or I32 [nq][k].
// What the symmetric self-join adds to a search: the queries ARE packed rows of the index (no packing pass),
// only database rows from row_begin on are scanned, and every score is also offered to its database row.
struct SearchExtra {
    const __nv_bfloat16* q_packed = nullptr;  // packed query rows (nq of them used)
    int64_t q_rows_avail = 0;                 // rows readable from q_packed on
    int64_t row_begin = 0;                    // first database row scanned (multiple of 128)
    int64_t row_end = 0;                      // database rows scanned end here (0: ntotal)
    bool col = false;                         // column direction on (Index::col_*)
    int64_t col_row_min = 0;                  // database rows below this do not collect
    const int32_t* q_ids = nullptr;           // [nq] ids the column lists record for the queries (null: self_ids)
};

// `keys_out` (optional, instead of D/I): the merged top-k as keys carrying the caller's ids (shard exchange).
// `q_norm`: where the squared norms of the packed queries go ([nq] f32, device).
int search_device(Index* ix, const void* q_dev, int64_t nq, int dtype, int k, float* D, int64_t* I64, int32_t* I32,
                  const int32_t* self_ids, const int32_t* group_q, const cvdb_search_opts* opts, cudaStream_t st,
                  uint64_t* keys_out = nullptr, float* q_norm = nullptr, const SearchExtra* ex = nullptr) {
    const int E = pick_E(k);
    const int C = 32 * (E ? E : 1);
    const int l2 = ix->metric == CVDB_METRIC_L2;
    const int64_t id_base = opts ? opts->id_base : 0;

    // the packed query matrix is padded with zero rows to a whole 256-query tile, so the A-operand
    // TMA boxes never run out of bounds (partially out-of-bounds boxes load measurably slower)
    const int64_t nq_pad = ceil_div(nq, 256) * 256;
    if (!q_norm) {
        TRY(ix->q_norm.ensure(static_cast<size_t>(nq) * 4));
        q_norm = ix->q_norm.as<float>();
    }
    const __nv_bfloat16* q_rows = nullptr;  // the packed queries
    int64_t q_rows_n = nq_pad;              // rows the query tensor map covers
    if (ex && ex->q_packed) {
        q_rows = ex->q_packed;
        q_rows_n = std::min(nq_pad, ex->q_rows_avail);
    } else {
        TRY(ix->q_pack.ensure(static_cast<size_t>(nq_pad) * ix->row_elems * 2));
        TRY(pack_dispatch(q_dev, dtype, nq, ix->q_pack.as<__nv_bfloat16>(), ix, 1, q_norm, st, nq_pad));
        q_rows = ix->q_pack.as<__nv_bfloat16>();
    }
    const int32_t* row_ids = ix->permuted() ? ix->row_ids.as<int32_t>() : nullptr;

    if (ix->ntotal == 0) {
        // nothing to search: all padding
        TRY(ix->part.ensure(static_cast<size_t>(nq) * k * 8));
        CU_TRY(cudaMemsetAsync(ix->part.p, 0, static_cast<size_t>(nq) * k * 8, st));
        return launch_merge(ix->part.as<uint64_t>(), nq, 1, k, k, l2, q_norm, id_base, k, k, nullptr, 0, D, I64, I32,
                            keys_out, st);
    }

    GemmTopkParams p{};
    const int64_t scan_end = (ex && ex->row_end > 0) ? std::min<int64_t>(ex->row_end, ix->ntotal) : ix->ntotal;
    p.nq = static_cast<int>(nq);
    p.n_rows = static_cast<int>(scan_end);
    p.k = k;
    p.room = compaction_trigger(k, C);
    p.dbg = opts ? opts->debug_flags : 0;
    p.nkb = static_cast<int>(ceil_div(ix->Kp, 64));
    p.k16 = (opts && (opts->debug_flags & 16)) ? 4 * p.nkb : static_cast<int>(ceil_div(ix->Kd, 16));
    p.plane_cols = ix->Kp;
    if (ix->planes == 1) {
        p.n_combo = 1;
        p.a_planes = p.b_planes = 0;
    } else {
        // smallest terms first: (lo,hi) (hi,lo) (mid,mid) (mid,hi) (hi,mid) (hi,hi)
        p.n_combo = 6;
        p.a_planes = 0x001102u;  // nibble c = A plane of combo c
        p.b_planes = 0x010120u;  // nibble c = B plane of combo c
    }
    // Kernel variant (measured on B200, see DESIGN.md "variant selection"):
    //   1 = single-CTA kernel streaming both operands (any K, any storage): small (HBM-bound)
    //       batches, exact-split storage, K > 768,
    //   2 = CTA pair with the queries resident on chip (bf16 storage, padded K <= 768; TMEM for the
    //       first 512 of K, shared-memory tail beyond; 128-column accumulators): large batches,
    //   3 = CTA pair streaming both operands (M = 256, N = 256; any K, any storage): short K with a real top-k, and
    //       every batch of more than 128 queries the resident-query kernel cannot take (exact storage, K > 832),
    //   4 = CTA pair with all of K <= 768 in TMEM and 64-column accumulators (comparison only).
    const bool ts2_ok = ix->planes == 1 && p.nkb <= 13;  // padded K <= 832 (768 + the L2 norm columns)
    int variant = 1;
    if (ts2_ok && nq > 128) {
        // more than one 128-query tile: the CTA-pair kernels take 256 queries per pass over the database (129..256
        // queries on 12.5M x 768: 4.1-4.8 ms against 5.5-6.6 ms for two passes of the single-CTA kernel --
        // tools/ridge_probe.py).  Resident queries win for K >= 448 and for k = 1; for short K with a real top-k the
        // 256 x 256 streaming pair kernel is ahead (K = 384, k = 10: 1325 vs 1236 TFLOP/s), and for K <= 128 the
        // single-CTA kernel (841 vs 804 / 707) -- tools/variant_sweep.py
        if (p.nkb <= 2) variant = 1;
        else if (p.nkb <= 6 && k > 1) variant = 3;
        else variant = 2;
    } else if (nq > 128) {
        // exact (three-plane) storage or padded K > 832: the queries cannot stay resident, but the streaming CTA pair
        // still fetches every database tile once per 256 queries.  tools/exact_variant_probe.py, variant 1 -> 3:
        // configs[0] (100k x 384 exact, 10k queries) 3.68 -> 3.22 ms; 2M x 768 exact 152 -> 137 ms (10k queries),
        // 4.4 -> 3.5 ms (256 queries); bf16 K = 1024 / 1536: 1180 -> 1360 TFLOP/s; identical results
        variant = 3;
    }
    if (opts && opts->force_variant == 1) variant = 1;
    if (opts && opts->force_variant == 3) variant = 3;
    if (opts && opts->force_variant == 2) {
        if (!ts2_ok) return fail(CVDB_EINVAL, "variant 2 needs bf16 storage and padded d <= 832");
        variant = 2;
    }
    if (opts && opts->force_variant == 4) {
        if (!(ts2_ok && p.nkb <= 12)) return fail(CVDB_EINVAL, "variant 4 needs bf16 storage and padded d <= 768");
        variant = 4;
    }
    if (ex && (ex->col || ex->row_begin > 0)) {
        if (!(ts2_ok && p.nkb <= 12)) return fail(CVDB_EINVAL, "the symmetric self-join needs bf16 storage and padded d <= 768");
        variant = 2;
    }
    const bool resident = variant == 2 || variant == 4;
    const int block_n = !resident ? kBlockN : (variant == 4 ? 64 : 128);
    const int q_tile = variant == 1 ? 128 : 256;
    const int sms = ix->num_sms;
    const int workers = variant == 1 ? sms : sms / 2;  // CTAs or CTA pairs
    p.q_tiles = static_cast<int>(ceil_div(nq, q_tile));
    p.n_tiles = static_cast<int>(ceil_div(scan_end, block_n));  // END tile of the scan
    p.tile0 = ex ? static_cast<int>(ex->row_begin / block_n) : 0;
    const int scan_tiles = p.n_tiles - p.tile0;
    // keep the per-slice result scratch under ~1 GiB
    const int64_t max_slices = std::max<int64_t>(1, (int64_t(1) << 30) / std::max<int64_t>(1, nq * k * 8));
    if (opts && opts->force_slices > 0) {
        const int64_t s = std::min<int64_t>(opts->force_slices, scan_tiles);
        p.tiles_per_slice = static_cast<int>(ceil_div(scan_tiles, s));
        p.n_slices = static_cast<int>(ceil_div(scan_tiles, p.tiles_per_slice));
    } else {
        // per-item overhead in tile times: the resident-query kernel reloads its queries into TMEM and drains
        // the pipeline at every item boundary (~10 us), the streaming kernels only flush their results
        int ovh = (variant == 2 || variant == 4) ? 8 : 4;
        if (const char* env = getenv("CVDB_PLAN_OVH")) ovh = atoi(env);  // tuning experiments
        choose_slices(p.q_tiles, scan_tiles, workers, max_slices, p.n_slices, p.tiles_per_slice, ovh);
    }
    const int64_t n_items = static_cast<int64_t>(p.q_tiles) * p.n_slices;
    const int grid = static_cast<int>(std::min<int64_t>(workers, n_items)) * (variant == 1 ? 1 : 2);
    p.self_ids = self_ids;
    p.group_q = group_q;
    p.group_db = (group_q && ix->has_groups) ? ix->groups.as<int32_t>() : nullptr;
    if (E > 0) TRY(ix->cand.ensure(static_cast<size_t>(grid) * 128 * C * 8));
    TRY(ix->part.ensure(static_cast<size_t>(nq) * p.n_slices * k * 8));
    p.cand = ix->cand.as<uint64_t>();
    p.part = ix->part.as<uint64_t>();
    {
        // shared thresholds [nq], wave counters [n_waves] and the per-item "flushed" counters [q_tiles * n_slices] live
        // in one buffer: one memset per launch
        const int64_t n_waves = ceil_div(n_items, std::min<int64_t>(workers, n_items));
        // + the pooled per-slice order statistics (gemm_topk.cuh, "pooled thresholds"): worth it from 3 slices on and
        // when an item is long enough for the 16-step publish chain at its end to vanish (debug flag 256: off)
        // (a launch of a single wave -- small HBM-bound batches -- has nobody to publish to: all items start together)
        const bool pool = E > 0 && E <= 16 && p.n_slices >= 3 &&
                          ((p.tiles_per_slice >= 128 && n_waves > 1) || (opts && (opts->debug_flags & 512))) &&
                          !(opts && (opts->debug_flags & (4 | 128 | 256)));
        const size_t pool_off = (static_cast<size_t>(nq + n_waves + n_items) * 4 + 15) & ~size_t(15);  // uint4 loads
        const size_t gthr_bytes = pool_off + (pool ? static_cast<size_t>(nq) * kPoolSlots * 4 : 0);
        TRY(ix->gthr.ensure(gthr_bytes));
        CU_TRY(cudaMemsetAsync(ix->gthr.p, 0, gthr_bytes, st));
        p.gpool = pool ? reinterpret_cast<uint32_t*>(static_cast<char*>(ix->gthr.p) + pool_off) : nullptr;
        p.gthr = (opts && (opts->debug_flags & 4)) ? nullptr : ix->gthr.as<uint32_t>();
        p.wave_cnt = (opts && (opts->debug_flags & 8)) ? nullptr : ix->gthr.as<uint32_t>() + nq;
        // Optional (debug flag 128): a slice starts from the result of the latest finished earlier slice of its query
        // tile (gemm_topk.cuh, item_begin).  Exact either way.  Same-box A/B (profiles/r2_inheritance_ab.jsonl): +1 % at
        // k = 50 on the headline shape, but -4 % on the mining chunk and -21 % at k = 200, so it is off by default.
        const bool inherit = E > 0 && E <= 16 && p.n_slices > 1 && p.tiles_per_slice >= 32 && opts && (opts->debug_flags & 128);
        p.done = inherit ? ix->gthr.as<uint32_t>() + nq + n_waves : nullptr;
        p.done_full = variant == 1 ? 4 : 8;
    }

    // a single partial query tile on the single-CTA kernel (small, HBM-bound batches): give every epilogue warp a
    // quarter of the queries instead of filling the TMEM lanes from 0 up, where 32 queries are one warp's work
    p.a_quarter = (variant == 1 && nq < 128 && !(opts && (opts->debug_flags & 64))) ? static_cast<int>(ceil_div(nq, 4)) : 0;
    CUtensorMap tq, tx;
    TRY(get_tmap(ix->tm_q, &tq, q_rows, q_rows_n, ix->row_elems, p.a_quarter > 0 ? 32 : 128));
    TRY(get_tmap(ix->tm_x, &tx, ix->x, ix->ntotal, ix->row_elems, variant == 1 ? kBlockN : block_n / 2));

    const bool prof = opts && opts->profile;
    const int slot = ix->prof_head;
    if (prof) {
        if (!ix->ev0[slot]) {
            CU_TRY(cudaEventCreate(&ix->ev0[slot]));
            CU_TRY(cudaEventCreate(&ix->ev1[slot]));
        }
        CU_TRY(cudaEventRecord(ix->ev0[slot], st));
    }
    if (variant == 4) {
        LAUNCH(launch_ts2(3, E, tx, tq, q_rows, ix->row_elems, p, grid, st));
    } else if (variant == 2) {
        const int cfg = p.nkb <= 8 ? 0 : (p.nkb <= 12 ? 1 : 2);  // all of K in TMEM / + 4-block tail / + 5-block tail
        if (ex && ex->col) {
            p.col_thr = ix->col_thr.as<float>();
            p.col_log = ix->col_log.as<uint4>();
            p.col_log_cnt = ix->col_scal.as<uint32_t>() + 4;
            p.col_log_cap = ix->col_log_cap;
            p.col_row_min = static_cast<int>(ex->col_row_min);
            p.q_ids = ex->q_ids;
            LAUNCH(launch_ts2_col(cfg, E, tx, tq, q_rows, ix->row_elems, p, grid, st));
        } else {
            LAUNCH(launch_ts2(cfg, E, tx, tq, q_rows, ix->row_elems, p, grid, st));
        }
    } else if (variant == 3) {
        LAUNCH(launch_ss2(E, tq, tx, p, grid, st));
    } else {
        LAUNCH(launch_ss1(E, tq, tx, p, grid, st));
    }
    if (prof) {
        CU_TRY(cudaEventRecord(ix->ev1[slot], st));
        ix->prof_head = (slot + 1) % Index::kProfSlots;
        ix->prof_count = std::min(ix->prof_count + 1, Index::kProfSlots);
    }
    const double rows_scanned = double(scan_end) - (ex ? double(ex->row_begin) : 0.0);
    ix->last_flops = 2.0 * double(nq) * rows_scanned * double(ix->d);
    ix->last_bytes = rows_scanned * double(ix->row_elems) * 2.0;
    ix->last_slices = p.n_slices;
    ix->last_grid = grid;
    ix->last_variant = variant;

    TRY(launch_merge(p.part, nq, p.n_slices, k, k, l2, q_norm, id_base, static_cast<int64_t>(p.n_slices) * k, k, row_ids,
                     ix->row_ids_n, D, I64, I32, keys_out, st));

    if (ix->planes == 3 && D != nullptr && !keys_out) {
        const unsigned rb = static_cast<unsigned>(ceil_div(nq * k, 8));
        if (I64)
            rescore_exact_kernel<int64_t><<<rb, 256, 0, st>>>(ix->q_pack.as<__nv_bfloat16>(), ix->x, nq, k, ix->d, ix->Kp,
                                                              l2, id_base, I64, D);
        else
            rescore_exact_kernel<int32_t><<<rb, 256, 0, st>>>(ix->q_pack.as<__nv_bfloat16>(), ix->x, nq, k, ix->d, ix->Kp,
                                                              l2, id_base, I32, D);
        ++g_launches;
        CU_TRY(cudaGetLastError());
    }
    return CVDB_OK;
}

static_assert(sizeof(IvfItem) == sizeof(GroupItem), "item layouts must agree");
constexpr int kGroupedBlockN = 128;

// Search `nq` queries, each restricted to the `nprobe` inverted lists named in probes[nq][nprobe]
// (device pointers; D/I device outputs).
int search_lists_device(Index* ix, const void* q_dev, int64_t nq, int dtype, int k, const int32_t* probes, int nprobe,
                        float* D, int64_t* I, cudaStream_t st) {
    // Every (query, list) pair starts its scan cold and a list is only a few hundred rows, so about
    // k * (1 + ln(rows / k)) rows pass the filter per pair.  Larger buffers (fewer sorts) were measured and do not
    // help (4M x 768, nlist 8192, nprobe 8, 10k queries: 3.58 / 3.60 / 3.95 ms per batch with 1x / 2x / 4x the flat
    // buffer size): the cost is the candidate walk itself, serialised in the one warp that owns the few valid
    // query rows of an item.  CVDB_IVF_EMULT = 2, 4 re-runs that experiment.
    int E = pick_E(k);
    if (const char* env = getenv("CVDB_IVF_EMULT")) {
        const int mult = std::max(1, atoi(env));
        if (E >= 1 && E < 16) E = std::min(16, E * mult);
    }
    const int C = 32 * (E ? E : 1);
    const int l2 = ix->metric == CVDB_METRIC_L2;
    const int64_t n_pairs = nq * nprobe;
    const int nlist = ix->nlist;
    const int row_vec16 = ix->row_elems * 2 / 16;

    TRY(ix->q_pack.ensure(static_cast<size_t>(nq) * ix->row_elems * 2));
    TRY(ix->q_norm.ensure(static_cast<size_t>(nq) * 4));
    TRY(pack_dispatch(q_dev, dtype, nq, ix->q_pack.as<__nv_bfloat16>(), ix, 1, ix->q_norm.as<float>(), st));

    // Which scan kernel: with few pairs per list (the usual IVF regime: many lists, few probes) the transposed kernel
    // (list rows on the M side, <= 16 queries per item) streams every list once at HBM rate; with many pairs per
    // list the grouped kernel (128 queries per item) re-reads a list fewer times.  CVDB_IVF_KERNEL=0/1 forces one.
    const int nkb_ = static_cast<int>(ceil_div(ix->Kp, 64));
    // Measured (10M x 768, nlist 16 384, 10k queries, k = 10): 4.9 pairs per list (nprobe 8) 3.43 ms against 4.01 ms,
    // 19.5 pairs per list (nprobe 32: two 16-query items for most lists) 9.1 ms against 7.9 ms.
    bool transposed = k <= 128 && nkb_ <= kIvfMaxKb && n_pairs <= 10 * static_cast<int64_t>(nlist);
    if (const char* env = getenv("CVDB_IVF_KERNEL")) transposed = atoi(env) != 0 && k <= 128 && nkb_ <= kIvfMaxKb;
    const int item_cap = transposed ? 16 : 128;
    // --- group the (query, probe) pairs by list
    const int64_t max_items = std::min<int64_t>(nlist, n_pairs) + n_pairs / item_cap + 1;
    const int64_t pairs_pad = n_pairs + 128;  // the last item's A box may run past the last gathered row
    TRY(ix->ivf_cnt.ensure(static_cast<size_t>(nlist) * 4));
    TRY(ix->ivf_cursor.ensure(static_cast<size_t>(nlist) * 4));
    TRY(ix->ivf_pair_off.ensure(static_cast<size_t>(nlist) * 4));
    TRY(ix->ivf_item_off.ensure(static_cast<size_t>(nlist) * 4));
    TRY(ix->ivf_scal.ensure(16));
    TRY(ix->ivf_items.ensure(static_cast<size_t>(max_items) * sizeof(IvfItem)));
    TRY(ix->ivf_pair_query.ensure(static_cast<size_t>(pairs_pad) * 4));
    TRY(ix->ivf_pair_dst.ensure(static_cast<size_t>(pairs_pad) * 4));
    TRY(ix->ivf_qg.ensure(static_cast<size_t>(pairs_pad) * ix->row_elems * 2));
    CU_TRY(cudaMemsetAsync(ix->ivf_cnt.p, 0, static_cast<size_t>(nlist) * 4, st));
    CU_TRY(cudaMemsetAsync(ix->ivf_cursor.p, 0, static_cast<size_t>(nlist) * 4, st));
    const int32_t* list_off = ix->list_off.as<int32_t>();
    ivf_count_pairs_kernel<<<static_cast<unsigned>(ceil_div(n_pairs, 256)), 256, 0, st>>>(probes, n_pairs, nlist, list_off,
                                                                                          ix->ivf_cnt.as<int32_t>());
    ivf_scan_lists_kernel<<<1, 1024, 0, st>>>(ix->ivf_cnt.as<int32_t>(), nlist, ix->ivf_pair_off.as<int32_t>(),
                                               ix->ivf_item_off.as<int32_t>(), ix->ivf_scal.as<int32_t>(), item_cap);
    ivf_make_items_kernel<<<static_cast<unsigned>(ceil_div(nlist, 256)), 256, 0, st>>>(
        ix->ivf_cnt.as<int32_t>(), ix->ivf_pair_off.as<int32_t>(), ix->ivf_item_off.as<int32_t>(), list_off, nlist,
        ix->ivf_items.as<IvfItem>(), item_cap);
    ivf_scatter_pairs_kernel<<<static_cast<unsigned>(ceil_div(n_pairs, 8)), 256, 0, st>>>(
        probes, n_pairs, nprobe, nlist, list_off, ix->ivf_pair_off.as<int32_t>(), ix->ivf_cursor.as<int32_t>(),
        ix->q_pack.as<uint4>(), row_vec16, ix->ivf_pair_query.as<int32_t>(), ix->ivf_pair_dst.as<int32_t>(),
        ix->ivf_qg.as<uint4>());
    g_launches += 4;
    CU_TRY(cudaGetLastError());
    // the item count stays on the device (the kernel reads it): no host synchronisation on this path
    const int n_items = static_cast<int>(std::min<int64_t>(max_items, INT32_MAX));  // upper bound, sizes the grid

    TRY(ix->part.ensure(static_cast<size_t>(std::max<int64_t>(n_pairs, 1)) * k * 8));
    CU_TRY(cudaMemsetAsync(ix->part.p, 0, static_cast<size_t>(n_pairs) * k * 8, st));  // dropped pairs stay empty
    TRY(ix->gthr.ensure(static_cast<size_t>(nq) * 4));
    CU_TRY(cudaMemsetAsync(ix->gthr.p, 0, static_cast<size_t>(nq) * 4, st));
    if (n_items > 0 && transposed) {
        const int grid = std::min(ix->num_sms, n_items);
        IvfScanParams p{};
        p.n_items_ptr = ix->ivf_scal.as<int32_t>();
        p.work_counter = ix->ivf_scal.as<unsigned int>() + 3;
        p.k = k;
        p.trigger = std::min(kIvfCap - 128, std::max(32, 2 * k));
        if (const char* env = getenv("CVDB_IVF_TRIGGER")) p.trigger = std::min(kIvfCap - 128, std::max(k, atoi(env)));
        p.nkb = nkb_;
        p.k16 = static_cast<int>(ceil_div(ix->Kd, 16));
        p.items = reinterpret_cast<const GroupItem*>(ix->ivf_items.p);
        p.pair_query = ix->ivf_pair_query.as<int32_t>();
        p.pair_dst = ix->ivf_pair_dst.as<int32_t>();
        p.row_ids = ix->row_ids.as<int32_t>();
        p.part = ix->part.as<uint64_t>();
        p.gthr = ix->gthr.as<uint32_t>();
        CUtensorMap tq, tx128, tx32;
        TRY(get_tmap(ix->tm_q_ivf, &tq, ix->ivf_qg.p, pairs_pad, ix->row_elems, 16));  // the item's <= 16 query rows
        TRY(get_tmap(ix->tm_x_ivf, &tx128, ix->x, ix->ntotal, ix->row_elems, 128));
        TRY(get_tmap(ix->tm_x32_ivf, &tx32, ix->x, ix->ntotal, ix->row_elems, 32));     // tail tiles of a list
        LAUNCH(launch_ivf_scan(tx128, tx32, tq, p, grid, st));
    } else if (n_items > 0) {
        const int grid = std::min(ix->num_sms, n_items);
        if (E > 0) TRY(ix->cand.ensure(static_cast<size_t>(grid) * 128 * C * 8));
        GroupedParams p{};
        p.n_items_ptr = ix->ivf_scal.as<int32_t>();
        p.k = k;
        p.room = compaction_trigger(k, C);
        p.nkb = static_cast<int>(ceil_div(ix->Kp, 64));
        p.k16 = static_cast<int>(ceil_div(ix->Kd, 16));
        p.items = reinterpret_cast<const GroupItem*>(ix->ivf_items.p);
        p.pair_query = ix->ivf_pair_query.as<int32_t>();
        p.pair_dst = ix->ivf_pair_dst.as<int32_t>();
        p.row_ids = ix->row_ids.as<int32_t>();
        p.cand = ix->cand.as<uint64_t>();
        p.part = ix->part.as<uint64_t>();
        p.gthr = ix->gthr.as<uint32_t>();
        CUtensorMap tq, tx;
        TRY(get_tmap(ix->tm_q_ivf, &tq, ix->ivf_qg.p, pairs_pad, ix->row_elems, 32));  // four 32-row boxes per A tile
        TRY(get_tmap(ix->tm_x_ivf, &tx, ix->x, ix->ntotal, ix->row_elems, kGroupedBlockN));
        LAUNCH(launch_grouped(E, tq, tx, p, grid, st));
    }
    ix->last_flops = 0;
    ix->last_slices = nprobe;
    ix->last_grid = n_items;
    ix->last_variant = transposed ? 6 : 5;
    // candidate keys of the list scan already carry the caller's row ids (GroupedParams::row_ids)
    return launch_merge(ix->part.as<uint64_t>(), nq, nprobe, k, k, l2, ix->q_norm.as<float>(), 0,
                        static_cast<int64_t>(nprobe) * k, k, nullptr, 0, D, I, nullptr, nullptr, st);
}

int check_index(cvdb_index_t h) {
    if (!h) return fail(CVDB_EINVAL, "null index handle");
    return CVDB_OK;
}

// queries per launch: bounds the packed-query and result scratch
int64_t query_chunk(const Index* ix, int k) {
    const int64_t by_pack = std::max<int64_t>(128, (int64_t(256) << 20) / (int64_t(ix->row_elems) * 2));
    const int64_t cap = k == 1 ? 262144 : 65536;
    return std::min(by_pack, cap) / 128 * 128;
}

}  // namespace

// =========================================================================== C ABI
extern "C" {

const char* cvdb_last_error(void) { return g_err.c_str(); }
int64_t cvdb_kernel_launches(void) { return g_launches.load(); }
int cvdb_version(void) { return 100; }

int cvdb_index_create(int d, int metric, int storage, int device, cvdb_index_t* out) {
    if (!out) return fail(CVDB_EINVAL, "out is null");
    *out = nullptr;
    if (d < 1 || d > 16384) return fail(CVDB_ELIMIT, "d=%d outside [1, 16384]", d);
    if (metric != CVDB_METRIC_IP && metric != CVDB_METRIC_L2) return fail(CVDB_EINVAL, "unknown metric %d", metric);
    if (storage != CVDB_STORE_BF16 && storage != CVDB_STORE_EXACT) return fail(CVDB_EINVAL, "unknown storage %d", storage);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(CVDB_ECUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(CVDB_EINVAL, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(CVDB_ECUDA, "device %d is sm_%d%d; the kernels are built for sm_100a (B200) only", device, prop.major,
                    prop.minor);
    Index* ix = new Index();
    ix->d = d;
    ix->metric = metric;
    ix->storage = storage;
    ix->device = device;
    ix->planes = storage == CVDB_STORE_EXACT ? 3 : 1;
    const int extra = metric == CVDB_METRIC_L2 ? 3 : 0;
    ix->Kd = d + extra;
    // Row width: whole 64-column K blocks when that costs at most 25 % more memory -- a TMA box that is partly out
    // of bounds along K loads measurably slower (L2 at d=768: 84.6k -> 94.9k queries/s with the 13th block padded) --
    // else the next multiple of 8 (16-byte rows for TMA).  CVDB_PAD64=0/1 overrides the rule (experiments).
    bool pad64 = storage == CVDB_STORE_EXACT || round_up(ix->Kd, 64) * 4 <= round_up(ix->Kd, 8) * 5;
    if (const char* env = getenv("CVDB_PAD64")) pad64 = storage == CVDB_STORE_EXACT || atoi(env) != 0;
    ix->Kp = pad64 ? round_up(ix->Kd, 64) : round_up(ix->Kd, 8);
    ix->row_elems = ix->planes * ix->Kp;
    ix->num_sms = prop.multiProcessorCount;
    {
        cvdb_guard g(device);
        if (ix->bad_rows.ensure(8) != CVDB_OK || cudaMemset(ix->bad_rows.p, 0, 8) != cudaSuccess ||
            cudaStreamSynchronize(nullptr) != cudaSuccess) {
            ix->bad_rows.release();
            delete ix;
            return fail(CVDB_ENOMEM, "cudaMalloc for the index state failed");
        }
    }
    *out = reinterpret_cast<cvdb_index_t>(ix);
    return CVDB_OK;
}

int cvdb_index_destroy(cvdb_index_t h) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    cvdb_guard g(ix->device);
    cudaDeviceSynchronize();
    if (ix->x) cudaFree(ix->x);
    for (DevBuf* b : {&ix->row_ids, &ix->list_off, &ix->ivf_cnt, &ix->ivf_pair_off, &ix->ivf_item_off, &ix->ivf_cursor,
                      &ix->ivf_scal, &ix->ivf_items, &ix->ivf_pair_query, &ix->ivf_pair_dst, &ix->ivf_qg, &ix->ivf_probes})
        b->release();
    for (DevBuf* b : {&ix->col_thr, &ix->col_cnt, &ix->col_base, &ix->col_buf, &ix->col_dirty, &ix->col_scal, &ix->col_log})
        b->release();
    for (DevBuf* b : {&ix->gthr, &ix->waves, &ix->stage_in, &ix->q_pack, &ix->q_norm, &ix->cand, &ix->part, &ix->out_d, &ix->out_i, &ix->ids_a,
                      &ix->ids_b, &ix->groups, &ix->bad_rows})
        b->release();
    for (int i = 0; i < Index::kProfSlots; ++i) {
        if (ix->ev0[i]) cudaEventDestroy(ix->ev0[i]);
        if (ix->ev1[i]) cudaEventDestroy(ix->ev1[i]);
    }
    for (int b = 0; b < 2; ++b) {
        if (ix->pin[b]) cudaFreeHost(ix->pin[b]);
        if (ix->pin_free[b]) cudaEventDestroy(ix->pin_free[b]);
        if (ix->stage_free[b]) cudaEventDestroy(ix->stage_free[b]);
        ix->stage2[b].release();
    }
    if (ix->copy_stream) cudaStreamDestroy(ix->copy_stream);
    if (ix->done_ev) cudaEventDestroy(ix->done_ev);
    delete ix;
    return CVDB_OK;
}

int cvdb_index_reset(cvdb_index_t h) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    ix->ntotal = 0;
    ix->has_groups = false;
    ix->grouped = false;
    ix->row_ids_n = 0;
    return CVDB_OK;
}

int cvdb_index_truncate(cvdb_index_t h, int64_t n) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (n < 0 || n > ix->ntotal)
        return fail(CVDB_EINVAL, "truncate to %lld rows: the index holds %lld", static_cast<long long>(n),
                    static_cast<long long>(ix->ntotal));
    if (ix->permuted() && n < ix->row_ids_n)
        return fail(CVDB_EINVAL, "rows are stored list-major after cvdb_index_group_by_list; the last rows added are not "
                                 "the last rows stored (only rows added since the grouping can be dropped; cvdb_index_reset clears)");
    ix->ntotal = n;
    return CVDB_OK;
}

int cvdb_index_nonfinite_rows(cvdb_index_t h, int64_t* count_out, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (!count_out) return fail(CVDB_EINVAL, "count_out is null");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long v = 0;
    CU_TRY(cudaMemcpyAsync(&v, ix->bad_rows.p, 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    *count_out = static_cast<int64_t>(v);
    return CVDB_OK;
}

int cvdb_index_reserve(cvdb_index_t h, int64_t n) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (n < 0 || n > 0x7FFFFF00LL) return fail(CVDB_ELIMIT, "n=%lld outside [0, 2^31)", static_cast<long long>(n));
    cvdb_guard g(ix->device);
    return grow(ix, n, nullptr);
}

int64_t cvdb_index_ntotal(cvdb_index_t h) { return h ? reinterpret_cast<Index*>(h)->ntotal : -1; }
int cvdb_index_dim(cvdb_index_t h) { return h ? reinterpret_cast<Index*>(h)->d : -1; }

int cvdb_index_add(cvdb_index_t h, const void* x, int64_t n, int dtype, int on_device, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (n < 0) return fail(CVDB_EINVAL, "n < 0");
    if (n == 0) return CVDB_OK;
    if (!x) return fail(CVDB_EINVAL, "x is null");
    if (dtype != CVDB_DTYPE_F32 && dtype != CVDB_DTYPE_BF16 && dtype != CVDB_DTYPE_F16)
        return fail(CVDB_EINVAL, "unknown dtype %d", dtype);
    if (ix->ntotal + n > 0x7FFFFF00LL) return fail(CVDB_ELIMIT, "an index holds fewer than 2^31 rows per GPU");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TRY(grow(ix, ix->ntotal + n, st));
    const size_t esz = dtype == CVDB_DTYPE_F32 ? 4 : 2;  // bf16 and fp16 are 2 bytes
    __nv_bfloat16* dst = ix->x + ix->ntotal * ix->row_elems;
    StreamOrder order(ix, st);
    if (on_device) {
        TRY(pack_dispatch(x, dtype, n, dst, ix, 0, nullptr, st));
    } else {
        TRY(ingest_host_rows(ix, static_cast<const char*>(x), n, dtype, esz, dst, st));
    }
    ix->ntotal += n;
    ix->has_groups = false;
    ix->grouped = false;  // new rows are not in any list yet
    return CVDB_OK;
}

int cvdb_index_set_groups(cvdb_index_t h, const int32_t* group_db, int on_device, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (!group_db) {
        ix->has_groups = false;
        return CVDB_OK;
    }
    if (ix->permuted())
        return fail(CVDB_EINVAL, "group ids are indexed by stored position, and cvdb_index_group_by_list re-stored the rows "
                                 "list-major: exclusion is not available on such an index (cvdb_index_reset clears)");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TRY(ix->groups.ensure(static_cast<size_t>(std::max<int64_t>(ix->ntotal, 1)) * 4));
    CU_TRY(cudaMemcpyAsync(ix->groups.p, group_db, static_cast<size_t>(ix->ntotal) * 4,
                           on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (!on_device) CU_TRY(cudaStreamSynchronize(st));
    ix->has_groups = true;
    return CVDB_OK;
}

int cvdb_index_search(cvdb_index_t h, const void* q, int64_t nq, int dtype, int k, float* D, int64_t* I, int on_device,
                      const cvdb_search_opts* opts, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (nq < 0) return fail(CVDB_EINVAL, "nq < 0");
    if (k < 1 || k > CVDB_MAX_K) return fail(CVDB_ELIMIT, "k=%d outside [1, %d]", k, CVDB_MAX_K);
    if (dtype != CVDB_DTYPE_F32 && dtype != CVDB_DTYPE_BF16 && dtype != CVDB_DTYPE_F16)
        return fail(CVDB_EINVAL, "unknown dtype %d", dtype);
    if (nq == 0) return CVDB_OK;
    if (!q || !D || !I) return fail(CVDB_EINVAL, "q, D and I must be non-null");
    if (opts && opts->group_q && !ix->has_groups)
        return fail(CVDB_EINVAL, "group_q given but the index has no groups (cvdb_index_set_groups)");
    if (ix->permuted() && opts && (opts->self_ids || opts->group_q))
        return fail(CVDB_EINVAL, "self / group exclusion works on stored positions, and cvdb_index_group_by_list re-stored "
                                 "the rows list-major: not available on such an index");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamOrder order(ix, st);
    const size_t esz = dtype == CVDB_DTYPE_F32 ? 4 : 2;  // bf16 and fp16 are 2 bytes
    const int64_t chunk = query_chunk(ix, k);
    for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
        const int64_t m = std::min(chunk, nq - q0);
        const void* qsrc = static_cast<const char*>(q) + q0 * ix->d * esz;
        const int32_t* self = (opts && opts->self_ids) ? opts->self_ids + q0 : nullptr;
        const int32_t* grp = (opts && opts->group_q) ? opts->group_q + q0 : nullptr;
        float* Dd = D + q0 * k;
        int64_t* Id = I + q0 * k;
        if (!on_device) {
            TRY(ix->stage_in.ensure(static_cast<size_t>(m) * ix->d * esz));
            CU_TRY(cudaMemcpyAsync(ix->stage_in.p, qsrc, static_cast<size_t>(m) * ix->d * esz, cudaMemcpyHostToDevice, st));
            qsrc = ix->stage_in.p;
            if (self) {
                TRY(ix->ids_a.ensure(static_cast<size_t>(m) * 4));
                CU_TRY(cudaMemcpyAsync(ix->ids_a.p, self, static_cast<size_t>(m) * 4, cudaMemcpyHostToDevice, st));
                self = ix->ids_a.as<int32_t>();
            }
            if (grp) {
                TRY(ix->ids_b.ensure(static_cast<size_t>(m) * 4));
                CU_TRY(cudaMemcpyAsync(ix->ids_b.p, grp, static_cast<size_t>(m) * 4, cudaMemcpyHostToDevice, st));
                grp = ix->ids_b.as<int32_t>();
            }
            TRY(ix->out_d.ensure(static_cast<size_t>(m) * k * 4));
            TRY(ix->out_i.ensure(static_cast<size_t>(m) * k * 8));
            Dd = ix->out_d.as<float>();
            Id = ix->out_i.as<int64_t>();
        }
        TRY(search_device(ix, qsrc, m, dtype, k, Dd, Id, nullptr, self, grp, opts, st));
        if (!on_device) {
            CU_TRY(cudaMemcpyAsync(D + q0 * k, Dd, static_cast<size_t>(m) * k * 4, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaMemcpyAsync(I + q0 * k, Id, static_cast<size_t>(m) * k * 8, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
        }
    }
    return CVDB_OK;
}

int cvdb_index_assign(cvdb_index_t h, const void* x, int64_t n, int dtype, int32_t* assign, float* dist, int on_device,
                      void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (n < 0) return fail(CVDB_EINVAL, "n < 0");
    if (dtype != CVDB_DTYPE_F32 && dtype != CVDB_DTYPE_BF16 && dtype != CVDB_DTYPE_F16)
        return fail(CVDB_EINVAL, "unknown dtype %d", dtype);
    if (n == 0) return CVDB_OK;
    if (!x || !assign) return fail(CVDB_EINVAL, "x and assign must be non-null");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamOrder order(ix, st);
    const size_t esz = dtype == CVDB_DTYPE_F32 ? 4 : 2;  // bf16 and fp16 are 2 bytes
    const int64_t chunk = query_chunk(ix, 1);
    for (int64_t q0 = 0; q0 < n; q0 += chunk) {
        const int64_t m = std::min(chunk, n - q0);
        const void* qsrc = static_cast<const char*>(x) + q0 * ix->d * esz;
        int32_t* Ad = assign + q0;
        float* Dd = dist ? dist + q0 : nullptr;
        if (!on_device) {
            TRY(ix->stage_in.ensure(static_cast<size_t>(m) * ix->d * esz));
            CU_TRY(cudaMemcpyAsync(ix->stage_in.p, qsrc, static_cast<size_t>(m) * ix->d * esz, cudaMemcpyHostToDevice, st));
            qsrc = ix->stage_in.p;
            TRY(ix->out_i.ensure(static_cast<size_t>(m) * 4));
            Ad = ix->out_i.as<int32_t>();
        }
        if (!on_device || !Dd) {
            TRY(ix->out_d.ensure(static_cast<size_t>(m) * 4));
            Dd = ix->out_d.as<float>();
        }
        TRY(search_device(ix, qsrc, m, dtype, 1, Dd, nullptr, Ad, nullptr, nullptr, nullptr, st));
        if (!on_device) {
            CU_TRY(cudaMemcpyAsync(assign + q0, Ad, static_cast<size_t>(m) * 4, cudaMemcpyDeviceToHost, st));
            if (dist) CU_TRY(cudaMemcpyAsync(dist + q0, Dd, static_cast<size_t>(m) * 4, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
        }
    }
    return CVDB_OK;
}

int cvdb_index_profile_ms(cvdb_index_t h, float* out_ms, int max_n) {
    if (!h || !out_ms || max_n < 0) return fail(CVDB_EINVAL, "bad arguments");
    Index* ix = reinterpret_cast<Index*>(h);
    cvdb_guard g(ix->device);
    const int n = std::min(ix->prof_count, max_n);
    // oldest first
    for (int i = 0; i < n; ++i) {
        const int slot = ((ix->prof_head - ix->prof_count + i) % Index::kProfSlots + Index::kProfSlots) % Index::kProfSlots;
        if (cudaEventSynchronize(ix->ev1[slot]) != cudaSuccess ||
            cudaEventElapsedTime(&out_ms[i], ix->ev0[slot], ix->ev1[slot]) != cudaSuccess)
            return fail(CVDB_ECUDA, "profiling events could not be read");
    }
    ix->prof_count = 0;
    return n;
}

float cvdb_index_last_kernel_ms(cvdb_index_t h) {
    if (!h) return -1.f;
    Index* ix = reinterpret_cast<Index*>(h);
    if (ix->prof_count < 1) return -1.f;
    cvdb_guard g(ix->device);
    const int slot = (ix->prof_head - 1 + Index::kProfSlots) % Index::kProfSlots;
    if (cudaEventSynchronize(ix->ev1[slot]) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, ix->ev0[slot], ix->ev1[slot]) != cudaSuccess) return -1.f;
    return ms;
}

int cvdb_index_last_work(cvdb_index_t h, double* flops, double* db_bytes, int* n_slices, int* grid) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (flops) *flops = ix->last_flops;
    if (db_bytes) *db_bytes = ix->last_bytes;
    if (n_slices) *n_slices = ix->last_slices;
    if (grid) *grid = ix->last_grid;
    return CVDB_OK;
}

int cvdb_index_last_variant(cvdb_index_t h) { return h ? reinterpret_cast<Index*>(h)->last_variant : -1; }

int cvdb_index_group_by_list(cvdb_index_t h, const int32_t* list_of_id, int nlist, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (ix->planes != 1) return fail(CVDB_EINVAL, "inverted lists need bf16 storage");
    if (nlist < 1) return fail(CVDB_EINVAL, "nlist < 1");
    if (!list_of_id) return fail(CVDB_EINVAL, "null pointer");
    if (ix->ntotal == 0) return fail(CVDB_EINVAL, "empty index");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamOrder order(ix, st);
    const int64_t n = ix->ntotal;
    const int row_vec16 = ix->row_elems * 2 / 16;
    // ids of the rows as stored now: what an earlier grouping left, then insertion order for rows added since
    TRY(ix->row_ids.grow_keep(static_cast<size_t>(n) * 4, static_cast<size_t>(ix->row_ids_n) * 4, st));
    if (ix->row_ids_n < n) {
        const int64_t m = n - ix->row_ids_n;
        iota_kernel<<<static_cast<unsigned>(ceil_div(m, 256)), 256, 0, st>>>(ix->row_ids.as<int32_t>() + ix->row_ids_n,
                                                                              ix->row_ids_n, m);
        ++g_launches;
        ix->row_ids_n = n;
    }
    TRY(ix->ivf_cnt.ensure(static_cast<size_t>(nlist) * 4));
    TRY(ix->ivf_cursor.ensure(static_cast<size_t>(nlist) * 4));
    TRY(ix->ivf_item_off.ensure(static_cast<size_t>(nlist) * 4));
    TRY(ix->ivf_scal.ensure(16));
    TRY(ix->list_off.ensure(static_cast<size_t>(nlist + 1) * 4));
    TRY(ix->ids_a.ensure(static_cast<size_t>(n) * 4));  // new row ids
    CU_TRY(cudaMemsetAsync(ix->ivf_cnt.p, 0, static_cast<size_t>(nlist) * 4, st));
    CU_TRY(cudaMemsetAsync(ix->ivf_cursor.p, 0, static_cast<size_t>(nlist) * 4, st));
    CU_TRY(cudaMemsetAsync(ix->ivf_scal.p, 0, 16, st));
    int32_t* bad = ix->ivf_scal.as<int32_t>() + 2;
    ivf_count_rows_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, st>>>(list_of_id, ix->row_ids.as<int32_t>(), n,
                                                                                  nlist, ix->ivf_cnt.as<int32_t>(), bad);
    ivf_scan_lists_kernel<<<1, 1024, 0, st>>>(ix->ivf_cnt.as<int32_t>(), nlist, ix->list_off.as<int32_t>(),
                                               ix->ivf_item_off.as<int32_t>(), ix->ivf_scal.as<int32_t>());
    ivf_close_offsets_kernel<<<1, 32, 0, st>>>(ix->list_off.as<int32_t>(), nlist, ix->ivf_scal.as<int32_t>());
    g_launches += 3;
    int32_t n_bad = 0;
    CU_TRY(cudaMemcpyAsync(&n_bad, bad, 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (n_bad) return fail(CVDB_EINVAL, "%d rows name a list outside [0, %d)", n_bad, nlist);
    const size_t bytes = static_cast<size_t>(n) * ix->row_elems * 2;
    __nv_bfloat16* nx = nullptr;
    cudaError_t e = cudaMalloc(&nx, bytes);
    if (e != cudaSuccess) return fail(CVDB_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    ivf_scatter_rows_kernel<<<static_cast<unsigned>(ceil_div(n, 8)), 256, 0, st>>>(
        list_of_id, ix->row_ids.as<int32_t>(), n, nlist, ix->list_off.as<int32_t>(), ix->ivf_cursor.as<int32_t>(),
        reinterpret_cast<const uint4*>(ix->x), row_vec16, reinterpret_cast<uint4*>(nx), ix->ids_a.as<int32_t>());
    ++g_launches;
    e = cudaMemcpyAsync(ix->row_ids.p, ix->ids_a.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(nx);
        return fail(CVDB_ECUDA, "regrouping the rows failed: %s", cudaGetErrorString(e));
    }
    cudaFree(ix->x);
    ix->x = nx;
    ix->capacity = n;
    ix->grouped = true;
    ix->nlist = nlist;
    ix->has_groups = false;
    return CVDB_OK;
}

int cvdb_index_list_offsets(cvdb_index_t h, int32_t* out_device, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (!ix->grouped) return fail(CVDB_EINVAL, "rows are not grouped into lists");
    if (!out_device) return fail(CVDB_EINVAL, "null pointer");
    cvdb_guard g(ix->device);
    CU_TRY(cudaMemcpyAsync(out_device, ix->list_off.p, static_cast<size_t>(ix->nlist + 1) * 4, cudaMemcpyDeviceToDevice,
                           static_cast<cudaStream_t>(stream)));
    return CVDB_OK;
}

int cvdb_index_search_lists(cvdb_index_t h, const void* q, int64_t nq, int dtype, int k, const int32_t* probes, int nprobe,
                            float* D, int64_t* I, int on_device, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (!ix->grouped) return fail(CVDB_EINVAL, "rows are not grouped into lists (cvdb_index_group_by_list)");
    if (nq < 0) return fail(CVDB_EINVAL, "nq < 0");
    if (k < 1 || k > CVDB_MAX_K) return fail(CVDB_ELIMIT, "k=%d outside [1, %d]", k, CVDB_MAX_K);
    if (nprobe < 1 || nprobe > 2048) return fail(CVDB_ELIMIT, "nprobe=%d outside [1, 2048]", nprobe);
    if (dtype != CVDB_DTYPE_F32 && dtype != CVDB_DTYPE_BF16 && dtype != CVDB_DTYPE_F16)
        return fail(CVDB_EINVAL, "unknown dtype %d", dtype);
    if (nq == 0) return CVDB_OK;
    if (!q || !probes || !D || !I) return fail(CVDB_EINVAL, "q, probes, D and I must be non-null");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamOrder order(ix, st);
    const size_t esz = dtype == CVDB_DTYPE_F32 ? 4 : 2;  // bf16 and fp16 are 2 bytes
    // bound the gathered-query scratch: at most 2^18 (query, probe) pairs per launch
    const int64_t chunk = std::max<int64_t>(1, (int64_t(1) << 18) / nprobe);
    for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
        const int64_t m = std::min(chunk, nq - q0);
        const void* qsrc = static_cast<const char*>(q) + q0 * ix->d * esz;
        const int32_t* psrc = probes + q0 * nprobe;
        float* Dd = D + q0 * k;
        int64_t* Id = I + q0 * k;
        if (!on_device) {
            TRY(ix->stage_in.ensure(static_cast<size_t>(m) * ix->d * esz));
            CU_TRY(cudaMemcpyAsync(ix->stage_in.p, qsrc, static_cast<size_t>(m) * ix->d * esz, cudaMemcpyHostToDevice, st));
            qsrc = ix->stage_in.p;
            TRY(ix->ivf_probes.ensure(static_cast<size_t>(m) * nprobe * 4));
            CU_TRY(cudaMemcpyAsync(ix->ivf_probes.p, psrc, static_cast<size_t>(m) * nprobe * 4, cudaMemcpyHostToDevice, st));
            psrc = ix->ivf_probes.as<int32_t>();
            TRY(ix->out_d.ensure(static_cast<size_t>(m) * k * 4));
            TRY(ix->out_i.ensure(static_cast<size_t>(m) * k * 8));
            Dd = ix->out_d.as<float>();
            Id = ix->out_i.as<int64_t>();
        }
        TRY(search_lists_device(ix, qsrc, m, dtype, k, psrc, nprobe, Dd, Id, st));
        if (!on_device) {
            CU_TRY(cudaMemcpyAsync(D + q0 * k, Dd, static_cast<size_t>(m) * k * 4, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaMemcpyAsync(I + q0 * k, Id, static_cast<size_t>(m) * k * 8, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
        }
    }
    return CVDB_OK;
}

int cvdb_build_triplets(const int64_t* I, const float* D, int64_t n, int k, const int64_t* pos, int64_t anchor_base,
                        int skip_top, int per_anchor, int metric, float limit, int use_limit, int64_t* out, void* stream) {
    if (n < 0 || k < 1 || skip_top < 0 || per_anchor < 1) return fail(CVDB_EINVAL, "bad sizes");
    if (metric != CVDB_METRIC_IP && metric != CVDB_METRIC_L2) return fail(CVDB_EINVAL, "unknown metric %d", metric);
    if (n == 0) return CVDB_OK;
    if (!I || !D || !pos || !out) return fail(CVDB_EINVAL, "null pointer");
    ptr_guard g(I);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    build_triplets_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, st>>>(I, D, n, k, pos, anchor_base, skip_top,
                                                                                    per_anchor, metric == CVDB_METRIC_L2,
                                                                                    limit, use_limit, out);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return CVDB_OK;
}

int64_t cvdb_index_row_bytes(cvdb_index_t h) { return h ? int64_t(reinterpret_cast<Index*>(h)->row_elems) * 2 : -1; }

int cvdb_index_export_rows(cvdb_index_t h, int64_t row0, int64_t nrows, void* host_dst, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (row0 < 0 || nrows < 0 || row0 + nrows > ix->ntotal) return fail(CVDB_EINVAL, "row range outside the index");
    if (nrows == 0) return CVDB_OK;
    if (!host_dst) return fail(CVDB_EINVAL, "null pointer");
    if (ix->permuted())
        return fail(CVDB_EINVAL, "rows are stored list-major (cvdb_index_group_by_list); the dump format has no row ids: "
                                 "export the index before grouping it");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t rb = static_cast<size_t>(ix->row_elems) * 2;
    CU_TRY(cudaMemcpyAsync(host_dst, reinterpret_cast<const char*>(ix->x) + row0 * rb, nrows * rb, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return CVDB_OK;
}

int cvdb_index_import_rows(cvdb_index_t h, const void* host_src, int64_t nrows, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (nrows < 0) return fail(CVDB_EINVAL, "nrows < 0");
    if (nrows == 0) return CVDB_OK;
    if (!host_src) return fail(CVDB_EINVAL, "null pointer");
    if (ix->ntotal + nrows > 0x7FFFFF00LL) return fail(CVDB_ELIMIT, "an index holds fewer than 2^31 rows per GPU");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TRY(grow(ix, ix->ntotal + nrows, st));
    const size_t rb = static_cast<size_t>(ix->row_elems) * 2;
    CU_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(ix->x) + ix->ntotal * rb, host_src, nrows * rb, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));
    ix->ntotal += nrows;
    ix->has_groups = false;
    ix->grouped = false;
    return CVDB_OK;
}

int cvdb_plan_slices(int q_tiles, int n_tiles, int workers, int64_t max_slices, int* n_slices, int* tiles_per_slice) {
    if (q_tiles < 1 || n_tiles < 1 || workers < 1 || !n_slices || !tiles_per_slice) return fail(CVDB_EINVAL, "bad arguments");
    choose_slices(q_tiles, n_tiles, workers, max_slices, *n_slices, *tiles_per_slice);
    return CVDB_OK;
}

int cvdb_merge_topk(const float* Dc, const int64_t* Ic, int64_t nq, int nlists, int k_in, int k, int metric, float* D,
                    int64_t* I, int on_device, void* stream) {
    if (nq < 0 || nlists < 1 || k_in < 1 || k < 1) return fail(CVDB_EINVAL, "bad sizes");
    if (metric != CVDB_METRIC_IP && metric != CVDB_METRIC_L2) return fail(CVDB_EINVAL, "unknown metric %d", metric);
    if (nq == 0) return CVDB_OK;
    if (!Dc || !Ic || !D || !I) return fail(CVDB_EINVAL, "null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t n_in = static_cast<size_t>(nq) * nlists * k_in, n_out = static_cast<size_t>(nq) * k;
    const unsigned blocks = static_cast<unsigned>(ceil_div(nq, 8));
    if (on_device) {
        ptr_guard g(Dc);
        merge_lists_kernel<<<blocks, 256, 0, st>>>(Dc, Ic, nq, nlists, k_in, k, metric == CVDB_METRIC_L2, D, I);
        ++g_launches;
        CU_TRY(cudaGetLastError());
        return CVDB_OK;
    }
    // host pointers: one stream-ordered scratch allocation (pooled by the driver, no cudaMalloc per call),
    // laid out [Dc | D | Ic | I] so the 8-byte arrays stay aligned; every copy is checked
    const size_t off_d = (n_in * 4 + 255) & ~size_t(255);
    const size_t off_ic = (off_d + n_out * 4 + 255) & ~size_t(255);
    const size_t off_i = off_ic + n_in * 8;
    char* scratch = nullptr;
    CU_TRY(cudaMallocAsync(reinterpret_cast<void**>(&scratch), off_i + n_out * 8, st));
    cudaError_t e = cudaMemcpyAsync(scratch, Dc, n_in * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(scratch + off_ic, Ic, n_in * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        merge_lists_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(scratch),
                                                   reinterpret_cast<const int64_t*>(scratch + off_ic), nq, nlists, k_in, k,
                                                   metric == CVDB_METRIC_L2, reinterpret_cast<float*>(scratch + off_d),
                                                   reinterpret_cast<int64_t*>(scratch + off_i));
        ++g_launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(D, scratch + off_d, n_out * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(I, scratch + off_i, n_out * 8, cudaMemcpyDeviceToHost, st);
    const cudaError_t ef = cudaFreeAsync(scratch, st);
    const cudaError_t es = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = ef != cudaSuccess ? ef : es;
    if (e != cudaSuccess) return fail(CVDB_ECUDA, "merge failed: %s", cudaGetErrorString(e));
    return CVDB_OK;
}

// ------------------------------------------------------------ symmetric self-join
int cvdb_selfjoin_begin(cvdb_index_t h, int k, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (ix->metric != CVDB_METRIC_IP || ix->planes != 1)
        return fail(CVDB_EINVAL, "the symmetric self-join needs an inner-product index with bf16 storage");
    if (ceil_div(ix->Kp, 64) > 12) return fail(CVDB_ELIMIT, "the symmetric self-join needs padded d <= 768");
    if (k < 2 || k > 124) return fail(CVDB_ELIMIT, "k=%d outside [2, 124]", k);
    if (ix->permuted()) return fail(CVDB_EINVAL, "rows are stored list-major: ids and stored positions differ");
    if (ix->ntotal < 1) return fail(CVDB_EINVAL, "empty index");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamOrder order(ix, st);
    const size_t n = static_cast<size_t>(ix->ntotal);
    TRY(ix->col_thr.ensure(n * 4));
    TRY(ix->col_cnt.ensure(n * 4));
    TRY(ix->col_base.ensure(n * 4));
    TRY(ix->col_dirty.ensure(n));
    TRY(ix->col_scal.ensure(32));  // [0] dirty-row count, [1] log overflow flag, [2] log cursor
    TRY(ix->col_buf.ensure(n * kColCap * 8));
    // the log takes the column candidates of ONE launch: about k per collecting row while chunks at most double (threads
    // reserve it in segments of 8 records).  The cursor is 32 bits wide:
    // the capacity stays below 2^31 records and a record past the capacity is dropped, not wrapped (the overflow flag
    // sends the affected rows to the exact fallback)
    ix->col_log_cap = static_cast<uint32_t>(std::min<unsigned long long>(
        (1ull << 31) - (1ull << 24), std::max<unsigned long long>(1ull << 22, 2ull * n * static_cast<unsigned long long>(k))));
    TRY(ix->col_log.ensure(static_cast<size_t>(ix->col_log_cap) * 16));
    CU_TRY(cudaMemsetAsync(ix->col_log.p, 0, static_cast<size_t>(ix->col_log_cap) * 16, st));
    CU_TRY(cudaMemsetAsync(ix->col_scal.p, 0, 32, st));
    CU_TRY(cudaMemsetAsync(ix->col_cnt.p, 0, n * 4, st));
    CU_TRY(cudaMemsetAsync(ix->col_base.p, 0, n * 4, st));
    CU_TRY(cudaMemsetAsync(ix->col_dirty.p, 0, n, st));
    fill_f32_kernel<<<static_cast<unsigned>(ceil_div(ix->ntotal, 256)), 256, 0, st>>>(ix->col_thr.as<float>(), ix->ntotal, -INFINITY);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    ix->sj_k = k;
    return CVDB_OK;
}

namespace {
int selfjoin_compact(Index* ix, int64_t row_min, int64_t row_end, cudaStream_t st) {
    // the launch's log of column candidates -> the rows' buffers
    col_scatter_kernel<<<148 * 16, 256, 0, st>>>(ix->col_log.as<uint4>(), ix->col_scal.as<uint32_t>() + 4, ix->col_log_cap,
                                                 ix->col_buf.as<uint64_t>(), ix->col_cnt.as<uint32_t>(),
                                                 ix->has_groups ? ix->groups.as<int32_t>() : nullptr);
    col_log_reset_kernel<<<1, 32, 0, st>>>(ix->col_scal.as<uint32_t>() + 4, ix->col_log_cap,
                                           ix->col_scal.as<unsigned long long>() + 1);
    g_launches += 2;
    CU_TRY(cudaGetLastError());
    if (row_min >= row_end) return CVDB_OK;
    const int64_t blocks = std::min<int64_t>(ceil_div(row_end - row_min, 8 * 32), 148 * 32);  // a warp per 32 rows
    col_compact_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(ix->col_buf.as<uint64_t>(), ix->col_cnt.as<uint32_t>(),
                                                                      ix->col_base.as<uint32_t>(), ix->col_thr.as<float>(),
                                                                      ix->col_dirty.as<uint8_t>(), ix->sj_k, row_min, row_end);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return CVDB_OK;
}
}  // namespace

int cvdb_selfjoin_chunk(cvdb_index_t h, int64_t row0, int64_t nrows, int64_t id_base, uint64_t* keys, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (ix->sj_k < 2) return fail(CVDB_EINVAL, "no self-join is open (cvdb_selfjoin_begin)");
    if (row0 < 0 || nrows < 1 || row0 + nrows > ix->ntotal) return fail(CVDB_EINVAL, "anchor rows outside the index");
    if (row0 % 128) return fail(CVDB_EINVAL, "row0 must be a multiple of 128 (the scan starts at a tile boundary)");
    if (nrows > 65536) return fail(CVDB_ELIMIT, "at most 65536 anchors per chunk");
    if (id_base < 0 || id_base + ix->ntotal > 0xFFFFFFFELL) return fail(CVDB_ELIMIT, "ids must stay below 2^32 - 1");
    if (!keys) return fail(CVDB_EINVAL, "null pointer");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamOrder order(ix, st);
    const int k = ix->sj_k;
    TRY(ix->ids_a.ensure(static_cast<size_t>(nrows) * 4));
    TRY(ix->ids_b.ensure(static_cast<size_t>(nrows) * 4));
    const unsigned ib = static_cast<unsigned>(ceil_div(nrows, 256));
    iota_kernel<<<ib, 256, 0, st>>>(ix->ids_a.as<int32_t>(), row0, nrows);            // the row that IS the anchor
    iota_kernel<<<ib, 256, 0, st>>>(ix->ids_b.as<int32_t>(), row0 + id_base, nrows);  // its id in the column lists
    g_launches += 2;
    SearchExtra ex;
    ex.q_packed = ix->x + row0 * ix->row_elems;
    ex.q_rows_avail = ix->ntotal - row0;
    ex.row_begin = row0;
    ex.col = true;
    ex.col_row_min = row0 + nrows;
    ex.q_ids = ix->ids_b.as<int32_t>();
    const int32_t* grp = ix->has_groups ? ix->groups.as<int32_t>() + row0 : nullptr;
    cvdb_search_opts o{};
    o.id_base = id_base;
    TRY(search_device(ix, nullptr, nrows, CVDB_DTYPE_BF16, k, nullptr, nullptr, nullptr, ix->ids_a.as<int32_t>(), grp, &o, st,
                      keys, nullptr, &ex));
    return selfjoin_compact(ix, row0 + nrows, ix->ntotal, st);
}

int cvdb_selfjoin_seed(cvdb_index_t h, int64_t seed_rows, int64_t id_base, uint64_t* keys_seed, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (ix->sj_k < 2) return fail(CVDB_EINVAL, "no self-join is open (cvdb_selfjoin_begin)");
    if (seed_rows < 1 || seed_rows > 65536 || seed_rows > ix->ntotal || (seed_rows % 128 && seed_rows != ix->ntotal))
        return fail(CVDB_EINVAL, "seed_rows must be a multiple of 128 in [128, min(65536, ntotal)] (or all rows)");
    if (id_base < 0 || id_base + ix->ntotal > 0xFFFFFFFELL) return fail(CVDB_ELIMIT, "ids must stay below 2^32 - 1");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamOrder order(ix, st);
    const int k = ix->sj_k;
    cvdb_search_opts o{};
    o.id_base = id_base;
    const int32_t* groups = ix->has_groups ? ix->groups.as<int32_t>() : nullptr;
    if (keys_seed) {
        // (a) the seed anchors themselves: a plain search of rows [0, seed_rows) over ALL rows of this index
        TRY(ix->ids_a.ensure(static_cast<size_t>(seed_rows) * 4));
        iota_kernel<<<static_cast<unsigned>(ceil_div(seed_rows, 256)), 256, 0, st>>>(ix->ids_a.as<int32_t>(), 0, seed_rows);
        ++g_launches;
        SearchExtra ex;
        ex.q_packed = ix->x;
        ex.q_rows_avail = ix->ntotal;
        TRY(search_device(ix, nullptr, seed_rows, CVDB_DTYPE_BF16, k, nullptr, nullptr, nullptr, ix->ids_a.as<int32_t>(), groups, &o,
                          st, keys_seed, nullptr, &ex));
    }
    // (b) every later row: its top-k among the seed anchors starts its column list (and its threshold)
    TRY(ix->out_i.ensure(static_cast<size_t>(65536) * k * 8));
    for (int64_t q0 = seed_rows; q0 < ix->ntotal; q0 += 65536) {
        const int64_t m = std::min<int64_t>(65536, ix->ntotal - q0);
        SearchExtra ex;
        ex.q_packed = ix->x + q0 * ix->row_elems;
        ex.q_rows_avail = ix->ntotal - q0;
        ex.row_end = seed_rows;
        TRY(search_device(ix, nullptr, m, CVDB_DTYPE_BF16, k, nullptr, nullptr, nullptr, nullptr, groups ? groups + q0 : nullptr, &o,
                          st, ix->out_i.as<uint64_t>(), nullptr, &ex));
        col_seed_kernel<<<static_cast<unsigned>(ceil_div(m, 8)), 256, 0, st>>>(ix->out_i.as<uint64_t>(), q0, m, k,
                                                                              ix->col_buf.as<uint64_t>(), ix->col_cnt.as<uint32_t>(),
                                                                              ix->col_base.as<uint32_t>(), ix->col_thr.as<float>());
        ++g_launches;
        CU_TRY(cudaGetLastError());
    }
    return CVDB_OK;
}

int cvdb_selfjoin_cross(cvdb_index_t h, const void* q, int64_t nq, int dtype, const int32_t* q_ids, const int32_t* group_q,
                        int64_t row_begin, int64_t row_end, int64_t col_row_min, int64_t id_base, uint64_t* keys, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (ix->sj_k < 2) return fail(CVDB_EINVAL, "no self-join is open (cvdb_selfjoin_begin)");
    if (nq < 1 || nq > 65536) return fail(CVDB_ELIMIT, "1..65536 anchors per call");
    if (dtype != CVDB_DTYPE_F32 && dtype != CVDB_DTYPE_BF16 && dtype != CVDB_DTYPE_F16)
        return fail(CVDB_EINVAL, "unknown dtype %d", dtype);
    if (row_end <= 0) row_end = ix->ntotal;
    if (row_begin < 0 || row_begin % 128 || row_begin >= row_end || row_end > ix->ntotal)
        return fail(CVDB_EINVAL, "row range [%lld, %lld) must start at a multiple of 128 inside the index",
                    static_cast<long long>(row_begin), static_cast<long long>(row_end));
    if (id_base < 0 || id_base + ix->ntotal > 0xFFFFFFFELL) return fail(CVDB_ELIMIT, "ids must stay below 2^32 - 1");
    if (!q || !q_ids || !keys) return fail(CVDB_EINVAL, "null pointer");
    if (group_q && !ix->has_groups) return fail(CVDB_EINVAL, "group_q given but the index has no groups");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamOrder order(ix, st);
    SearchExtra ex;
    ex.row_begin = row_begin;
    ex.row_end = row_end;
    ex.col = true;
    ex.col_row_min = std::max(row_begin, col_row_min);
    ex.q_ids = q_ids;
    cvdb_search_opts o{};
    o.id_base = id_base;
    TRY(search_device(ix, q, nq, dtype, ix->sj_k, nullptr, nullptr, nullptr, nullptr, group_q, &o, st, keys, nullptr, &ex));
    return selfjoin_compact(ix, ex.col_row_min, row_end, st);
}

int cvdb_selfjoin_finish(cvdb_index_t h, int64_t row0, int64_t nrows, const uint64_t* row_keys, float* D, int64_t* I,
                         void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (ix->sj_k < 2) return fail(CVDB_EINVAL, "no self-join is open (cvdb_selfjoin_begin)");
    if (row0 < 0 || nrows < 0 || row0 + nrows > ix->ntotal) return fail(CVDB_EINVAL, "rows outside the index");
    if (nrows == 0) return CVDB_OK;
    if (!row_keys || !D || !I) return fail(CVDB_EINVAL, "null pointer");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamOrder order(ix, st);
    selfjoin_finalize_kernel<<<static_cast<unsigned>(ceil_div(nrows, 8)), 256, 0, st>>>(
        row_keys, ix->col_buf.as<uint64_t>(), ix->col_cnt.as<uint32_t>(), row0, nrows, ix->sj_k, D, I);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return CVDB_OK;
}

int cvdb_selfjoin_dirty(cvdb_index_t h, int32_t* rows_out, int64_t max_out, int64_t* n_out, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (ix->sj_k < 2) return fail(CVDB_EINVAL, "no self-join is open (cvdb_selfjoin_begin)");
    if (!n_out || max_out < 0 || (max_out > 0 && !rows_out)) return fail(CVDB_EINVAL, "bad arguments");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamOrder order(ix, st);
    CU_TRY(cudaMemsetAsync(ix->col_scal.p, 0, 8, st));
    collect_flagged_kernel<<<static_cast<unsigned>(ceil_div(ix->ntotal, 256)), 256, 0, st>>>(
        ix->col_dirty.as<uint8_t>(), ix->ntotal, rows_out, max_out, ix->col_scal.as<unsigned long long>());
    ++g_launches;
    unsigned long long c[2] = {0, 0};
    CU_TRY(cudaMemcpyAsync(c, ix->col_scal.p, 16, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    // a log that ran over its capacity lost candidates of unknown rows: every row has to be recomputed
    *n_out = c[1] ? std::max<int64_t>(ix->ntotal, max_out + 1) : static_cast<int64_t>(c[0]);
    return CVDB_OK;
}

int cvdb_selfjoin_end(cvdb_index_t h) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    cvdb_guard g(ix->device);
    cudaDeviceSynchronize();
    ix->sj_k = 0;
    for (DevBuf* b : {&ix->col_thr, &ix->col_cnt, &ix->col_base, &ix->col_buf, &ix->col_dirty, &ix->col_log}) b->release();
    return CVDB_OK;
}

// ------------------------------------------------------------ shard exchange in keys
int cvdb_index_search_keys(cvdb_index_t h, const void* q, int64_t nq, int dtype, int k, uint64_t* keys,
                           const cvdb_search_opts* opts, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (nq < 0) return fail(CVDB_EINVAL, "nq < 0");
    if (k < 1 || k > CVDB_MAX_K) return fail(CVDB_ELIMIT, "k=%d outside [1, %d]", k, CVDB_MAX_K);
    if (dtype != CVDB_DTYPE_F32 && dtype != CVDB_DTYPE_BF16 && dtype != CVDB_DTYPE_F16)
        return fail(CVDB_EINVAL, "unknown dtype %d", dtype);
    if (nq == 0) return CVDB_OK;
    if (!q || !keys) return fail(CVDB_EINVAL, "q and keys must be non-null");
    if (opts && opts->group_q && !ix->has_groups)
        return fail(CVDB_EINVAL, "group_q given but the index has no groups (cvdb_index_set_groups)");
    if (ix->permuted() && opts && (opts->self_ids || opts->group_q))
        return fail(CVDB_EINVAL, "self / group exclusion is not available on an index re-stored list-major");
    const int64_t id_base = opts ? opts->id_base : 0;
    if (id_base < 0 || id_base + ix->ntotal > 0xFFFFFFFELL)
        return fail(CVDB_ELIMIT, "keys carry 32-bit ids: id_base + ntotal must stay below 2^32 - 1");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamOrder order(ix, st);
    const size_t esz = dtype == CVDB_DTYPE_F32 ? 4 : 2;
    TRY(ix->q_norm.ensure(static_cast<size_t>(nq) * 4));  // kept for cvdb_index_merge_keys (L2 needs |q|^2)
    const int64_t chunk = query_chunk(ix, k);
    for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
        const int64_t m = std::min(chunk, nq - q0);
        const void* qsrc = static_cast<const char*>(q) + q0 * ix->d * esz;
        const int32_t* self = (opts && opts->self_ids) ? opts->self_ids + q0 : nullptr;
        const int32_t* grp = (opts && opts->group_q) ? opts->group_q + q0 : nullptr;
        TRY(search_device(ix, qsrc, m, dtype, k, nullptr, nullptr, nullptr, self, grp, opts, st, keys + q0 * k,
                          ix->q_norm.as<float>() + q0));
    }
    return CVDB_OK;
}

int cvdb_merge_keys(const uint64_t* keys, int64_t nq, int nlists, int k_in, int k, uint64_t* keys_out, void* stream) {
    if (nq < 0 || nlists < 1 || k_in < 1 || k < 1 || k > CVDB_MAX_K) return fail(CVDB_EINVAL, "bad sizes");
    if (nq == 0) return CVDB_OK;
    if (!keys || !keys_out) return fail(CVDB_EINVAL, "null pointer");
    ptr_guard g(keys);
    return launch_merge(keys, nq, nlists, k_in, k, 0, nullptr, 0, k_in, static_cast<int64_t>(nq) * k_in, nullptr, 0, nullptr,
                        nullptr, nullptr, keys_out, static_cast<cudaStream_t>(stream));
}

int cvdb_index_merge_keys(cvdb_index_t h, const uint64_t* keys, int64_t nq, int nlists, int k_in, int k, float* D,
                          int64_t* I, void* stream) {
    TRY(check_index(h));
    Index* ix = reinterpret_cast<Index*>(h);
    if (nq < 0 || nlists < 1 || k_in < 1 || k < 1 || k > CVDB_MAX_K) return fail(CVDB_EINVAL, "bad sizes");
    if (nq == 0) return CVDB_OK;
    if (!keys || !D || !I) return fail(CVDB_EINVAL, "null pointer");
    const int l2 = ix->metric == CVDB_METRIC_L2;
    if (l2 && ix->q_norm.cap < static_cast<size_t>(nq) * 4)
        return fail(CVDB_EINVAL, "L2 distances need the query norms of the preceding cvdb_index_search_keys call "
                                 "(same queries, same handle)");
    cvdb_guard g(ix->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamOrder order(ix, st);
    return launch_merge(keys, nq, nlists, k_in, k, l2, ix->q_norm.as<float>(), 0, k_in, static_cast<int64_t>(nq) * k_in,
                        nullptr, 0, D, I, nullptr, nullptr, st);
}

int cvdb_kmeans_accumulate(const void* x, int64_t n, int d, int dtype, const int32_t* assign, float* sums,
                           int32_t* counts, void* stream) {
    if (n < 0 || d < 1) return fail(CVDB_EINVAL, "bad sizes");
    if (n == 0) return CVDB_OK;
    if (!x || !assign || !sums || !counts) return fail(CVDB_EINVAL, "null pointer");
    ptr_guard g(sums);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t blocks = std::min<int64_t>(ceil_div(n, 8), 148 * 16);
    if (dtype == CVDB_DTYPE_F32)
        kmeans_update_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const float*>(x), n, d,
                                                                                     assign, sums, counts);
    else if (dtype == CVDB_DTYPE_BF16)
        kmeans_update_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), n, d, assign, sums, counts);
    else if (dtype == CVDB_DTYPE_F16)
        kmeans_update_kernel<__half><<<static_cast<unsigned>(blocks), 256, 0, st>>>(static_cast<const __half*>(x), n, d,
                                                                                    assign, sums, counts);
    else
        return fail(CVDB_EINVAL, "unknown dtype %d", dtype);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return CVDB_OK;
}

int cvdb_kmeans_finalize(const float* sums, const int32_t* counts, int K, int d, float* centroids, void* stream) {
    if (K < 1 || d < 1) return fail(CVDB_EINVAL, "bad sizes");
    if (!sums || !counts || !centroids) return fail(CVDB_EINVAL, "null pointer");
    ptr_guard g(centroids);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t total = static_cast<int64_t>(K) * d;
    kmeans_finalize_kernel<<<static_cast<unsigned>(ceil_div(total, 256)), 256, 0, st>>>(sums, counts, K, d, centroids);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return CVDB_OK;
}

int cvdb_kmeans_split_empty(float* centroids, int32_t* counts, int K, int d, float eps, int32_t* n_split, void* stream) {
    if (K < 1 || d < 1) return fail(CVDB_EINVAL, "bad sizes");
    if (!centroids || !counts) return fail(CVDB_EINVAL, "null pointer");
    if (!(eps >= 0.f && eps < 1.f)) return fail(CVDB_EINVAL, "eps must be in [0, 1)");
    ptr_guard g(centroids);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    kmeans_split_empty_kernel<<<1, 1024, 0, st>>>(centroids, counts, K, d, eps, n_split);
    ++g_launches;
    CU_TRY(cudaGetLastError());
    return CVDB_OK;
}

}  // extern "C"
