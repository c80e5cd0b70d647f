/* cvdb_b200.h - C ABI of the B200-native exact nearest-neighbour engine.
 *
 * Drop-in boundary for the one data-parallel hot path of a CloudVectorDB-style
 * pipeline: brute-force inner-product / L2 distance over an embedding matrix,
 * fused with per-query top-k.
 *
 * Reference interface each entry point replaces: the reference
 * (dorenwick/CloudVectorDB) ships only README.md; README.md:2 names the stages
 * ("building a very large dataset of triplets", "building the vectordb") whose
 * inner search this library implements.  There is no reference FFI to mirror,
 * so the surface follows the FAISS IndexFlat / Kmeans convention that
 * BASELINE.json's north_star fixes: add() / search(queries, k) -> (D, I).
 *
 * Conventions
 *   - every function returns 0 on success, a negative CVDB_E* code on failure;
 *     cvdb_last_error() returns the message of the calling thread's last failure.
 *   - plain pointers and sizes only.  `on_device` says whether ALL data pointers
 *     of that call (inputs and outputs) are device (1) or host (0) pointers.
 *   - `stream` is a cudaStream_t (NULL = legacy default stream).  With device
 *     pointers a call only enqueues work; with host pointers it returns after
 *     the outputs have been copied back.
 *   - rows are C-contiguous [n, d]; ids are insertion order starting at 0.
 *   - search results: IP sorted by descending score, L2 by ascending SQUARED
 *     distance, ties broken by lower id; missing results have I = -1 and
 *     D = -inf (IP) / +inf (L2).
 *   - one index handle is not thread-safe; distinct handles are.  Calls on one handle may use different
 *     streams: a call first waits (on the device) for the previous call's work when its stream differs,
 *     because the handle's scratch buffers are shared between calls.
 *   - functions without a handle take device pointers and run on the device that owns them, whatever the
 *     caller's current device is.
 *   - there is no CPU fallback: without a B200-class GPU every compute entry
 *     point fails with CVDB_ECUDA.
 */
#ifndef CVDB_B200_H
#define CVDB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVDB_OK 0
#define CVDB_EINVAL (-1)   /* bad argument */
#define CVDB_ECUDA (-2)    /* CUDA runtime / driver failure, or no usable GPU */
#define CVDB_ENOMEM (-3)   /* device allocation failed */
#define CVDB_ELIMIT (-4)   /* argument outside the supported range (k, d) */

#define CVDB_METRIC_IP 0
#define CVDB_METRIC_L2 1

#define CVDB_DTYPE_F32 0
#define CVDB_DTYPE_BF16 1
#define CVDB_DTYPE_F16 2 /* input only: rows and queries are converted to the index storage on the way in */

/* how database rows are kept in HBM */
#define CVDB_STORE_BF16 0  /* one bf16 plane: bf16 x bf16 -> fp32 tensor-core scores */
#define CVDB_STORE_EXACT 1 /* three bf16 planes (hi+mid+lo == the fp32 value): fp32-fidelity scores */

#define CVDB_MAX_K 2048

typedef struct cvdb_index_s* cvdb_index_t;

/* Optional per-search arguments (zero-initialise, then set what is needed).
 * self_ids / group_q follow the `on_device` flag of the search call. */
typedef struct cvdb_search_opts {
    const int32_t* self_ids; /* [nq] database row to exclude for query i (-1: none), or NULL          */
    const int32_t* group_q;  /* [nq] group id of query i (<0: none); rows of the same group
                                (see cvdb_index_set_groups) are excluded, or NULL                     */
    int64_t id_base;         /* added to every returned id (shard offset)                             */
    int profile;             /* 1: bracket the GEMM+top-k kernel with CUDA events (cvdb_index_last_kernel_ms) */
    int force_slices;        /* >0: override the database-slice heuristic (testing)                   */
    int force_variant;       /* 0: auto; 1..4: force that kernel variant (cvdb_index_last_variant)   */
    int debug_flags;         /* kernel-tuning experiments only (results become invalid): 1 = skip the
                                top-k scan, 2 = skip the TMEM read as well, 4 = no threshold sharing between slices,
                                8 = no wave alignment of the producers, 16 = run all K-steps of the padded row width,
                                32 = always sort the whole candidate buffer at the end of a work item,
                                64 = small batches: fill the query tile from row 0 up instead of one quarter per
                                epilogue warp, 128 = every database slice starts from the result list of the latest finished
                                earlier slice of its query tile instead of an empty candidate buffer, 256 = no pooled
                                thresholds across slices, 512 = pooled thresholds also for short work items (4, 8, 16,
                                32, 64, 128, 256 and 512 leave the results valid) */
} cvdb_search_opts;

/* -- index lifetime -------------------------------------------------------
 * FAISS IndexFlatIP(d) / IndexFlatL2(d) analogue.  `device` is the CUDA ordinal. */
int cvdb_index_create(int d, int metric, int storage, int device, cvdb_index_t* out);
int cvdb_index_destroy(cvdb_index_t idx);
int cvdb_index_reset(cvdb_index_t idx);               /* ntotal = 0, keeps the allocation   */
int cvdb_index_reserve(cvdb_index_t idx, int64_t n);  /* capacity for n rows in total       */
int cvdb_index_truncate(cvdb_index_t idx, int64_t n); /* drop the rows added last: ntotal = n <= ntotal; after
                                                         cvdb_index_group_by_list only rows added since can go */
int64_t cvdb_index_ntotal(cvdb_index_t idx);
int cvdb_index_dim(cvdb_index_t idx);

/* FAISS add(x): append n rows (copied; caller keeps ownership of x).  Host rows stream through two pinned
 * 64 MB staging buffers (host copy, PCIe copy and the packing kernel of consecutive chunks overlap); memory the
 * caller pinned itself is read in place.  Returns once x may be reused. */
int cvdb_index_add(cvdb_index_t idx, const void* x, int64_t n, int dtype, int on_device, void* stream);

/* Input validation without a hidden synchronisation on the hot path (SURVEY.md 4.2 "NaN/Inf rejection"): the
 * kernel that packs rows -- database rows in add(), query rows in search()/assign() -- counts the rows whose
 * squared norm is not finite (a NaN or infinite element, or values too large for bf16).  This call waits for
 * `stream` and returns the running total since cvdb_index_create.  A caller that wants to reject such input
 * compares the total before and after a call and undoes an add() with cvdb_index_truncate. */
int cvdb_index_nonfinite_rows(cvdb_index_t idx, int64_t* count_out, void* stream);

/* Group id per database row [ntotal] for positive exclusion in hard-negative
 * mining (README.md:2 "dataset of triplets"); NULL clears. */
int cvdb_index_set_groups(cvdb_index_t idx, const int32_t* group_db, int on_device, void* stream);

/* FAISS search(q, k) -> (D [nq,k] f32, I [nq,k] i64). */
int cvdb_index_search(cvdb_index_t idx, const void* q, int64_t nq, int dtype, int k, float* D, int64_t* I,
                      int on_device, const cvdb_search_opts* opts, void* stream);

/* k = 1 assignment against the index rows: assign [n] int32 (+id_base not applied), dist [n] f32 or NULL.
 * Used as the k-means assignment step with the centroids as index rows. */
int cvdb_index_assign(cvdb_index_t idx, const void* x, int64_t n, int dtype, int32_t* assign, float* dist,
                      int on_device, void* stream);

/* device time of the last profiled GEMM+top-k kernel (blocks until it finished); <0 if none */
float cvdb_index_last_kernel_ms(cvdb_index_t idx);
/* device times (ms) of the profiled GEMM+top-k launches since the previous call, oldest first (at most the
 * last 64); blocks until they finished.  Returns how many were written to out_ms, or a negative error. */
int cvdb_index_profile_ms(cvdb_index_t idx, float* out_ms, int max_n);
/* algorithmic work of the last search: flops = 2*nq*ntotal*d*planes_factor, db bytes streamed once */
int cvdb_index_last_work(cvdb_index_t idx, double* flops, double* db_bytes, int* n_slices, int* grid);
/* kernel variant the last search used: 1 = single-CTA streaming kernel, 2 = CTA-pair kernel with the queries
 * resident on chip (TMEM, plus a shared-memory tail for 512 < K <= 768; 128-column accumulators),
 * 3 = CTA-pair streaming kernel, 4 = CTA-pair kernel with all of K <= 768 in TMEM (64-column accumulators) */
int cvdb_index_last_variant(cvdb_index_t idx);

/* -- inverted lists (IVF) on top of the coarse quantizer -------------------------
 * cvdb_index_group_by_list re-stores the rows list-major.  list_of_id [ntotal] (int32, DEVICE pointer) names the
 * list (0..nlist-1) of every row, indexed by row id (= insertion order).  Rows added after an earlier grouping are
 * picked up by calling it again with the complete vector.  bf16 storage only.  The order of rows inside a list is
 * unspecified; results do not depend on it (candidates carry their row id).  add() un-groups the index: list
 * search is refused until it is grouped again, while cvdb_index_search / cvdb_index_assign keep working on the
 * re-stored rows and keep returning the ids the rows were added under.  What addresses rows by stored position
 * -- cvdb_index_set_groups, self / group exclusion, cvdb_index_export_rows, truncating below the grouped rows --
 * fails with CVDB_EINVAL from the first grouping until cvdb_index_reset.
 * cvdb_index_list_offsets copies the nlist+1 list boundaries (stored positions) to a DEVICE buffer.
 * cvdb_index_search_lists: like cvdb_index_search, but query i only scans the lists probes[i][0..nprobe)
 * (int32, follows on_device; entries < 0 are skipped).  Returned ids are row ids. */
int cvdb_index_group_by_list(cvdb_index_t idx, const int32_t* list_of_id, int nlist, void* stream);
int cvdb_index_list_offsets(cvdb_index_t idx, int32_t* out_device, void* stream);
int cvdb_index_search_lists(cvdb_index_t idx, const void* q, int64_t nq, int dtype, int k, const int32_t* probes,
                            int nprobe, float* D, int64_t* I, int on_device, void* stream);

/* -- triplet assembly from mined neighbours (README.md:2 "dataset of triplets") ---------
 * For anchor i (global id anchor_base + i) with positive pos[i] (< 0: none), take up to per_anchor hard negatives
 * from ranks [skip_top, k) of its mined list (I, D as returned by a search with exclusion), skipping rows on the
 * "probably a false negative" side of `limit` when use_limit (IP: score > limit, L2: distance < limit).
 * out [n][per_anchor][3] int64 = (anchor, positive, negative), unused slots -1.  Device pointers only. */
int cvdb_build_triplets(const int64_t* I, const float* D, int64_t n, int k, const int64_t* pos, int64_t anchor_base,
                        int skip_top, int per_anchor, int metric, float limit, int use_limit, int64_t* out, void* stream);

/* -- persistence: the packed rows as stored in HBM (row_bytes each), to / from HOST memory ------------
 * A file written from export_rows can be re-loaded with import_rows into an index created with the same
 * (d, metric, storage).  Rows keep their order; groups / list grouping are not part of the dump. */
int64_t cvdb_index_row_bytes(cvdb_index_t idx);
int cvdb_index_export_rows(cvdb_index_t idx, int64_t row0, int64_t nrows, void* host_dst, void* stream);
int cvdb_index_import_rows(cvdb_index_t idx, const void* host_src, int64_t nrows, void* stream);

/* -- shard/merge layer ----------------------------------------------------
 * k-way select over `nlists` per-shard result lists, laid out [nlists][nq][k_in]
 * (what an all-gather of per-rank (D, I) produces).  Output [nq][k]. */
int cvdb_merge_topk(const float* Dc, const int64_t* Ic, int64_t nq, int nlists, int k_in, int k, int metric,
                    float* D, int64_t* I, int on_device, void* stream);

/* The same exchange in the engine's own 64-bit candidate keys (8 bytes per candidate instead of 12, one buffer
 * instead of two, and the final merge is a head-pointer merge of sorted lists):
 *   cvdb_index_search_keys: like cvdb_index_search with DEVICE pointers, but writes keys [nq][k] (uint64, sorted
 *     descending, 0 = no result) whose id field already includes opts->id_base (must stay below 2^32 - 1);
 *   cvdb_index_merge_keys: k-way merge of `nlists` such lists laid out [nlists][nq][k_in] (an
 *     all_gather_into_tensor of every rank's keys) -> D [nq][k] f32, I [nq][k] i64, DEVICE pointers.  For L2 it uses
 *     the query norms the preceding cvdb_index_search_keys call on this handle computed (same queries). */
int cvdb_index_search_keys(cvdb_index_t idx, const void* q, int64_t nq, int dtype, int k, uint64_t* keys,
                           const cvdb_search_opts* opts, void* stream);
int cvdb_index_merge_keys(cvdb_index_t idx, const uint64_t* keys, int64_t nq, int nlists, int k_in, int k, float* D,
                          int64_t* I, void* stream);
/* the same merge with the result left in keys [nq][k] (lists of lists: merging what several steps produced) */
int cvdb_merge_keys(const uint64_t* keys, int64_t nq, int nlists, int k_in, int k, uint64_t* keys_out, void* stream);

/* -- symmetric self-join (hard-negative mining over the index's OWN rows; README.md:2 "dataset of triplets") -----
 * A self-join computes every score twice: (anchor i, row j) and (anchor j, row i).  These calls compute each tile of
 * X.X^T once and select in both directions: the anchors of a chunk get their top-k among the rows AT OR AFTER the chunk
 * (row direction, the usual fused epilogue), and every row AFTER the chunk is offered the chunk's anchors as candidates
 * (column direction: per-row threshold / counter / 256-key buffer in HBM, compacted after every chunk).  Walking the
 * chunks in row order covers every ordered pair exactly once with half the flops of the plain join.
 * IP metric, bf16 storage, padded d <= 768, 2 <= k <= 124; self and same-group rows (cvdb_index_set_groups) are
 * excluded in both directions.  2 KB of HBM per index row while the join is open.  All pointers are DEVICE pointers.
 *   begin   opens a join for the current rows of the index
 *   chunk   anchors = rows [row0, row0 + nrows), row0 a multiple of 128, nrows <= 65536, chunks in increasing row order
 *           and contiguous; writes the row-direction result of the chunk as keys [nrows][k] (see cvdb_index_search_keys).
 *           Chunk sizes should start small and at most double (256, 256, 512, 1024, ...): a row's buffer takes the
 *           candidates of ONE chunk under the threshold of the chunks before it.  id_base is added to every id that
 *           leaves the index (row-direction keys, and the anchors' ids in the column lists): the shard offset
 *   seed    warms the column side up with plain searches instead of the first (smallest) chunks: the seed anchors
 *           [0, seed_rows) get their final row-direction result over ALL rows (keys_seed [seed_rows][k], may be NULL
 *           when the caller computes them elsewhere), and every later row starts its column list with its top-k among
 *           the seed anchors.  The chunks then start at row seed_rows with chunk sizes from seed_rows up
 *   cross   the block (anchors of ANOTHER shard) x (rows [row_begin, row_end) of this index; row_end <= 0: ntotal);
 *           only rows >= col_row_min collect in the column direction:
 *           q [nq][d] are the anchors' vectors, q_ids [nq] their global ids, group_q [nq] their groups (or NULL);
 *           row direction -> keys [nq][k] with ids + id_base, column direction -> the lists of this index's rows.
 *           One of the two shards of a pair computes the block; the other receives the keys (see sharded.py)
 *   finish  merges the row-direction keys of rows [row0, row0 + nrows) with their column lists -> D, I (any row range,
 *           after the last chunk)
 *   dirty   rows whose column buffer overflowed (their lists lost candidates): writes up to max_out row ids, returns
 *           the count; the caller recomputes those rows with cvdb_index_search (exact, with the same exclusion)
 *   end     frees the join state */
int cvdb_selfjoin_begin(cvdb_index_t idx, int k, void* stream);
int cvdb_selfjoin_chunk(cvdb_index_t idx, int64_t row0, int64_t nrows, int64_t id_base, uint64_t* keys, void* stream);
int cvdb_selfjoin_seed(cvdb_index_t idx, int64_t seed_rows, int64_t id_base, uint64_t* keys_seed, void* stream);
int cvdb_selfjoin_cross(cvdb_index_t idx, const void* q, int64_t nq, int dtype, const int32_t* q_ids, const int32_t* group_q,
                        int64_t row_begin, int64_t row_end, int64_t col_row_min, int64_t id_base, uint64_t* keys, void* stream);
int cvdb_selfjoin_finish(cvdb_index_t idx, int64_t row0, int64_t nrows, const uint64_t* row_keys, float* D, int64_t* I,
                         void* stream);
int cvdb_selfjoin_dirty(cvdb_index_t idx, int32_t* rows_out, int64_t max_out, int64_t* n_out, void* stream);
int cvdb_selfjoin_end(cvdb_index_t idx);

/* -- k-means update (IVF coarse quantizer) --------------------------------
 * sums [K,d] f32 += x[i] for assign[i]; counts [K] i32 += 1.  The caller zeroes
 * sums/counts and all-reduces them across ranks.  Device pointers only. */
int cvdb_kmeans_accumulate(const void* x, int64_t n, int d, int dtype, const int32_t* assign, float* sums,
                           int32_t* counts, void* stream);
/* centroids[j] = sums[j] / counts[j] where counts[j] > 0 (else unchanged).  Device pointers only. */
int cvdb_kmeans_finalize(const float* sums, const int32_t* counts, int K, int d, float* centroids, void* stream);
/* Re-seed empty clusters after finalize (FAISS Kmeans convention: an empty cluster takes half of a big one).  In
 * ascending order every cluster with counts == 0 takes the currently largest cluster j (ties -> lower id; stops when
 * it has fewer than 2 points): centroid[e] = centroid[j] * (1 +- eps) alternating by dimension, centroid[j] the other
 * way round, counts[e] = counts[j] / 2, counts[j] -= counts[e].  Deterministic, so ranks holding the same all-reduced
 * counts agree.  n_split (device pointer, may be NULL) receives the number of clusters re-seeded.  Device pointers. */
int cvdb_kmeans_split_empty(float* centroids, int32_t* counts, int K, int d, float eps, int32_t* n_split, void* stream);

/* -- host-side planning (no GPU needed; exposed for tests) -----------------------------------------
 * How a search splits the database: work item = (query tile, database slice).  Given the number of query tiles,
 * database tiles and workers (CTAs or CTA pairs), returns the slice count and tiles per slice that minimise
 * waves * (tiles per slice + per-item overhead), with at most max_slices slices. */
int cvdb_plan_slices(int q_tiles, int n_tiles, int workers, int64_t max_slices, int* n_slices, int* tiles_per_slice);

/* -- diagnostics ------------------------------------------------------------ */
const char* cvdb_last_error(void);
int64_t cvdb_kernel_launches(void); /* kernels launched by this library so far (process-wide) */
int cvdb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CVDB_B200_H */
