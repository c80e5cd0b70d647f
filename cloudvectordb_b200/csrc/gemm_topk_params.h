// Parameter blocks of the GEMM+top-k kernels (plain data, shared by the kernels and the C ABI).
#pragma once
#include <stdint.h>
#include <vector_types.h>

namespace cvdb {

constexpr int kMaxCombos = 6;
// pooled thresholds (gemm_topk.cuh): levels m = 2, 4, .., 2^kPoolLevels; level m keeps m words at word offset m - 2
#ifndef CVDB_POOL_LEVELS
#define CVDB_POOL_LEVELS 4
#endif
constexpr int kPoolLevels = CVDB_POOL_LEVELS;
constexpr int kPoolSlots = 2 << kPoolLevels;  // 32 words per query for 4 levels (30 used)

struct GemmTopkParams {
    int nq;               // queries in this launch
    int n_rows;           // database rows in this launch
    int k;                // results kept per (query, slice); k <= 32*E - 8 (E>0) or 1 (E==0)
    int room;             // compact a query's buffer once it holds more than this many candidates (<= 32*E - 8)
    int q_tiles;          // ceil(nq / 128)
    int n_tiles;          // ceil(n_rows / BLOCK_N)
    int n_slices;         // database slices
    int tiles_per_slice;  // in BLOCK_N units
    int nkb;              // 64-element K blocks per plane
    int k16;              // 16-element MMA K steps per plane (ceil(Kp / 16)): the last K block may need fewer than 4
    int n_combo;          // (A plane, B plane) pairs accumulated per tile: 1 (bf16) or 6 (exact split)
    int plane_cols;       // columns per plane (Kp)
    uint32_t a_planes;    // 4 bits per combo: plane of the A (query) operand
    uint32_t b_planes;    // 4 bits per combo: plane of the B (database) operand
    const int32_t* self_ids;  // [nq] database row to drop for each query, or null
    const int32_t* group_q;   // [nq] group id per query (<0: none), or null
    const int32_t* group_db;  // [n_rows] group id per database row, or null
    uint64_t* cand;  // [gridDim.x][128][32*E] candidate scratch (E>0)
    uint64_t* part;  // [nq][n_slices][k] per-slice results (keys)
    uint32_t* gthr;  // [nq] shared per-query threshold (ordered-float), zeroed before the launch
    uint32_t* gpool;  // [nq][kPoolSlots] pooled per-slice order statistics (gemm_topk.cuh, "pooled thresholds"), zeroed; or null
    uint32_t* wave_cnt;  // [waves] producers that finished issuing the loads of their item in that wave (or null)
    uint32_t* done;      // [q_tiles][n_slices] epilogue warps that have flushed that item (zeroed; null: no inheritance)
    int done_full;       // warps that flush one item: 4 (single CTA) or 8 (CTA pair)
    int tile0;       // first database tile of the launch (n_tiles stays the END tile; rows before tile0 are not scanned)
    // column direction of a symmetric self-join (gemm_topk.cuh, scan_chunk_col); col_thr == null: off
    const float* col_thr;  // [n_rows] a score must beat it to become a candidate of that database row
    uint4* col_log;                      // append-only log of column candidates {row, 0, key lo, key hi}
    uint32_t* col_log_cnt;               // records reserved so far (in segments of 8)
    uint32_t col_log_cap;                // capacity of the log in records (< 2^31)
    int col_row_min;       // database rows below this do not collect
    const int32_t* q_ids;  // [nq] id of query i in the id space of the column lists (null: self_ids[i])
    int dbg;         // tuning experiments: 1 = skip scan, 2 = skip TMEM read too
    int a_quarter;   // single-CTA kernel with one partial query tile: query rows per epilogue warp (the A tile is
                     // loaded as four 32-row boxes, box j = queries [j*a_quarter, j*a_quarter + 32)); 0 = one box
};

struct GroupItem {
    int a_row0;  // first gathered query row of the item
    int a_rows;  // valid gathered rows (<= 128)
    int x_row0;  // first database row of the list (list-major storage)
    int x_rows;  // rows in the list
};

struct GroupedParams {
    int room;                // compaction trigger, as in GemmTopkParams
    const int* n_items_ptr;  // device scalar: number of work items (written by the bucketing kernels)
    int k;
    int nkb, k16;
    const GroupItem* items;
    const int32_t* pair_query;  // [pairs] query id of each gathered row
    const int32_t* pair_dst;    // [pairs] output slot (query * nprobe + probe) of each gathered row
    const int32_t* row_ids;     // [n_rows] caller-visible id of each stored row (or null: position)
    uint64_t* cand;             // [gridDim.x][128][32*E]
    uint64_t* part;             // [nq * nprobe][k]
    uint32_t* gthr;             // [nq]
};

}  // namespace cvdb
