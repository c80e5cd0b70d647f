// Host-callable launchers of the GEMM+top-k kernel families.  Each family is
// compiled in its own translation unit (k_*.cu) so the build parallelises; the
// C ABI (cvdb_api.cu) only sees these declarations.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "gemm_topk_params.h"

namespace cvdb {

// E = candidate-buffer registers per lane (0: top-1, else 1/2/4/8/16).  All return a cudaError_t.
cudaError_t launch_ss1(int E, const CUtensorMap& tq, const CUtensorMap& tx, const GemmTopkParams& p, int grid,
                       cudaStream_t st);
cudaError_t launch_ss2(int E, const CUtensorMap& tq, const CUtensorMap& tx, const GemmTopkParams& p, int grid,
                       cudaStream_t st);
// cfg: 0 = K <= 512 all in TMEM (N=128), 1 = K <= 768 TMEM + 4-block tail, 2 = K <= 832 TMEM + 5-block tail,
//      3 = K <= 768 all in TMEM with N=64 accumulators
cudaError_t launch_ts2(int cfg, int E, const CUtensorMap& tx, const CUtensorMap& tq, const __nv_bfloat16* q_pack,
                       int q_row_elems, const GemmTopkParams& p, int grid, cudaStream_t st);
// the same kernel with the column direction of the symmetric self-join (cfg 0 or 1, E in {1, 2, 4, 8})
cudaError_t launch_ts2_col(int cfg, int E, const CUtensorMap& tx, const CUtensorMap& tq, const __nv_bfloat16* q_pack,
                           int q_row_elems, const GemmTopkParams& p, int grid, cudaStream_t st);
cudaError_t launch_grouped(int E, const CUtensorMap& tq, const CUtensorMap& tx, const GroupedParams& p, int grid,
                           cudaStream_t st);
// transposed inverted-list scan (ivf_scan.cuh): list rows on the M side, <= 16 gathered queries per item, k <= 128
struct IvfScanParams;
cudaError_t launch_ivf_scan(const CUtensorMap& tx128, const CUtensorMap& tx32, const CUtensorMap& tq, const IvfScanParams& p,
                            int grid, cudaStream_t st);

}  // namespace cvdb
