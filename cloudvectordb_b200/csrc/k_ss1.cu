// Variant 1: single-CTA kernel streaming both operands (128 queries x 256 rows, 4 stages).
#include "k_common.cuh"
namespace cvdb {
template <int E>
static cudaError_t go(const CUtensorMap& tq, const CUtensorMap& tx, const GemmTopkParams& p, int grid, cudaStream_t st) {
    static unsigned long long configured = 0;
    return launch_kernel(gemm_topk_ss_kernel<256, 4, E>, gemm_topk_ss_smem_bytes<256, 4>(), configured, grid, st, tq, tx, p);
}
cudaError_t launch_ss1(int E_, const CUtensorMap& tq, const CUtensorMap& tx, const GemmTopkParams& p, int grid,
                       cudaStream_t st) {
    CVDB_DISPATCH_E(E_, (go<E>(tq, tx, p, grid, st)))
}
}  // namespace cvdb
