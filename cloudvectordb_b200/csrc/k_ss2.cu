// Variant 3: CTA-pair kernel streaming both operands (256 queries x 256 rows, 6 stages).
#include "k_common.cuh"
namespace cvdb {
template <int E>
static cudaError_t go(const CUtensorMap& tq, const CUtensorMap& tx, const GemmTopkParams& p, int grid, cudaStream_t st) {
    static unsigned long long configured = 0;
    return launch_kernel(gemm_topk_ss2_kernel<256, 6, E>, gemm_topk_ss2_smem_bytes<256, 6>(), configured, grid, st, tq, tx, p);
}
cudaError_t launch_ss2(int E_, const CUtensorMap& tq, const CUtensorMap& tx, const GemmTopkParams& p, int grid,
                       cudaStream_t st) {
    CVDB_DISPATCH_E(E_, (go<E>(tq, tx, p, grid, st)))
}
}  // namespace cvdb
