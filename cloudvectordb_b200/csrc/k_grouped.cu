// Grouped (inverted-list) kernel: 128 gathered query rows x 128 list rows, 6 stages.
#include "k_common.cuh"
namespace cvdb {
template <int E>
static cudaError_t go(const CUtensorMap& tq, const CUtensorMap& tx, const GroupedParams& p, int grid, cudaStream_t st) {
    static unsigned long long configured = 0;
    return launch_kernel(gemm_topk_grouped_kernel<128, 6, E>, gemm_topk_ss_smem_bytes<128, 6>(), configured, grid, st, tq, tx, p);
}
cudaError_t launch_grouped(int E_, const CUtensorMap& tq, const CUtensorMap& tx, const GroupedParams& p, int grid,
                           cudaStream_t st) {
    CVDB_DISPATCH_E(E_, (go<E>(tq, tx, p, grid, st)))
}
}  // namespace cvdb
