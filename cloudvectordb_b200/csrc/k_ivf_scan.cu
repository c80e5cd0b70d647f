// Transposed inverted-list scan: 128 list rows x 16 gathered queries per MMA, 8 stages (ivf_scan.cuh).
#include "ivf_scan.cuh"
#include "k_common.cuh"
namespace cvdb {
cudaError_t launch_ivf_scan(const CUtensorMap& tx128, const CUtensorMap& tx32, const CUtensorMap& tq, const IvfScanParams& p,
                            int grid, cudaStream_t st) {
    static unsigned long long configured = 0;
    constexpr size_t smem = ivf_scan_smem_bytes<16, 8>();
    static_assert(smem <= 232448, "shared memory budget");
    return launch_kernel(ivf_scan_kernel<16, 8>, smem, configured, grid, st, tx128, tx32, tq, p);
}
}  // namespace cvdb
