// Resident-query CTA-pair kernel WITH the column direction of the symmetric self-join (gemm_topk.cuh,
// scan_chunk_col): configurations 0 (K <= 512) and 1 (K <= 768), candidate buffers for k <= 124.
#include "k_common.cuh"
namespace cvdb {
template <int BLOCK_N, int KB_T, int KB_S, int KB_STAGE, int STAGES, int E>
static cudaError_t go(const CUtensorMap& tx, const CUtensorMap& tq, const __nv_bfloat16* q_pack, int q_row_elems,
                      const GemmTopkParams& p, int grid, cudaStream_t st) {
    static unsigned long long configured = 0;
    constexpr size_t smem = gemm_topk_ts2_smem_bytes<BLOCK_N, KB_S, KB_STAGE, STAGES>();
    static_assert(smem <= 232448, "shared memory budget");
    return launch_kernel(gemm_topk_ts2_kernel<BLOCK_N, KB_T, KB_S, KB_STAGE, STAGES, E, true>, smem, configured, grid, st, tx,
                         tq, q_pack, q_row_elems, p);
}
#define CVDB_DISPATCH_E_COL(E_, CALL)                      \
    switch (E_) {                                          \
        case 1: { constexpr int E = 1; return CALL; }      \
        case 2: { constexpr int E = 2; return CALL; }      \
        case 4: { constexpr int E = 4; return CALL; }      \
        case 8: { constexpr int E = 8; return CALL; }      \
        default: return cudaErrorInvalidValue;             \
    }
cudaError_t launch_ts2_col(int cfg, int E_, const CUtensorMap& tx, const CUtensorMap& tq, const __nv_bfloat16* q_pack,
                           int q_row_elems, const GemmTopkParams& p, int grid, cudaStream_t st) {
    if (cfg == 0) { CVDB_DISPATCH_E_COL(E_, (go<128, 8, 0, 4, 6, E>(tx, tq, q_pack, q_row_elems, p, grid, st))) }
    if (cfg == 1) { CVDB_DISPATCH_E_COL(E_, (go<128, 8, 4, 4, 5, E>(tx, tq, q_pack, q_row_elems, p, grid, st))) }
    return cudaErrorInvalidValue;
}
}  // namespace cvdb
