// Dispatch over the configurations of the resident-query CTA-pair kernel (variants 2 and 4).
#include "launchers.h"
namespace cvdb {
#define DECL(n) cudaError_t launch_ts2_cfg##n(int, const CUtensorMap&, const CUtensorMap&, const __nv_bfloat16*, int, \
                                              const GemmTopkParams&, int, cudaStream_t);
DECL(0) DECL(1) DECL(2) DECL(3)
#undef DECL
cudaError_t launch_ts2(int cfg, int E, const CUtensorMap& tx, const CUtensorMap& tq, const __nv_bfloat16* q_pack,
                       int q_row_elems, const GemmTopkParams& p, int grid, cudaStream_t st) {
    switch (cfg) {
        case 0: return launch_ts2_cfg0(E, tx, tq, q_pack, q_row_elems, p, grid, st);
        case 1: return launch_ts2_cfg1(E, tx, tq, q_pack, q_row_elems, p, grid, st);
        case 2: return launch_ts2_cfg2(E, tx, tq, q_pack, q_row_elems, p, grid, st);
        case 3: return launch_ts2_cfg3(E, tx, tq, q_pack, q_row_elems, p, grid, st);
        default: return cudaErrorInvalidValue;
    }
}
}  // namespace cvdb
