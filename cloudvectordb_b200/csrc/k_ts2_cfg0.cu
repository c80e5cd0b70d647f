// Resident-query CTA-pair kernel, configuration 0: BLOCK_N=128, 8 K blocks in TMEM, 0 in the shared-memory tail,
// 4 K blocks per stage, 6 stages.
#define TS2_FN launch_ts2_cfg0
#define TS2_BLOCK_N 128
#define TS2_KB_T 8
#define TS2_KB_S 0
#define TS2_KB_STAGE 4
#define TS2_STAGES 6
#include "k_ts2.inc"
