// Resident-query CTA-pair kernel, configuration 2: BLOCK_N=128, 8 K blocks in TMEM, 5 in the shared-memory tail,
// 5 K blocks per stage, 3 stages.
#define TS2_FN launch_ts2_cfg2
#define TS2_BLOCK_N 128
#define TS2_KB_T 8
#define TS2_KB_S 5
#define TS2_KB_STAGE 5
#define TS2_STAGES 3
#include "k_ts2.inc"
