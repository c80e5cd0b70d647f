// Resident-query CTA-pair kernel, configuration 1: BLOCK_N=128, 8 K blocks in TMEM, 4 in the shared-memory tail,
// 4 K blocks per stage, 5 stages.
#define TS2_FN launch_ts2_cfg1
#define TS2_BLOCK_N 128
#define TS2_KB_T 8
#define TS2_KB_S 4
#define TS2_KB_STAGE 4
#define TS2_STAGES 5
#include "k_ts2.inc"
