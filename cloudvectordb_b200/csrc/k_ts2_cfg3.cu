// Resident-query CTA-pair kernel, configuration 3: BLOCK_N=64, 12 K blocks in TMEM, 0 in the shared-memory tail,
// 12 K blocks per stage, 4 stages.
#define TS2_FN launch_ts2_cfg3
#define TS2_BLOCK_N 64
#define TS2_KB_T 12
#define TS2_KB_S 0
#define TS2_KB_STAGE 12
#define TS2_STAGES 4
#include "k_ts2.inc"
