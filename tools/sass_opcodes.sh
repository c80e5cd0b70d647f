#!/bin/bash
# Opcode histogram of every kernel family (cuobjdump -sass of the built objects) -> profiles/sass_opcodes_<family>.txt
# The mnemonics that prove the Blackwell path: UTCHMMA / UTCQMMA (tcgen05.mma), UTMALDG (TMA load), LDTM / STTM
# (tcgen05.ld / st), UTCBAR (tcgen05.commit), UTCATOMSWS (TMEM alloc), SYNCS (mbarrier); HMMA would be mma.sync.
set -e
cd "$(dirname "$0")/.."
B=cloudvectordb_b200/csrc/build
for f in k_ss1 k_ss2 k_grouped k_ivf_scan k_ts2 k_ts2_cfg0 k_ts2_cfg1 k_ts2_cfg2 k_ts2_cfg3 k_ts2_col cvdb_api; do
  [ -f $B/$f.o ] || continue
  out=profiles/sass_opcodes_$f.txt
  {
    echo "# cuobjdump -sass $B/$f.o : opcode histogram over all kernels of the translation unit (sm_100a)"
    echo "# kernels:"
    cuobjdump -sass $B/$f.o | grep -E "^\s*Function :" | sed 's/^\s*Function : /#   /' | c++filt | cut -c1-200
    echo "# tensor/TMA/TMEM opcodes:"
    cuobjdump -sass $B/$f.o | grep -E "^\s+/\*[0-9a-f]{4,}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/;$//' \
      | sort | uniq -c | sort -rn | awk '$2 ~ /^(UTC|UTMA|LDTM|STTM|HMMA|SYNCS|UBLKCP|REDUX|BAR|ATOMS|ATOMG|RED)/ {printf "%8d %s\n", $1, $2}'
    echo "# all opcodes (top 40):"
    cuobjdump -sass $B/$f.o | grep -E "^\s+/\*[0-9a-f]{4,}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/;$//' \
      | sort | uniq -c | sort -rn | head -40 | awk '{printf "%8d %s\n", $1, $2}'
  } > $out
  echo "wrote $out"
done
