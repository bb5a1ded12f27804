#!/bin/bash
# "switch parts off" / tuning sweep of the fused E1 -> E4 kernel (csrc/b2b.cuh); run on the GPU box.
#   DBGS="0 31"   MM_B2B_DBG values;   XFLAGS="-DMM_B2B_STG=0|-DMM_B2B_S1=2"   '|'-separated extra flag sets
IFS='|' read -ra SETS <<< "${XFLAGS:- }"
for xf in "${SETS[@]}"; do
for dbg in ${DBGS:-0 31 127}; do
  MEDMOE_BUILD_ONLY=b2b.cu MEDMOE_NVCC_EXTRA="-DMM_B2B_DBG=$dbg $xf" python -m medmoe_b200.build --force > /dev/null 2>&1 || echo build failed
  echo "== MM_B2B_DBG=$dbg $xf"; python tools/b2b_probe.py 2>&1 | tail -3
done
done
MEDMOE_BUILD_ONLY=b2b.cu python -m medmoe_b200.build --force > /dev/null 2>&1
