#!/bin/bash
# per-layer times under the precision plans, with the epilogue knock-out switches of WSU_DBG (timing experiments only)
mkdir -p gpurun_out
for d in ${DBGS:-0 4 2 3}; do
  WSU_DBG=$d python tools/precision_profile.py 32 3 > gpurun_out/r02_profile_dbg$d.log 2>&1
done
