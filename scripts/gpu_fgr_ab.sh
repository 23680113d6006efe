#!/bin/bash
# FGR front end A/B: LIBS="new ab/libx.so ..." -> tests (new only), batch timing and launch list per library
bash scripts/gpu_fgr_batch.sh 2>&1 | head -${HEADN:-14}
for L in $LIBS; do
  [ "$L" = "new" ] && continue
  echo "== $L"
  MGICP_LIB=$PWD/$L timeout 300 python /tmp/fgr_batch.py 3 | tail -1
  MGICP_LIB=$PWD/$L timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fgr_ab.csv python /tmp/fgr_batch.py 1 > /dev/null 2>&1
  grep -E "k_fgr_match_tc|k_fgr_match_fb\(" gpurun_out/fgr_ab.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | cut -c1-120
done
