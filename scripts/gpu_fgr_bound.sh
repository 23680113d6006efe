#!/bin/bash
# what bounds k_fgr_match_tc: the kernel's time with the epilogue / the MMAs compiled out (results invalid, time only)
bash scripts/gpu_fgr_prof.sh >/dev/null 2>&1     # writes /tmp/fgr_one.py
for L in $LIBS; do
  MGICP_LIB=$PWD/$L timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fgr_bound.csv python /tmp/fgr_one.py > /dev/null 2>&1
  echo "== $L"; grep -E "k_fgr_match_tc" gpurun_out/fgr_bound.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | cut -c1-100
done
