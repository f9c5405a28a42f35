#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
G="python bench.py --workload c3 --steps 40 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gibbs_sweep -s 4 -c 1 -o $O/r2_prof_c3e $G > $O/r2_ncu_f_c3e.log 2>&1
if [ -f $O/r2_prof_c3e.ncu-rep ]; then
  ncu -i $O/r2_prof_c3e.ncu-rep --page raw --csv > $O/r2_prof_c3e_raw.csv 2>/dev/null
  ncu -i $O/r2_prof_c3e.ncu-rep --page source --csv > $O/r2_prof_c3e_source.csv 2>/dev/null
  rm -f $O/r2_prof_c3e.ncu-rep
fi
echo done
