#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
G="python bench.py --workload c3 --steps 40 --warmup 3 --no-e2e --no-cpu-baseline"
$G > $O/r2_plain_c3c.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gibbs_sweep -s 4 -c 1 -o $O/r2_prof_c3c $G > $O/r2_ncu_f_c3c.log 2>&1
if [ -f $O/r2_prof_c3c.ncu-rep ]; then
  ncu -i $O/r2_prof_c3c.ncu-rep --page raw --csv > $O/r2_prof_c3c_raw.csv 2>/dev/null
  ncu -i $O/r2_prof_c3c.ncu-rep --page source --csv > $O/r2_prof_c3c_source.csv 2>/dev/null
  rm -f $O/r2_prof_c3c.ncu-rep
fi
export NBMF_LIB_VARIANT=prof
timeout 300 python tools/role_profile.py 65536 3 64 > $O/r2_roles13_c4_fast.log 2>&1
timeout 300 python tools/role_profile.py 65536 3 64 512 > $O/r2_roles13_c4_fast_f512.log 2>&1
timeout 300 python tools/role_profile.py 65536 3 128 > $O/r2_roles13_64k_fast.log 2>&1
timeout 300 python tools/role_profile.py 65536 3 128 512 > $O/r2_roles13_64k_fast_f512.log 2>&1
echo done
