#!/bin/bash
# profiling call (1 GPU): launch lists with DRAM bytes, ncu --set full of the top kernels, small-K table
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
export NBMF_TP_SPIN_LIMIT=2000000000
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > $O/r2_tests6.log 2>&1
timeout 600 python tools/smallk_bench.py > $O/r2_smallk.log 2>&1
timeout 300 python tests/tools/gibbs_bench.py > $O/r2_gibbs_bench2.log 2>&1
B="python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$B > $O/r2_plain_c5.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 160 --csv --log-file $O/r2_launches_c5.csv $B > $O/r2_ncu_l_c5.log 2>&1
$B --precision 4 > $O/r2_plain_c5_lo8.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 160 --csv --log-file $O/r2_launches_c5_lo8.csv $B --precision 4 > $O/r2_ncu_l_c5_lo8.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tp_pass -s 6 -c 2 -o $O/r2_prof_c5_fast $B > $O/r2_ncu_f_c5.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tp_pass -s 6 -c 2 -o $O/r2_prof_c5_lo8 $B --precision 4 > $O/r2_ncu_f_c5_lo8.log 2>&1
G="python bench.py --workload c3 --steps 40 --warmup 3 --no-e2e --no-cpu-baseline"
$G > $O/r2_plain_c3.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gibbs_ -s 8 -c 4 -o $O/r2_prof_c3 $G > $O/r2_ncu_f_c3.log 2>&1
S="python tools/smallk_bench.py quick"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass_mma64\|pass_simt -s 12 -c 60 -o $O/r2_prof_smallk $S > $O/r2_ncu_f_smallk.log 2>&1
# reports stay on the box: only their raw / source pages come back (gpurun_out is capped at 64 MiB)
for r in r2_prof_c5_fast r2_prof_c5_lo8 r2_prof_c3 r2_prof_smallk; do
  if [ -f $O/$r.ncu-rep ]; then
    ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null
    ncu -i $O/$r.ncu-rep --page source --csv > $O/${r}_source.csv 2>/dev/null
    rm -f $O/$r.ncu-rep
  fi
done
du -sh $O > $O/r2_du.log
echo done
