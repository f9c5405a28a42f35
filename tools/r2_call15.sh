#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_pipeline.py -m gpu -q -x --timeout=600 > $O/r2_tests15.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests15.log
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c15_c4_auto.json 2> $O/r2_c15_c4_auto.err
timeout 300 python bench.py --workload c4 --precision 3 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_c15_c4_fast.json 2> $O/r2_c15_c4_fast.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_c15_c5_auto.json 2> $O/r2_c15_c5_auto.err
timeout 600 python bench.py --precision 4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_c15_c5_lo8.json 2> $O/r2_c15_c5_lo8.err
G="python bench.py --workload c3 --steps 40 --warmup 3 --no-e2e --no-cpu-baseline"
$G > $O/r2_plain_c3d.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gibbs_ -s 8 -c 2 -o $O/r2_prof_c3d $G > $O/r2_ncu_f_c3d.log 2>&1
if [ -f $O/r2_prof_c3d.ncu-rep ]; then
  ncu -i $O/r2_prof_c3d.ncu-rep --page raw --csv > $O/r2_prof_c3d_raw.csv 2>/dev/null
  ncu -i $O/r2_prof_c3d.ncu-rep --page source --csv > $O/r2_prof_c3d_source.csv 2>/dev/null
  rm -f $O/r2_prof_c3d.ncu-rep
fi
echo done
