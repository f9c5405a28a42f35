#!/bin/bash
# records of the round's final kernels: whole GPU suite, smoke, bench lines, launch list, ncu summaries
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $O/r2_gpu_tests_final.log 2>&1
echo "pytest rc=$?" >> $O/r2_gpu_tests_final.log
timeout 300 python __graft_entry__.py smoke > $O/r2_smoke_final.log 2>&1 || python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke_final.log 2>&1
timeout 300 python tests/tools/gibbs_bench.py > $O/r2_gibbs_bench_final.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2f_bench_c5_auto.json 2> $O/r2f_bench_c5_auto.err
timeout 600 python bench.py --precision 4 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/r2f_bench_c5_lo8.json 2> $O/r2f_bench_c5_lo8.err
timeout 600 python bench.py --precision 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2f_bench_c5_hilo.json 2> $O/r2f_bench_c5_hilo.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > $O/r2f_bench_c4_auto.json 2> $O/r2f_bench_c4_auto.err
timeout 300 python bench.py --workload c4 --precision 3 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2f_bench_c4_fast.json 2> $O/r2f_bench_c4_fast.err
timeout 300 python bench.py --workload c3 --steps 200 --warmup 5 > $O/r2f_bench_c3.json 2> $O/r2f_bench_c3.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2f_bench_reference_arm.json 2> $O/r2f_bench_reference_arm.err
export NBMF_TP_SPIN_LIMIT=2000000000
B="python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 160 --csv --log-file $O/r2f_launches_c5.csv $B > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:tp_pass -s 6 -c 2 -o $O/r2f_prof_c5_fast $B > $O/r2f_ncu_c5_fast.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:tp_pass -s 6 -c 2 -o $O/r2f_prof_c5_lo8 $B --precision 4 > $O/r2f_ncu_c5_lo8.log 2>&1
G="python bench.py --workload c3 --steps 40 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none -k regex:gibbs_sweep -s 4 -c 1 -o $O/r2f_prof_c3 $G > $O/r2f_ncu_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 80 --csv --log-file $O/r2f_launches_c3.csv $G > /dev/null 2>&1
for r in r2f_prof_c5_fast r2f_prof_c5_lo8 r2f_prof_c3; do
  if [ -f $O/$r.ncu-rep ]; then ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null; rm -f $O/$r.ncu-rep; fi
done
echo done
