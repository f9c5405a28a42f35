#!/bin/bash
# final single-GPU call: whole GPU suite, smoke, bench lines of record, short ncu recaptures
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > $O/r2_gpu_tests.log 2>&1
echo "pytest rc=$?" >> $O/r2_gpu_tests.log
timeout 300 python __graft_entry__.py smoke > $O/r2_smoke.log 2>&1
timeout 300 python tests/tools/gibbs_bench.py > $O/r2_gibbs_bench4.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_c5_auto.json 2> $O/r2_bench_c5_auto.err
timeout 600 python bench.py --precision 4 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/r2_bench_c5_lo8.json 2> $O/r2_bench_c5_lo8.err
timeout 600 python bench.py --precision 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2_bench_c5_hilo.json 2> $O/r2_bench_c5_hilo.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_bench_c4_auto.json 2> $O/r2_bench_c4_auto.err
timeout 300 python bench.py --workload c3 --steps 200 --warmup 5 > $O/r2_bench_c3.json 2> $O/r2_bench_c3.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
timeout 600 python tools/smallk_bench.py quick shapes=4,5,6 > $O/r2_smallk_fp64log.log 2>&1
export NBMF_TP_SPIN_LIMIT=2000000000
B="python bench.py --workload c5 --precision 4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$B > $O/r2_plain_c5_lo8b.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 160 --csv --log-file $O/r2_launches_c5_lo8b.csv $B > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:tp_pass -s 6 -c 2 -o $O/r2_prof_c5_lo8b $B > $O/r2_ncu_f_c5_lo8b.log 2>&1
G="python bench.py --workload c3 --steps 40 --warmup 3 --no-e2e --no-cpu-baseline"
$G > $O/r2_plain_c3b.log 2>&1 && timeout 600 ncu --set full --clock-control none -k regex:gibbs_sweep -s 4 -c 2 -o $O/r2_prof_c3b $G > $O/r2_ncu_f_c3b.log 2>&1
for r in r2_prof_c5_lo8b r2_prof_c3b; do
  if [ -f $O/$r.ncu-rep ]; then ncu -i $O/$r.ncu-rep --page raw --csv > $O/${r}_raw.csv 2>/dev/null; rm -f $O/$r.ncu-rep; fi
done
echo done
