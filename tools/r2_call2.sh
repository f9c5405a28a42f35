#!/bin/bash
# second GPU call of round 2: all tests (no -x), C5 error growth, perturbation growth, gamma=1 curves
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -s > $O/r2_tests2.log 2>&1
echo "pytest rc=$?" >> $O/r2_tests2.log
timeout 400 python tools/error_growth.py 65536 128 1,2,5,10,20,40,60,80,100 pert > $O/r2_eg_64k_pert.log 2>&1
timeout 400 python tools/error_growth.py 65536 128 1,2,5,10,20,40,60,80,100 tensor,fast,pert 128 1.0 > $O/r2_eg_64k_gamma1.log 2>&1
timeout 300 python tools/error_growth.py 65536 64 1,2,5,10,20,40,60,80,100 tensor,fast,pert 128 1.0 > $O/r2_eg_c4_gamma1.log 2>&1
timeout 2400 python tools/error_growth.py 262144 128 1,2,5,10,20,25,30,40,50,60,80,100 tensor,fast 128,512 > $O/r2_eg_c5.log 2>&1
echo done
