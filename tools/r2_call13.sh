#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
export NBMF_LIB_VARIANT=prof
timeout 300 python tools/role_profile.py 65536 3 64 > $O/r2_roles13_c4_fast.log 2>&1
timeout 300 python tools/role_profile.py 65536 3 64 512 > $O/r2_roles13_c4_fast_f512.log 2>&1
timeout 300 python tools/role_profile.py 65536 3 128 > $O/r2_roles13_64k_fast.log 2>&1
timeout 300 python tools/role_profile.py 65536 3 128 512 > $O/r2_roles13_64k_fast_f512.log 2>&1
echo done
