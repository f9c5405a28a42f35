#!/bin/bash
set -u
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python tools/err_anatomy.py 65536 128 lo8 1024 512 128 > $O/r2_anatomy_64k_lo8.log 2>&1
timeout 600 python tools/err_anatomy.py 65536 128 fast 512 128 > $O/r2_anatomy_64k_fast.log 2>&1
echo done
