#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
rm -f gpurun_out/b_status.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -q --maxfail=30 > gpurun_out/b_parity.log 2>&1
echo "parity exit $?" >> gpurun_out/b_status.txt
timeout 600 python scripts/debug_cheng_train.py > gpurun_out/b_debug_train.log 2>&1
echo "debug exit $?" >> gpurun_out/b_status.txt
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 --deselect tests/test_gpu_parity.py > gpurun_out/b_all.log 2>&1
echo "all exit $?" >> gpurun_out/b_status.txt
tail -15 gpurun_out/b_parity.log; cat gpurun_out/b_debug_train.log | tail -20; tail -15 gpurun_out/b_all.log; cat gpurun_out/b_status.txt
