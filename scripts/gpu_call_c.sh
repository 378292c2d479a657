#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
rm -f gpurun_out/c_status.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -q --maxfail=30 -k "persistent or streaming or deterministic or fused or active" > gpurun_out/c_kernels.log 2>&1
echo "kernels exit $?" >> gpurun_out/c_status.txt
if grep -q "passed" gpurun_out/c_kernels.log && ! grep -q "failed" gpurun_out/c_kernels.log; then
  ICADV_TC_STREAM_BWD=0 timeout 600 python scripts/launch_table.py 64 gpurun_out/c_table_stream0.json > gpurun_out/c_table_stream0.log 2>&1
  echo "table0 exit $?" >> gpurun_out/c_status.txt
  ICADV_TC_STREAM_BWD=1 timeout 600 python scripts/launch_table.py 64 gpurun_out/c_table_stream1.json > gpurun_out/c_table_stream1.log 2>&1
  echo "table1 exit $?" >> gpurun_out/c_status.txt
  ICADV_TC_STREAM_BWD=1 timeout 600 python scripts/launch_table.py 8 gpurun_out/c_table_stream1_n8.json > gpurun_out/c_table_stream1_n8.log 2>&1
fi
timeout 1500 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_train.py -q --maxfail=30 -k "graph or msssim or autoregressive or trajectory or context_q4" > gpurun_out/c_parity.log 2>&1
echo "parity+train exit $?" >> gpurun_out/c_status.txt
tail -12 gpurun_out/c_kernels.log; grep -E "bwd|step_ms" gpurun_out/c_table_stream0.log gpurun_out/c_table_stream1.log; tail -8 gpurun_out/c_parity.log; cat gpurun_out/c_status.txt
