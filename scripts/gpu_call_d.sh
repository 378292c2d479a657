#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
rm -f gpurun_out/d_status.txt
ICADV_TC_STREAM_BWD=2 timeout 900 python -m pytest tests/test_gpu_kernels.py -q --maxfail=30 -k "persistent or streaming or deterministic or fused or active or gdn" > gpurun_out/d_kernels.log 2>&1
echo "kernels exit $?" >> gpurun_out/d_status.txt
timeout 300 python scripts/persistent_timeline.py 16 > gpurun_out/d_timeline.log 2>&1
ICADV_TC_STREAM_BWD=0 timeout 600 python scripts/launch_table.py 64 gpurun_out/d_table_stream0.json > gpurun_out/d_table_stream0.log 2>&1
ICADV_TC_STREAM_BWD=1 timeout 600 python scripts/launch_table.py 64 gpurun_out/d_table_stream1.json > gpurun_out/d_table_stream1.log 2>&1
tail -5 gpurun_out/d_kernels.log; cat gpurun_out/d_timeline.log; grep -E "bwd|step_ms" gpurun_out/d_table_stream0.log gpurun_out/d_table_stream1.log; cat gpurun_out/d_status.txt
