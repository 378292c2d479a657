#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
rm -f gpurun_out/f_status.txt
timeout 1200 python -m pytest tests/test_gpu_e2e.py -q --maxfail=30 -k "traced or cheng2020 or quantize_modes or checkpoint_file or mean_scale or unwritten" > gpurun_out/f_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/f_status.txt
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/f_bench_c4.json 2> gpurun_out/f_bench_c4.err
echo "bench c4 exit $?" >> gpurun_out/f_status.txt
tail -25 gpurun_out/f_tests.log; cat gpurun_out/f_status.txt; tail -5 gpurun_out/f_bench_c4.err; head -c 700 gpurun_out/f_bench_c4.json
