#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
rm -f gpurun_out/e_status.txt
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/e_all.log 2>&1
echo "all exit $?" >> gpurun_out/e_status.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
echo "bench exit $?" >> gpurun_out/e_status.txt
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/e_bench_c3.json 2> gpurun_out/e_bench_c3.err
echo "bench c3 exit $?" >> gpurun_out/e_status.txt
timeout 600 python bench.py --config 1 --steps 100 --warmup 3 > gpurun_out/e_bench_c1.json 2> gpurun_out/e_bench_c1.err
echo "bench c1 exit $?" >> gpurun_out/e_status.txt
tail -8 gpurun_out/e_all.log; cat gpurun_out/e_status.txt; head -c 600 gpurun_out/e_bench.json; echo; tail -3 gpurun_out/e_bench_c3.err; head -c 400 gpurun_out/e_bench_c3.json; echo; tail -3 gpurun_out/e_bench_c1.err; head -c 400 gpurun_out/e_bench_c1.json
