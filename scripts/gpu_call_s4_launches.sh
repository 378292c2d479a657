#!/bin/bash
# session-4 measurement call: ncu launch lists of the final code (default config and config 3), each after its plain run
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 400 python bench.py --steps 2 --warmup 3 > gpurun_out/s4_bench_short.json 2> gpurun_out/s4_bench_short.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s4_launches.csv \
  python bench.py --steps 2 --warmup 3 > gpurun_out/s4_ncu.log 2>&1
echo "ncu config 2 exit $?"
timeout 400 python bench.py --config 3 --steps 2 --warmup 3 > gpurun_out/s4_bench_short_c3.json 2> gpurun_out/s4_bench_short_c3.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s4_launches_c3.csv \
  python bench.py --config 3 --steps 2 --warmup 3 > gpurun_out/s4_ncu_c3.log 2>&1
echo "ncu config 3 exit $?"
python scripts/summarise_launches.py gpurun_out/s4_launches.csv | tail -30
python scripts/summarise_launches.py gpurun_out/s4_launches_c3.csv | tail -40
