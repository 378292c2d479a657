set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/s4_bench_v14.json 2> gpurun_out/s4_bench_v14.err; tail -c 400 gpurun_out/s4_bench_v14.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s4_bench_ref.json 2> gpurun_out/s4_bench_ref.err; tail -c 300 gpurun_out/s4_bench_ref.json
timeout 300 python bench.py --config 3 > gpurun_out/s4_bench_config3c.json 2> gpurun_out/s4_bench_config3c.err; tail -c 200 gpurun_out/s4_bench_config3c.json
timeout 300 python scripts/cli_paths.py 100 > gpurun_out/s4_cli_paths.txt 2>&1; cat gpurun_out/s4_cli_paths.txt
