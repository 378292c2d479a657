#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
rm -f gpurun_out/g_status.txt
timeout 1500 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_parity.py tests/test_gpu_train.py -q --maxfail=30 > gpurun_out/g_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/g_status.txt
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/g_bench_c3.json 2> gpurun_out/g_bench_c3.err
echo "bench c3 exit $?" >> gpurun_out/g_status.txt
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/g_bench_c4.json 2> gpurun_out/g_bench_c4.err
echo "bench c4 exit $?" >> gpurun_out/g_status.txt
timeout 900 python bench.py --config 5 --steps 3 --warmup 1 > gpurun_out/g_bench_c5.json 2> gpurun_out/g_bench_c5.err
echo "bench c5 exit $?" >> gpurun_out/g_status.txt
tail -12 gpurun_out/g_tests.log; cat gpurun_out/g_status.txt
for c in 3 4 5; do tail -2 gpurun_out/g_bench_c$c.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/g_bench_c$c.json"))
    print("config $c:", d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "eager", (d.get("gpu_eager_baseline") or {}).get("value"), d.get("parts_ms"))
except Exception as e:
    print("config $c: no line", e)
PY
done
