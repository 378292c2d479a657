#!/bin/bash
# session-3 measurement call: default bench line, ncu launch list of the same command, the other BASELINE configs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
echo "bench exit $?"
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/h_bench_short.json 2> gpurun_out/h_bench_short.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/h_launches.csv \
  python bench.py --steps 2 --warmup 3 > gpurun_out/h_ncu.log 2>&1
echo "ncu exit $?"
for c in 1 3 4; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/h_bench_c$c.json 2> gpurun_out/h_bench_c$c.err
  echo "bench c$c exit $?"
done
timeout 900 python bench.py --config 5 --steps 3 --warmup 1 > gpurun_out/h_bench_c5.json 2> gpurun_out/h_bench_c5.err
echo "bench c5 exit $?"
python - <<PY
import json
for n in ["h_bench", "h_bench_c1", "h_bench_c3", "h_bench_c4", "h_bench_c5"]:
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "eager", (d.get("gpu_eager_baseline") or {}).get("value"), "roofline", d["roofline"].get("kernel"), d["roofline"].get("frac"), d.get("clocks"))
    except Exception as e:
        print(n, "no line", e)
PY
