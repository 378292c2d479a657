#!/bin/bash
# round-2 GPU call A: new parity-mode tests, ADVICE regression tests, the bench line, ncu of the bandwidth-bound kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -q --maxfail=20 -x -k "split or philox" > gpurun_out/a_parity_unit.log 2>&1
echo "parity unit exit $?" >> gpurun_out/a_status.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -q --maxfail=20 -k "not split and not philox" > gpurun_out/a_parity_traj.log 2>&1
echo "parity traj exit $?" >> gpurun_out/a_status.txt
timeout 900 python -m pytest tests/test_gpu_train.py -q --maxfail=20 -k "autoregressive or resume" > gpurun_out/a_train.log 2>&1
echo "train exit $?" >> gpurun_out/a_status.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench exit $?" >> gpurun_out/a_status.txt
timeout 300 python scripts/profile_aux_kernels.py > gpurun_out/a_aux_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:'perturb_|output_loss|pad_rgb4|eb_forward|gc_forward|ssim_level' -c 20 -o gpurun_out/a_aux python scripts/profile_aux_kernels.py > gpurun_out/a_aux_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/a_status.txt
tail -5 gpurun_out/a_parity_unit.log gpurun_out/a_parity_traj.log gpurun_out/a_train.log; cat gpurun_out/a_status.txt; head -c 1500 gpurun_out/a_bench.json
