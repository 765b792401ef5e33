#!/bin/bash
# first GPU session of round 2: parity tests, connect A/B, a short bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
timeout 600 python scripts/perf_ab.py CBbunny 1024 768 64 5 BDPT_CONNECT_MINB=5 BDPT_CONNECT_MINB=6 BDPT_CONNECT_MINB=8 BDPT_CONNECT_MINB=3 > gpurun_out/s1_ab_bunny.log 2>&1
cat gpurun_out/s1_ab_bunny.log
timeout 600 python scripts/perf_ab.py CBgems 1024 768 32 16 BDPT_CONNECT_MINB=4 BDPT_CONNECT_MINB=6 BDPT_CONNECT_MINB=8 > gpurun_out/s1_ab_gems.log 2>&1
cat gpurun_out/s1_ab_gems.log
timeout 600 python scripts/perf_ab.py CBspheres 480 360 128 5 BDPT_CONNECT_MINB=4 BDPT_CONNECT_MINB=6 > gpurun_out/s1_ab_spheres.log 2>&1
cat gpurun_out/s1_ab_spheres.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
cat gpurun_out/s1_bench.json | head -c 3000
