#!/bin/bash
# validation (tests, smoke, bench line), knob A/B, then the ncu launch list + full capture of the shipped build.  usage: gpu_round_end.sh TAG
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r02h}
bash scripts/gpu_validate.sh ${TAG}
timeout 300 python scripts/perf_ab.py dae:CBbunny4 1024 768 128 5 BDPT_CACHE_CONFIG=1 BDPT_CACHE_CONFIG=3 BDPT_CACHE_CONFIG=2 2>&1 | tee gpurun_out/${TAG}_knobs.log
timeout 900 bash scripts/gpu_ncu2.sh ${TAG}
