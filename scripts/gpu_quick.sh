#!/bin/bash
# parity tests + the three A/B scenes: usage gpu_quick.sh TAG [VARIANT ...]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-q}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python scripts/perf_ab.py CBbunny 1024 768 64 5 "$@" 2>&1 | tee gpurun_out/${TAG}_ab_bunny.log
timeout 600 python scripts/perf_ab.py CBgems 1024 768 32 16 "$@" 2>&1 | tee gpurun_out/${TAG}_ab_gems.log
timeout 600 python scripts/perf_ab.py CBspheres 480 360 128 5 "$@" 2>&1 | tee gpurun_out/${TAG}_ab_spheres.log
timeout 600 python scripts/perf_ab.py dae:CBbunny4 1024 768 64 5 2>&1 | tee gpurun_out/${TAG}_ab_bunny4.log
