#!/bin/bash
# round-end validation on one box: GPU parity tests, smoke, the default bench line; then (optional) an A/B of library builds.
# usage: gpu_validate.sh TAG [other.so ...]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-v}; shift
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/${TAG}_smoke.log 2>&1; tail -4 gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.log 2>&1; tail -3 gpurun_out/${TAG}_bench.log | cut -c1-1500
if [ $# -gt 0 ]; then timeout 600 bash scripts/gpu_ab_lib.sh ${TAG}_ab "dae:CBbunny4 1024 768 128 5" "$@"; fi
