#!/bin/bash
# A/B only: usage gpu_ab.sh TAG "scene W H spp m" VARIANT...
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=$1; SPEC=$2; shift; shift
mkdir -p gpurun_out
timeout 900 python scripts/perf_ab.py $SPEC "$@" 2>&1 | tee gpurun_out/${TAG}.log
