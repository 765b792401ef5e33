#!/bin/bash
# A/B of BUILDS of the library on one box: usage gpu_ab_lib.sh TAG "scene W H spp m" other.so...   (ab/ is scratch)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=$1; SPEC=$2; shift; shift
LIB=bidirectional-pathtracing_b200/libbdpt_b200.so
mkdir -p gpurun_out; cp $LIB /tmp/lib_new.so
for rep in 1 2; do
  cp /tmp/lib_new.so $LIB; echo "shipped:";   timeout 600 python scripts/perf_ab.py $SPEC 2>&1 | tail -1
  for o in "$@"; do cp $o $LIB; echo "$o:"; timeout 600 python scripts/perf_ab.py $SPEC 2>&1 | tail -1; done
done 2>&1 | tee gpurun_out/${TAG}.log
cp /tmp/lib_new.so $LIB
