#!/bin/bash
# ncu captures of the three hot kernels: usage gpu_ncu.sh TAG [scene W H spp m]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r02a}; SCENE=${2:-CBbunny}; W=${3:-1024}; H=${4:-768}; SPP=${5:-21}; M=${6:-5}
mkdir -p gpurun_out
CMD="python scripts/one_render.py $SCENE $W $H $SPP $M"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_connect|k_shadow" -c 3 -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log; cat gpurun_out/${TAG}_plain.log
