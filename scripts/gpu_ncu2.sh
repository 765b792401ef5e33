#!/bin/bash
# launch list of the bench command + full capture of the hot kernels (CBbunny4): usage gpu_ncu2.sh TAG
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r02b}
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
$BENCH > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
CMD="python scripts/one_render.py dae:CBbunny4 1024 768 21 5"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_connect|k_shadow" -c 3 -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log; cat gpurun_out/${TAG}_plain.log; tail -c 400 gpurun_out/${TAG}_bench_plain.json
