#!/bin/bash
# One `ncu --set full` capture per kernel of the hot path (run under gpurun, 1 GPU), plus the launch list.
# Usage: tools/ncu_capture.sh <tag>      -> gpurun_out/<tag>_<kernel>.ncu-rep, gpurun_out/<tag>_launches.csv
set -u
TAG=${1:-r02}
N=${2:-20000}
OUT=gpurun_out
python tools/ncu_fit.py $N 2500 > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
cat $OUT/${TAG}_plain.log
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $OUT/${TAG}_$1 python tools/ncu_fit.py $N 2500 > $OUT/${TAG}_$1.log 2>&1
  echo "$1: rc=$?"
}
cap kbuild kbuild_lower_kernel 0
cap kcross kcross_kernel 0
cap varpartial var_partial_kernel 0
cap bwd bwd_step_kernel 30
cap fwd fwd_step_kernel 30
cap gemm gemm_dmma_kernel 12      # launch 12 = first far update (lookahead off, outer panel 6 tiles at N = 20000)
cap rowwise gemm_dmma_kernel 660  # a launch inside the variance solve
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv python tools/ncu_fit.py $N 2500 > $OUT/${TAG}_launchlist.log 2>&1
echo "launch list rc=$?"
