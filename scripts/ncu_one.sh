#!/bin/bash
# One full ncu capture of the kernels matching $1 in the command that follows; report lands in
# gpurun_out/$2.ncu-rep.  usage: scripts/ncu_one.sh <kernel-regex> <name> <skip> <count> <command...>
mkdir -p gpurun_out
K=$1; NAME=$2; SKIP=$3; COUNT=$4; shift 4
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $COUNT -o gpurun_out/$NAME -f "$@" > gpurun_out/$NAME.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/$NAME.ncu-rep
