#!/bin/bash
# usage: tools/gpurun_once.sh <timeout_s> <gpus> '<command>' -- one attempt, no retry (a busy multi-GPU attempt blocks this repo's GPU access for ~20 min)
T=$1; G=$2; CMD=$3
if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD"; fi
