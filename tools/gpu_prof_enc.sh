#!/bin/bash
# ncu --set full of the short-row attention kernels (forward + backward) at the encoder shape
TAG=${1:-enc}
mkdir -p gpurun_out
python tools/attn_enc_one.py > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_.*enc -s 2 -c 2 -f -o gpurun_out/prof_$TAG python tools/attn_enc_one.py > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$TAG.log
