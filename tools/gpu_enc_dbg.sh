#!/bin/bash
# decomposition of the short-row attention time (MAE_ENC_DBG bits: 1 no softmax arithmetic, 2 no loads after the first
# stages, 4 one chunk per row, 8 contiguous addresses)
for d in ${@:-0 1 2 3}; do echo "dbg=$d"; MAE_ENC_DBG=$d timeout 120 python tools/attn_time.py 2>&1 | grep "L=36"; done
