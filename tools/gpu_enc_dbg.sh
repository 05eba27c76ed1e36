#!/bin/bash
# decomposition of the short-row attention time: MAE_ENC_DBG 0 (real), 1 (no softmax arithmetic), 2 (no TMA), 3 (neither)
for d in 0 1 2 3; do echo "dbg=$d"; MAE_ENC_DBG=$d timeout 120 python tools/attn_time.py 2>&1 | grep "L=36"; done
