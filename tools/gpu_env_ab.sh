#!/bin/bash
# whole-step A/B of environment switches: prints ms/step of `bench.py --steps 10 --warmup 4 --no-cpu` for each setting
run() { env "$@" timeout 120 python bench.py --steps 10 --warmup 4 --no-cpu 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],3))"; }
run X=0
for s in "$@"; do run $s; done
run X=0
