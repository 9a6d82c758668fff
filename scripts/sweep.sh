#!/bin/bash
# build variants of the library with -D overrides and bench each (development)
set -e
cd "$(dirname "$0")/.."
run() { name=$1; shift; make -s -C fora_b200 clean >/dev/null; make -s -C fora_b200 EXTRA="$*" >/dev/null 2>&1; echo -n "[$name $*] "; ./scripts/bench_quick.sh ${SLOTS:-32}; }
run base
run ub2 -DCFG_PUSH_UB=2
run ub8 -DCFG_PUSH_UB=8 -DCFG_PUSH_WQ=512
run batch2048 -DCFG_PUSH_BATCH=2048
run ua8 -DCFG_PUSH_UA=8
run chunk4096 -DCFG_WALK_CHUNK=4096
run chunk1024 -DCFG_WALK_CHUNK=1024
make -s -C fora_b200 clean >/dev/null; make -s -C fora_b200 >/dev/null 2>&1
FORA_WALK_GRID=16 ./scripts/bench_quick.sh 32 FORA_WALK_GRID=16
SLOTS=48 ./scripts/bench_quick.sh 48
SLOTS=64 ./scripts/bench_quick.sh 64
