#!/bin/bash
set -e
cd "$(dirname "$0")/.."
run() { name=$1; shift; make -s -C fora_b200 clean >/dev/null; make -s -C fora_b200 EXTRA="$*" >/dev/null 2>&1; echo -n "[$name $*] "; ./scripts/bench_quick.sh ${SLOTS:-32}; }
run c512 -DCFG_WALK_CHUNK=512
run c256 -DCFG_WALK_CHUNK=256
run c512ub2 -DCFG_WALK_CHUNK=512 -DCFG_PUSH_UB=2
run c512ub2b512 -DCFG_WALK_CHUNK=512 -DCFG_PUSH_UB=2 -DCFG_PUSH_BATCH=512
make -s -C fora_b200 clean >/dev/null; make -s -C fora_b200 >/dev/null 2>&1
