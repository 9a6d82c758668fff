#!/bin/bash
# BASELINE config 3: ./fora build --opt on the Pokec-shape graph with the index sharded over 1/2/4/8 GPUs, then --with_idx queries.
# usage: scripts/r2_pokec_build_curve.sh <workdir> [gpu counts ...]
set -e
W=${1:-/tmp/fora_pokec}; shift || true
GS=${@:-1 2 4 8}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $W/data/pokec $W/result
python - <<PY
import sys, time, numpy as np, pandas as pd
sys.path.insert(0, "$ROOT")
import fora_b200 as fb
n, m = 1632803, 30622564
t = time.time()
src, dst = fb.synth_edges(n, m, 42)
pd.DataFrame({"s": src, "d": dst}).to_csv("$W/data/pokec/graph.txt", sep=" ", header=False, index=False)
open("$W/data/pokec/attribute.txt", "w").write("n=%d\nm=%d\n" % (n, m))
np.savetxt("$W/data/pokec/ssquery.txt", np.random.default_rng(43).integers(0, n, 1000), fmt="%d")
print("dataset written in %.1fs" % (time.time() - t))
PY
for g in $GS; do
  echo "=== ./fora build --opt --gpus $g"
  rm -f $W/data/pokec/randwalks.*
  $ROOT/fora_b200/fora build --prefix $W/data/ --dataset pokec --epsilon 0.5 --opt --gpus $g --seed 7 --result_dir $W/result 2>&1 | grep -E "index walks|GPU [0-9]+:|tuned_index_size|Memory"
done
echo "=== ./fora query --with_idx --opt --gpus 1 (index of the last build)"
$ROOT/fora_b200/fora query --algo fora --prefix $W/data/ --dataset pokec --epsilon 0.5 --opt --with_idx --query_size 200 --gpus 1 --seed 7 --result_dir $W/result 2>&1 | grep -E "Average|average|idx|Total" | head
