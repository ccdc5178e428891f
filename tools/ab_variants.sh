# A/B of several builds of libipp.so (same C ABI) on one GPU box: bash tools/ab_variants.sh _ab/a.so _ab/b.so ...
# Prints the config 2 bench (3 steps) and the luis4 cold evaluation (200 plans x 30 viewpoints, with a hash of the fitness rows) for every
# build, then restores the tree's own build.
set -u
cp inspection_path_planning_b200/libipp.so /tmp/tree.so
run() { python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 c2', round(d['value'],1), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; python tools/diag_gpu.py luis4 1024 200 30 2>&1 | grep -A1 "^iter 0" | sed "s/^/$1 luis4 /" | sed 's/counters.*//' | cut -c1-200; }
for so in "$@"; do cp $so inspection_path_planning_b200/libipp.so; run $(basename $so .so); done
cp /tmp/tree.so inspection_path_planning_b200/libipp.so
