# A/B of two builds of libipp.so on one GPU box: the tree holds the candidate, _ab/base/libipp.so the build to compare with.
# Usage: mkdir -p _ab/base && cp inspection_path_planning_b200/libipp.so _ab/base/ && <edit, rebuild> && gpurun -- bash tools/ab_visibility.sh
set -u
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), d['kernel_ms_per_step'])"; python tools/diag_gpu.py luis4 1024 200 30 2>&1 | grep "^iter 0" | cut -c1-60; }
run new
for v in base; do cp _ab/$v/libipp.so inspection_path_planning_b200/libipp.so; run $v; done
