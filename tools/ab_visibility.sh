# A/B of two builds of the package on one GPU box: the tree holds the candidate, _ab/base_pkg/ (python files + libipp.so of the
# build to compare with) the baseline.  Usage: gpurun -- bash tools/ab_visibility.sh
set -u
run() { python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 c2', round(d['value'],1), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; python tools/diag_gpu.py luis4 1024 200 30 2>&1 | grep -A1 "^iter 0" | grep "kernel ms" | sed "s/^/$1 luis4 /"; }
run new
mkdir -p /tmp/newpkg && cp inspection_path_planning_b200/*.py inspection_path_planning_b200/libipp.so /tmp/newpkg/
cp _ab/base_pkg/*.py _ab/base_pkg/libipp.so inspection_path_planning_b200/; run base
cp /tmp/newpkg/*.py /tmp/newpkg/libipp.so inspection_path_planning_b200/; run new
