# A/B of environment switches of one build: bash tools/ab_env.sh "IPP_X=0" "IPP_X=1" ...
set -u
run() { env $1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 c2', round(d['value'],1), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['snapshots_per_step'], d['snapshots_rendered_per_step'])"; env $1 python tools/diag_gpu.py luis4 1024 200 30 2>&1 | grep -A1 "^iter 0" | sed "s/^/$1 luis4 /" | cut -c1-330; }
for e in "$@"; do run "$e"; done
