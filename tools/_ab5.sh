set -u
run() { env $2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2 c2', round(d['value'],1), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; }
cp inspection_path_planning_b200/libipp.so /tmp/tree.so
for so in _ab/hdr16.so _ab/dd2.so; do cp $so inspection_path_planning_b200/libipp.so; run $(basename $so .so) IPP_POSE_DEDUPE=0; run $(basename $so .so) IPP_POSE_DEDUPE=1; done
cp /tmp/tree.so inspection_path_planning_b200/libipp.so
