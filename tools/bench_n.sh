# bench.py on N GPUs of one box under torch.distributed.run: bash tools/bench_n.sh <N>   (gpurun --gpus N)
set -u
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 900 $TR --master-port 29542 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err; echo "bench rc $?"
python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench_n$N.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['c3']['cold']['s'], d['c3']['warm']['plans_per_s'], d['c3']['nsga2']['evals_per_s_incl_python'], d['c3']['nsga2_reference_defaults']['wall_s'])"
