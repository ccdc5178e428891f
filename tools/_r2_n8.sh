set -u
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 900 $TR --master-port 29541 tools/run_sharded.py ico8 1000,10000,100000,1000000 50 > gpurun_out/r2_c5_sweep_n$N.jsonl 2> gpurun_out/r2_c5_sweep_n$N.err; echo "c5 rc $?"
if [ "${2:-}" = "bench" ]; then
timeout -s KILL 900 $TR --master-port 29542 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc $?"
fi
cut -c1-400 gpurun_out/r2_c5_sweep_n$N.jsonl
