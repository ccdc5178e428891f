nvidia-smi -L; nvidia-smi topo -m 2>/dev/null | head -8
one() { CUDA_VISIBLE_DEVICES=$1 python bench.py --steps 3 --warmup 3 --no-c3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2 gpu$1', round(d['value'],1), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; }
one 0 solo
one 1 solo
one 0 pair & one 1 pair & wait
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-c3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('torchrun N=2', round(d['value'],1), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
timeout -s KILL 280 python -m pytest tests/test_multi_gpu.py -q -x 2>&1 | tail -5
