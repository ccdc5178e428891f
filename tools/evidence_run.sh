# The evidence run of a round on one B200 (under gpurun): GPU tests, the bench line, the reference arm, the launch list of the bench
# command, --set full captures of K4 + cluster tables, config 1 on GPU and CPU.  Outputs go to gpurun_out/; profiles/README.md says which are kept.
set -u
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2f_pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "reference arm rc $?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c3 > gpurun_out/r2f_plain_ll.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c3 > gpurun_out/r2f_ncu_ll.log 2>&1; echo "launch list rc $?"
python tools/diag_gpu.py mockup_only 1024 1000 20 > gpurun_out/r2f_diag_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_visibility_bins|k_leaf_table" -c 2 -o gpurun_out/r2f_bins python tools/diag_gpu.py mockup_only 1024 1000 20 > gpurun_out/r2f_diag_ncu.log 2>&1; echo "ncu K4 rc $?"
python tools/run_configs.py c1 > gpurun_out/r2f_c1.jsonl 2>&1; python tools/run_configs.py c1cpu > gpurun_out/r2f_c1_cpu.jsonl 2>&1; echo "c1 rc $?"
