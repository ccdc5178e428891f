set -u
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2b_pytest_gpu.log
python tools/diag_gpu.py mockup_only 1024 1000 20 > gpurun_out/r2b_diag_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_visibility_bins" -c 1 -o gpurun_out/r2b_bins python tools/diag_gpu.py mockup_only 1024 1000 20 > gpurun_out/r2b_diag_ncu.log 2>&1; echo "ncu K4 rc $?"
head -4 gpurun_out/r2b_diag_plain.log
