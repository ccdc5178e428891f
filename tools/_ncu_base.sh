set -u
mkdir -p /tmp/newpkg && cp inspection_path_planning_b200/*.py inspection_path_planning_b200/libipp.so /tmp/newpkg/
cp _ab/base_pkg/*.py _ab/base_pkg/libipp.so inspection_path_planning_b200/
python tools/diag_gpu.py luis4 1024 60 30 > gpurun_out/diag_base_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_visibility_bins" -c 1 -o gpurun_out/r2_bins_base_luis4 python tools/diag_gpu.py luis4 1024 60 30 > gpurun_out/diag_base_ncu.log 2>&1
cp /tmp/newpkg/*.py /tmp/newpkg/libipp.so inspection_path_planning_b200/
cp _ab/sub16.so inspection_path_planning_b200/libipp.so
ncu --set full --clock-control none --import-source on -k regex:"k_visibility_bins" -c 1 -o gpurun_out/r2_bins_sub16_luis4 python tools/diag_gpu.py luis4 1024 60 30 > gpurun_out/diag_sub16_ncu.log 2>&1
cp /tmp/newpkg/libipp.so inspection_path_planning_b200/libipp.so
grep "kernel ms" gpurun_out/diag_base_plain.log | head -1
