# partition sizes of cfg4 (what 8/4/2 GPUs own), plain timing; then ncu --set full of the step kernel on cfg4 (LDD) and cfg5 (NZ=250)
set -x
for sz in "350 250" "500 350" "700 500"; do python tools/perf_run.py cfg4 $sz 10 | tail -1; done > gpurun_out/r2_partition_sizes.txt 2>&1
python tools/perf_run.py cfg3 220 200 10 | tail -1 >> gpurun_out/r2_partition_sizes.txt 2>&1
for sz in "110 50" "110 100" "220 100"; do python tools/perf_run.py cfg3 $sz 10 | tail -1; done >> gpurun_out/r2_partition_sizes.txt 2>&1
cat gpurun_out/r2_partition_sizes.txt
(cd tools/micro && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_micro fp64_micro.cu && ./fp64_micro) > gpurun_out/r2_fp64_micro.txt 2>&1; tail -12 gpurun_out/r2_fp64_micro.txt
python tools/perf_run.py cfg4 400 300 8 > gpurun_out/plain_cfg4.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:kpp_step_kernel -s 6 -c 1 -f -o gpurun_out/r2_step_cfg4 python tools/perf_run.py cfg4 400 300 8 > gpurun_out/ncu_cfg4.log 2>&1
python tools/perf_run.py cfg5 300 200 8 > gpurun_out/plain_cfg5.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:kpp_step_kernel -s 6 -c 1 -f -o gpurun_out/r2_step_cfg5 python tools/perf_run.py cfg5 300 200 8 > gpurun_out/ncu_cfg5.log 2>&1
tail -2 gpurun_out/plain_cfg4.log gpurun_out/plain_cfg5.log gpurun_out/ncu_cfg4.log gpurun_out/ncu_cfg5.log
ls -la gpurun_out/
