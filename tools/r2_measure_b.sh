set -x
python tools/size_sweep.py cfg4 "350x250,500x350,700x500,1000x700" "0" 8 2>&1 | tee gpurun_out/r2_sizes_b.txt
python tools/size_sweep.py cfg5 "300x200" "0" 8 2>&1 | tee -a gpurun_out/r2_sizes_b.txt
python tools/perf_run.py cfg4 400 300 8 > gpurun_out/plain_cfg4.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:kpp_step_kernel -s 6 -c 1 -f -o gpurun_out/r2b_step_cfg4 python tools/perf_run.py cfg4 400 300 8 > gpurun_out/ncu_cfg4.log 2>&1
python tools/perf_run.py cfg5 300 200 8 > gpurun_out/plain_cfg5.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:kpp_step_kernel -s 6 -c 1 -f -o gpurun_out/r2b_step_cfg5 python tools/perf_run.py cfg5 300 200 8 > gpurun_out/ncu_cfg5.log 2>&1
python tools/perf_run.py cfg2 300 200 8 > gpurun_out/plain_cfg2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:kpp_step_kernel -s 6 -c 1 -f -o gpurun_out/r2b_step_cfg2 python tools/perf_run.py cfg2 300 200 8 > gpurun_out/ncu_cfg2.log 2>&1
grep median gpurun_out/plain_cfg*.log
