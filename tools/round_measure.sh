# Round-end measurement pass on one B200 (run from the repo root under gpurun): tests, bench lines, launch-shape
# sweeps, then the ncu launch list of the bench command and one full capture of the step kernel per shape.
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench_cfg4_1gpu.json 2> gpurun_out/r2_bench_cfg4_1gpu.err; tail -c 600 gpurun_out/r2_bench_cfg4_1gpu.json
python bench.py --config cfg2 --no-other-shapes > gpurun_out/r2_bench_cfg2_1gpu.json 2> gpurun_out/r2_bench_cfg2_1gpu.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
python tools/size_sweep.py cfg4 "350x250,500x350,700x500,1000x700" "0" 8 2>&1 | tee gpurun_out/r2_sizes_final.txt
python tools/size_sweep.py cfg3 "220x200,220x100,110x100,110x50" "0" 8 2>&1 | tee -a gpurun_out/r2_sizes_final.txt
python tools/size_sweep.py cfg2 "300x200" "0" 8 2>&1 | tee -a gpurun_out/r2_sizes_final.txt
python tools/size_sweep.py cfg5 "300x200" "0" 8 2>&1 | tee -a gpurun_out/r2_sizes_final.txt
B="python bench.py --steps 4 --warmup 3 --spinup 10 --no-cpu-baseline --no-output-e2e --no-other-shapes"
$B > gpurun_out/plain_bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launch_list_bench.csv $B > gpurun_out/ncu_launch.log 2>&1
for shape in "cfg4 1000 700" "cfg2 300 200" "cfg5 300 200"; do
  set -- $shape
  python tools/perf_run.py $1 $2 $3 8 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --import-source on --clock-control none -k regex:kpp_step_kernel -s 6 -c 1 -f -o gpurun_out/r2_step_$1 python tools/perf_run.py $1 $2 $3 8 > gpurun_out/ncu_$1.log 2>&1
done
grep median gpurun_out/plain_cfg*.log
ls -la gpurun_out/ | tail -20
