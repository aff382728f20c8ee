# source-level counters of the step kernel at step 21 of cfg2 (60,000 columns): gpurun -- 'bash tools/ncu_src.sh name'
ncu --section SourceCounters --import-source on --clock-control none -k regex:kpp_step_kernel -s 20 -c 1 -f -o gpurun_out/src_${1:-cur} python tools/perf_run.py cfg2 300 200 22 0 > /dev/null 2>&1
ls -la gpurun_out/src_${1:-cur}.ncu-rep
