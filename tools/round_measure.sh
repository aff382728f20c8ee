# Round-end measurement pass on one B200 (run from the repo root under gpurun):
#   tests, bench, straggler scan, then the ncu launch list and one full capture of the step kernel.
set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
KPP_PASS_BUDGET=1 KPP_TEST_PASS_BUDGET=1 timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 400 gpurun_out/bench_default.json
python bench.py --steps 140 --warmup 4 --no-cpu-baseline > gpurun_out/bench_140.json 2>/dev/null
timeout 300 python tools/iter_scan.py cfg2 300 200 130 > gpurun_out/iter_scan_budget6.txt 2>&1
KPP_PASS_BUDGET=0 timeout 300 python tools/iter_scan.py cfg2 300 200 100 > gpurun_out/iter_scan_budget0.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:kpp_step_kernel -s 20 -c 1 -f -o gpurun_out/step_full python tools/perf_run.py cfg2 300 200 22 0 > /dev/null 2>&1
ls -la gpurun_out/
