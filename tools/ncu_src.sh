for d in _old .; do
 (cd $d && ncu --section SourceCounters --import-source on --clock-control none -k regex:kpp_step_kernel -s 20 -c 1 -f -o $OLDPWD/gpurun_out/src_$(basename $d | tr -d .)x python tools/perf_run.py cfg2 300 200 22 0 > /dev/null 2>&1)
done
ls -la gpurun_out/*.ncu-rep
