for cfgargs in "cfg5 300 200 8" "cfg2 300 200 8" "cfg3 220 200 8"; do
for b in 0 6 7 8 10; do
echo -n "budget $b: "; KPP_PASS_BUDGET=$b python tools/perf_run.py $cfgargs 0 2>&1 | tail -1
done; done
