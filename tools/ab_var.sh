for i in 1 2; do
for v in "$@"; do
KPP_LIB_PATH=$PWD/_var/$v/libkpp_gpu.so python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'ms', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e6,3))"
done; done
