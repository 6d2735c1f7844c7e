for i in 1 2; do
for a in 1 0; do
FDC_E2E_ASYNC_READ=$a python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('async=$a', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['latency_ms'], d['e2e_present']['ms_per_step'])"
done; done
