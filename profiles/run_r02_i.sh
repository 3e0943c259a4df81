mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_gpu_tests_i.log; cat gpurun_out/r02_gpu_tests_i.log
oracle/_ref/dropin_test | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_i.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['e2e_ops_only']['ms_per_step'], d['gpu_launches'])"
