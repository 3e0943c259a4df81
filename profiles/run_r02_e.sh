mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_packed.py tests/test_gpu_outputs.py tests/test_gpu_gotoh.py tests/test_gpu_pp.py tests/test_gpu_subcommands.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_gpu_tests_e.log; cat gpurun_out/r02_gpu_tests_e.log
python bench.py --no-cpu-baseline > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_i.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_ops_only','kernel_ms','gpu_launches')})
PY
TRACY_B200_TRACE=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_trace_bench.json 2> gpurun_out/r02_e2e_timeline.txt; grep chunk gpurun_out/r02_e2e_timeline.txt | tail -13
