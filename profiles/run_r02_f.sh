mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/r02_bench_j.json 2> gpurun_out/r02_bench_j.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_j.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_ops_only','kernel_ms','gpu_launches')})
PY
TRACY_B200_CHUNK_TARGET=16384 TRACY_B200_CHUNK_PARTS=8 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_j2.json 2> gpurun_out/r02_bench_j2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_j2.json').read().strip().splitlines()[-1])
print("7-wave chunks:", {k:d[k] for k in ('e2e','e2e_ops_only')})
PY
timeout 900 python -m pytest tests/test_gpu_packed.py -m gpu -q -x 2>&1 | tail -3
