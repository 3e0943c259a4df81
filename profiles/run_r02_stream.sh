mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_packed.py -q -x -k "streamed or ramp" 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_stream.json 2> gpurun_out/r02_bench_stream.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_stream.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches') if k in d})
PY
TRACY_B200_NO_STREAM=1 timeout 600 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NO_STREAM', d['e2e'])"
