mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_cpp_binding.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_gpu_tests_multi_n2.log; cat gpurun_out/r02_gpu_tests_multi_n2.log
oracle/_ref/dropin_test | tail -3
python profiles/bench_multi.py 100000 > gpurun_out/r02_bench_multi_n2.json 2> gpurun_out/r02_bench_multi_n2.err; cat gpurun_out/r02_bench_multi_n2.json; tail -3 gpurun_out/r02_bench_multi_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','n_gpus','reference_broadcast')})
PY
