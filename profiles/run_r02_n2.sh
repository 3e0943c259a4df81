mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['n_gpus'])"
oracle/_ref/dropin_test | tail -1
