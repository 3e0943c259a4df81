mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_packed.py -q -x -k "streamed or ramp" 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['n_gpus'])"
timeout 600 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', d['value'], d['e2e'])"
