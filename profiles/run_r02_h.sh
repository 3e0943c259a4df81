mkdir -p gpurun_out
nvidia-smi -L | wc -l
python profiles/bench_multi.py 100000 > gpurun_out/r02_bench_multi_n8.json 2> gpurun_out/r02_bench_multi_n8.err; cat gpurun_out/r02_bench_multi_n8.json; tail -3 gpurun_out/r02_bench_multi_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -c 1500 gpurun_out/r02_bench_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 --total-pairs 1000000 > gpurun_out/r02_bench_n8_strong.json 2> gpurun_out/r02_bench_n8_strong.err; tail -c 1500 gpurun_out/r02_bench_n8_strong.json
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
