mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_gpu_tests_c.log; cat gpurun_out/r02_gpu_tests_c.log
g++ -std=c++17 -O2 -I include profiles/bench_decompose.cpp -o /tmp/bench_decompose -Ltracy_b200 -ltracy_b200 -Wl,-rpath,$PWD/tracy_b200 && TRACY_B200_TIMING=1 /tmp/bench_decompose 10000 > gpurun_out/r02_bench_decompose_cpp3.json 2> gpurun_out/r02_bench_decompose_cpp3.err; cat gpurun_out/r02_bench_decompose_cpp3.json; cat gpurun_out/r02_bench_decompose_cpp3.err | tail -10
TRACY_B200_TRACE=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_trace_bench.json 2> gpurun_out/r02_e2e_timeline.txt; tail -60 gpurun_out/r02_e2e_timeline.txt
