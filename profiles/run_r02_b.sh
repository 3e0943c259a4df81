mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_gpu_tests_b.log; cat gpurun_out/r02_gpu_tests_b.log
g++ -std=c++17 -O2 -I include profiles/bench_decompose.cpp -o /tmp/bench_decompose -Ltracy_b200 -ltracy_b200 -Wl,-rpath,$PWD/tracy_b200 && TRACY_B200_TIMING=1 /tmp/bench_decompose 10000 > gpurun_out/r02_bench_decompose_cpp2.json 2> gpurun_out/r02_bench_decompose_cpp2.err; cat gpurun_out/r02_bench_decompose_cpp2.json; cat gpurun_out/r02_bench_decompose_cpp2.err | tail -12
python profiles/tb_cost_probe.py > gpurun_out/r02_tb_cost.json 2>&1; cat gpurun_out/r02_tb_cost.json
