mkdir -p gpurun_out
g++ -std=c++17 -O2 -I include profiles/bench_decompose.cpp -o /tmp/bench_decompose -Ltracy_b200 -ltracy_b200 -Wl,-rpath,$PWD/tracy_b200 && /tmp/bench_decompose 10000 > gpurun_out/r02_bench_decompose_cpp.json 2> gpurun_out/r02_bench_decompose_cpp.err; cat gpurun_out/r02_bench_decompose_cpp.json
timeout 600 python profiles/bench_decompose.py 2000 > gpurun_out/r02_bench_decompose_py.json 2> gpurun_out/r02_bench_decompose_py.err; tail -c 600 gpurun_out/r02_bench_decompose_py.json
bash profiles/run_profile.sh r02
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gotoh_pp -c 1 -f -o gpurun_out/prof_r02_pp python profiles/pp_probe.py > gpurun_out/ncu_full_r02_pp.log 2>&1; tail -3 gpurun_out/ncu_full_r02_pp.log
