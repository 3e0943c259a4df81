mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_gpu_tests_g.log; cat gpurun_out/r02_gpu_tests_g.log
oracle/_ref/dropin_test | tail -2
g++ -std=c++17 -O2 -I include profiles/bench_assemble.cpp -o /tmp/bench_assemble -Ltracy_b200 -ltracy_b200 -Wl,-rpath,$PWD/tracy_b200 && /tmp/bench_assemble 512 > gpurun_out/r02_bench_assemble_cpp.json 2> gpurun_out/r02_bench_assemble_cpp.err; cat gpurun_out/r02_bench_assemble_cpp.json
timeout 600 python profiles/prof_assemble_stages.py > gpurun_out/r02_assemble_stages.json 2>&1; tail -c 900 gpurun_out/r02_assemble_stages.json
timeout 900 python profiles/bench_assemble_files.py > gpurun_out/r02_bench_assemble_files.json 2> gpurun_out/r02_bench_assemble_files.err; tail -c 600 gpurun_out/r02_bench_assemble_files.json
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gotoh_pp -c 1 -f -o gpurun_out/prof_r02_pp_screen python profiles/pp_probe.py > gpurun_out/ncu_full_r02_pp_screen.log 2>&1; tail -3 gpurun_out/ncu_full_r02_pp_screen.log
timeout 600 python profiles/pp_probe.py > gpurun_out/r02_pp_probe.json 2> gpurun_out/r02_pp_probe.err; grep -A3 "score_fp32x2_arr\"\|traceback_fp32x2_arr\"" gpurun_out/r02_pp_probe.json | grep gcups
