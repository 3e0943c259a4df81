mkdir -p gpurun_out
oracle/_ref/dropin_test | tail -2
g++ -std=c++17 -O2 -I include profiles/bench_assemble.cpp -o /tmp/bench_assemble -Ltracy_b200 -ltracy_b200 -Wl,-rpath,$PWD/tracy_b200 && for i in 1 2; do /tmp/bench_assemble 512 > gpurun_out/r02_bench_assemble_cpp.json 2> gpurun_out/r02_bench_assemble_cpp.err; cat gpurun_out/r02_bench_assemble_cpp.json; done
timeout 600 python profiles/prof_assemble_stages.py > gpurun_out/r02_assemble_stages.json 2>&1; tail -c 900 gpurun_out/r02_assemble_stages.json
timeout 900 python profiles/bench_assemble_files.py > gpurun_out/r02_bench_assemble_files.json 2> gpurun_out/r02_bench_assemble_files.err; head -c 300 gpurun_out/r02_bench_assemble_files.json
