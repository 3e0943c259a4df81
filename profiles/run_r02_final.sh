# Final pass of round 2 on one B200: smoke, every GPU test, the drop-in binary, the bench line, the ncu launch list of the bench
# command at full size, one --set full capture of the packed kernel (small batch: replay passes), the pp probe.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_gpu_tests_final.log; cat gpurun_out/r02_gpu_tests_final.log
oracle/_ref/dropin_test | tail -2
python bench.py > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err; tail -c 1200 gpurun_out/bench_r02f.json; tail -3 gpurun_out/bench_r02f.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02f.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_r02f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gotoh_packed -s 1 -c 1 -f -o gpurun_out/prof_r02f python bench.py --steps 1 --warmup 1 --pairs 5328 --no-cpu-baseline > gpurun_out/ncu_full_r02f.log 2>&1
tail -3 gpurun_out/ncu_full_r02f.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02f_reference.json 2> /dev/null; tail -c 600 gpurun_out/bench_r02f_reference.json
