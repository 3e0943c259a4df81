mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_pp.py -q -x 2>&1 | tail -6 > gpurun_out/r02_pp_screen_tests.log; cat gpurun_out/r02_pp_screen_tests.log
timeout 600 python profiles/pp_probe.py > gpurun_out/r02_pp_probe_screen.json 2> gpurun_out/r02_pp_probe_screen.err; grep -A4 "score_fp32x2\|traceback_fp32x2\|big_pair" gpurun_out/r02_pp_probe_screen.json | grep "gcups\|all_pairs\|big_pair\|kernel_ms"
