mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_gpu_tests_final.log; cat gpurun_out/r02_gpu_tests_final.log
oracle/_ref/dropin_test | tail -2
bash profiles/run_profile.sh r02f
