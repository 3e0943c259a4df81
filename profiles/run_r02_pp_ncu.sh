mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gotoh_pp -c 1 -f -o gpurun_out/prof_r02_pp_screen python profiles/pp_probe.py > gpurun_out/ncu_full_r02_pp_screen.log 2>&1; tail -3 gpurun_out/ncu_full_r02_pp_screen.log
