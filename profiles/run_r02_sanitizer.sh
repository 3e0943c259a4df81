mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
( echo "== memcheck: fraction + outputs + pp (small) + device pipeline + packed shapes/spans"; timeout 1500 $SAN --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_fraction.py tests/test_gpu_outputs.py tests/test_gpu_device_pipeline.py "tests/test_gpu_packed.py" -m gpu -q -x -k "fraction or outputs or rows or packed_ops or pipeline or shapes or spans or string_pairs" 2>&1 | tail -8; echo "rc=$?" ) > gpurun_out/r02_sanitizer_memcheck.log 2>&1
( echo "== memcheck: pp kernel incl. big pairs"; timeout 900 $SAN --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_pp.py -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/r02_sanitizer_pp.log 2>&1
( echo "== racecheck: pp big pairs + outputs"; timeout 900 $SAN --tool racecheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_outputs.py -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/r02_sanitizer_race.log 2>&1
tail -12 gpurun_out/r02_sanitizer_memcheck.log gpurun_out/r02_sanitizer_pp.log gpurun_out/r02_sanitizer_race.log
