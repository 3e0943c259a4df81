mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_packed.py -q -x -k "streamed or ramp" 2>&1 | tail -2
timeout 300 python profiles/stream_sanity.py 2>&1 | grep -v "chunk [0-9]* lane\|streamed chunk" | tail -3
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_stream.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "ncu rc=$?"; grep -c gotoh_packed gpurun_out/r02_launches_stream.csv; tail -3 gpurun_out/r02_launches_stream.csv | cut -c1-200
timeout 300 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', d['value'], d['e2e'], d['e2e_ops_only']['ms_per_step'])"
