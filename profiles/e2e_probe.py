import os, sys, time, json, subprocess
for ch in ["3552","7104","12500","25000","50000"]:
    env=dict(os.environ, TRACY_B200_CHUNK=ch)
    out=subprocess.run([sys.executable,"bench.py","--no-cpu-baseline","--steps","3","--warmup","2"],env=env,capture_output=True,text=True).stdout.strip().splitlines()[-1]
    d=json.loads(out); print(ch, round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), flush=True)
