"""e2e (TB_MEM_HOST) vs lanes x chunk size; run on the GPU box."""
import os, sys, json, subprocess
for lanes in ["2", "3", "4"]:
    for ch in ["7104", "12432", "24864"]:
        env = dict(os.environ, TRACY_B200_CHUNK=ch, TRACY_B200_LANES=lanes)
        out = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--steps", "3", "--warmup", "2"], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
        j = json.loads(out); print("lanes", lanes, "chunk", ch, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), round(j["e2e"]["ms_per_step"], 1), flush=True)
