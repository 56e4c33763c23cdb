"""Sweeps the pipeline chunk size of dfpsr_session_render_views_host (DFPSR_PIPELINE_CHUNK) on the bench workload."""
import os, subprocess, sys, json
for chunk in (4, 8, 16, 32, 64):
    env = dict(os.environ, DFPSR_PIPELINE_CHUNK=str(chunk))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-extras", "--no-cpu-baseline"], capture_output=True, text=True, env=env)
    try:
        line = json.loads(out.stdout.strip().splitlines()[-1])
        print(chunk, "e2e fps", round(line["e2e"]["value"]), "device fps", round(line["value"]))
    except Exception as exc:
        print(chunk, "failed", exc, out.stderr[-500:])
