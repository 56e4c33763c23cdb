"""Rebuilds the CUDA library with different -D tuning macros (on the GPU box) and runs the headline bench for each.
usage: python tools/variant_sweep.py "-DRASTER_MIN_BLOCKS=4" "-DRASTER_MIN_BLOCKS=6" ..."""
import json, os, subprocess, sys
for flags in sys.argv[1:]:
    env = dict(os.environ, DFPSR_NVCC_EXTRA=flags)
    subprocess.check_call([sys.executable, "-c", "from dfpsr_b200 import build; build.build(force=True)"], env=env)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-extras", "--no-cpu-baseline"], capture_output=True, text=True)
    try:
        line = json.loads(out.stdout.strip().splitlines()[-1])
        k = line["roofline"]["per_kernel_us_per_frame"]
        single = subprocess.run([sys.executable, "tools/single_frame_profile.py"], capture_output=True, text=True).stdout.strip().splitlines()
        tiny = subprocess.run([sys.executable, "tools/tiny_profile.py"], capture_output=True, text=True).stdout.strip().splitlines()
        print(tiny[-1][:160] if tiny else "", flush=True)
        print(flags, "| fps", round(line["value"]), "| raster us", round(k.get("raster_kernel<false>", 0), 1), "| setup us", round(k.get("setup_kernel<true>", 0), 1), "|", single[0] if single else "", flush=True)
    except Exception as exc:
        print(flags, "failed", exc, out.stderr[-800:], flush=True)
