"""Lab: C4 group-by time under env-knob variants (GPU box only)."""
import json, os, subprocess, sys
scale = sys.argv[1] if len(sys.argv) > 1 else "1.0"
for v in (sys.argv[2:] or [""]):
    env = dict(os.environ)
    for kv in filter(None, v.split(",")):
        k, val = kv.split("=")
        env[k] = val
    out = subprocess.run([sys.executable, "bench.py", "--only", "groupby", "--no-cpu", "--no-e2e", "--scale", scale, "--steps", "5", "--warmup", "2"],
                         env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        w = d["workloads"]["groupby"]
        print("[%s] scale %s: %.3f ms ok=%s %s" % (v, scale, w["ms_per_step"], w["parity_properties_ok"],
              {k: round(x["ms_per_step"], 3) for k, x in w["kernels"].items()}), flush=True)
    except Exception as e:
        print("[%s] FAILED %s\n%s" % (v, e, out.stderr[-2000:]), flush=True)
