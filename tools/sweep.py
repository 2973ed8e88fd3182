"""Development helper: run bench.py for several values of an environment knob and print one summary line each."""
import json
import os
import subprocess
import sys

knob, values, workload, steps = sys.argv[1], sys.argv[2].split(","), sys.argv[3], sys.argv[4]
for v in values:
    env = dict(os.environ)
    if v != "default":
        env[knob] = v
    out = subprocess.run([sys.executable, "bench.py", "--workload", workload, "--steps", steps, "--warmup", "3", "--no-cpu-baseline"],
                         env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        p = d["phases_ms_per_step"]
        print(workload, knob, v, "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"] * 1e3, 2), "frac", round(d["roofline"]["frac"], 3),
              {k: round(x, 2) for k, x in p.items() if k.startswith("msm")}, flush=True)
    except Exception as e:
        print(workload, knob, v, "FAILED", e, out.stderr[-500:], flush=True)
