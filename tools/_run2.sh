set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/r02_d1.err | tail -1 > gpurun_out/r02_d1_k17.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --workload sha_k19 2>gpurun_out/r02_d2.err | tail -1 > gpurun_out/r02_d1_sha.json
python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras --workload agg_k22 2>gpurun_out/r02_d3.err | tail -1 > gpurun_out/r02_d1_agg.json
python - <<EOF
import json
for f in ["k17","sha","agg"]:
    try:
        d=json.load(open("gpurun_out/r02_d1_%s.json"%f)); print(f, d["ms_per_step"], d["e2e"]["value"], {k:v for k,v in d["phases_ms_per_step"].items() if k.startswith("ntt") or k.startswith("prove.q") or k.startswith("eval_p")})
    except Exception as e: print(f, "ERR", e)
EOF
