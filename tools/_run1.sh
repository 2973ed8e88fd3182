set -x
timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_msm.py -x -q 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_steps.py -x -q 2>&1 | tail -5
ZKC_MSM_ACCUM_OCC=4 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/r02_c1.err | tail -1 > gpurun_out/r02_c1_occ4.json
ZKC_MSM_ACCUM_OCC=3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/r02_c2.err | tail -1 > gpurun_out/r02_c1_occ3.json
python - <<EOF
import json
for f in ["occ4","occ3"]:
    d=json.load(open("gpurun_out/r02_c1_%s.json"%f)); print(f, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_msm_frac"], {k:v for k,v in d["phases_ms_per_step"].items() if k.startswith("msm")})
EOF
