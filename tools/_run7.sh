set -x
timeout 600 python -m pytest tests/test_gpu_prover.py -x -q 2>&1 | tail -3
for v in 0 1; do
  if [ $v = 1 ]; then export ZKC_NO_PROGRAM_FACTORING=1; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --workload sha_k19 2>/dev/null | tail -1 > gpurun_out/r02_e_sha_$v.json
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras --workload agg_k20 2>/dev/null | tail -1 > gpurun_out/r02_e_agg20_$v.json
done
python - <<EOF
import json
for f in ["sha_0","sha_1","agg20_0","agg20_1"]:
    d=json.load(open("gpurun_out/r02_e_%s.json"%f)); print(f, d["ms_per_step"], d["phases_ms_per_step"]["eval_program"], d["phases_ms_per_step"]["prove.quotient"])
EOF
