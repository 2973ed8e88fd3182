# Last pass of a round on ONE B200: the GPU suite, the default bench line (copied to profiles/ by hand) and two light ncu passes that
# list fmaheavy utilisation per launch for the NTT (SHA k=19 proof) and the MSM accumulate (RSA k=17 proof).
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_r02.err | tail -1 > gpurun_out/bench_r02_rsa_k17.json
M=gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum.pct_of_peak_sustained_elapsed,launch__grid_size
ncu --metrics $M --clock-control none -k regex:k_ntt_ -s 500 -c 260 --csv --log-file gpurun_out/r02_ntt_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --workload sha_k19 > gpurun_out/ncu_n2.log 2>&1
ncu --metrics $M --clock-control none -k regex:k_msm_accum -s 24 -c 14 --csv --log-file gpurun_out/r02_accum_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_a2.log 2>&1
python - <<EOF
import json
d=json.load(open("gpurun_out/bench_r02_rsa_k17.json"))
print("rsa_k17", d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("whole_msm_frac"), d["cpu_baseline"]["value"], d.get("parity"))
for k,v in d.get("other_workloads",{}).items():
    print(k, v.get("value"), v.get("e2e",{}).get("value") if isinstance(v.get("e2e"),dict) else None, (v.get("roofline") or {}).get("frac"), v.get("skipped"))
EOF
