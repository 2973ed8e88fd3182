# Round-2 measurement pass on ONE B200 (run with gpurun; outputs under gpurun_out/, summaries copied to profiles/ by hand).
set -x
R=r02
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_${R}.err | tail -1 > gpurun_out/bench_${R}_rsa_k17.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_${R}_reference.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1200 --csv --log-file gpurun_out/launches_${R}.csv $CMD > gpurun_out/ncu_l.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_${R}.csv gpurun_out/${R}_launches_summary.txt "ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1200 $CMD"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_msm_accum -s 24 -c 8 --csv --log-file gpurun_out/accum_traffic_${R}.csv $CMD > gpurun_out/ncu_t.log 2>&1
python tools/ncu_summary.py traffic gpurun_out/accum_traffic_${R}.csv gpurun_out/${R}_msm_accum_traffic.json k_msm_accum
ncu --set full --clock-control none --import-source on -k regex:k_msm_accum -s 26 -c 2 -o gpurun_out/prof_msm_accum_${R} -f $CMD > gpurun_out/ncu_f.log 2>&1
ncu -i gpurun_out/prof_msm_accum_${R}.ncu-rep --page raw --csv > gpurun_out/raw_accum.csv 2>/dev/null
python tools/ncu_summary.py metrics gpurun_out/raw_accum.csv gpurun_out/${R}_msm_accum_ncu_full.csv
ncu --set full --clock-control none --import-source on -k regex:k_ntt_ -s 40 -c 4 -o gpurun_out/prof_ntt_${R} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --workload sha_k19 > gpurun_out/ncu_n.log 2>&1
ncu -i gpurun_out/prof_ntt_${R}.ncu-rep --page raw --csv > gpurun_out/raw_ntt.csv 2>/dev/null
python tools/ncu_summary.py metrics gpurun_out/raw_ntt.csv gpurun_out/${R}_ntt_ncu_full.csv
ncu --set full --clock-control none --import-source on -k regex:k_eval_program -s 4 -c 2 -o gpurun_out/prof_eval_${R} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --workload sha_k19 > gpurun_out/ncu_e.log 2>&1
ncu -i gpurun_out/prof_eval_${R}.ncu-rep --page raw --csv > gpurun_out/raw_eval.csv 2>/dev/null
python tools/ncu_summary.py metrics gpurun_out/raw_eval.csv gpurun_out/${R}_eval_program_ncu_full.csv
rm -f gpurun_out/raw_accum.csv gpurun_out/raw_ntt.csv gpurun_out/raw_eval.csv
KS=15,17,19,22 python tools/opbench.py > gpurun_out/opbench_${R}.log 2>&1
cp gpurun_out/opbench.json gpurun_out/${R}_opbench.json
python - <<EOF
import json
d=json.load(open("gpurun_out/bench_${R}_rsa_k17.json"))
print("rsa_k17", d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("whole_msm_frac"), d["cpu_baseline"]["value"], d.get("parity"))
for k,v in d.get("other_workloads",{}).items():
    print(k, v.get("value"), v.get("e2e",{}).get("value") if isinstance(v.get("e2e"),dict) else None, (v.get("roofline") or {}).get("frac"), v.get("skipped"))
print(open("gpurun_out/bench_${R}_reference.json").read()[:600])
EOF
head -14 gpurun_out/${R}_launches_summary.txt
