set -x
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_final.err | tail -1 > gpurun_out/bench_r01_rsa_k17.json
python bench.py --steps 5 --warmup 3 --workload sha_k19 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_r01_sha_k19.json
python bench.py --steps 10 --warmup 3 --workload rsa_k15 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_r01_rsa_k15.json
python bench.py --steps 3 --warmup 3 --workload agg_k22 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_r01_agg_k22.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r01_reference.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1200 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_l.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_r01.csv gpurun_out/r01_launches_summary.txt "ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1200 $CMD"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_msm_accum -s 24 -c 8 --csv --log-file gpurun_out/accum_traffic_r01.csv $CMD > gpurun_out/ncu_t.log 2>&1
python tools/ncu_summary.py traffic gpurun_out/accum_traffic_r01.csv gpurun_out/r01_msm_accum_traffic.json k_msm_accum
ncu --set full --clock-control none --import-source on -k regex:k_msm_accum -s 26 -c 2 -o gpurun_out/prof_msm_accum_r01 -f $CMD > gpurun_out/ncu_f.log 2>&1
ncu -i gpurun_out/prof_msm_accum_r01.ncu-rep --page raw --csv > gpurun_out/raw_accum.csv 2>/dev/null
python tools/ncu_summary.py metrics gpurun_out/raw_accum.csv gpurun_out/r01_msm_accum_ncu_full.csv
ncu --set full --clock-control none --import-source on -k regex:k_ntt_ -s 40 -c 4 -o gpurun_out/prof_ntt_r01 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload sha_k19 > gpurun_out/ncu_n.log 2>&1
ncu -i gpurun_out/prof_ntt_r01.ncu-rep --page raw --csv > gpurun_out/raw_ntt.csv 2>/dev/null
python tools/ncu_summary.py metrics gpurun_out/raw_ntt.csv gpurun_out/r01_ntt_ncu_full.csv
rm -f gpurun_out/raw_accum.csv gpurun_out/raw_ntt.csv
KS=15,17,19,22 python tools/opbench.py > gpurun_out/opbench.log 2>&1
python - <<EOF
import json
for f in ["rsa_k17","sha_k19","rsa_k15","agg_k22","reference"]:
    try:
        d=json.load(open("gpurun_out/bench_r01_%s.json"%f)); print(f, d["ms_per_step"], d["e2e"]["value"], d.get("roofline",{}).get("frac"))
    except Exception as e: print(f, "ERR", e)
EOF
head -12 gpurun_out/r01_launches_summary.txt
