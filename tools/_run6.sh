set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
WORKLOAD=agg_k20 STEPS=2 $TR tools/team_check.py 2>gpurun_out/r02_tc8.err | tail -1 > gpurun_out/r02_team_check_8gpu.json
cat gpurun_out/r02_team_check_8gpu.json; tail -3 gpurun_out/r02_tc8.err
$TR bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/r02_s8.err | tail -1 > gpurun_out/r02_scale8.json
tail -3 gpurun_out/r02_s8.err
ZKC_TEAM_COMMIT_BY_COLUMN=1 $TR bench.py --gpus 8 --team --workload agg_k22 --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>gpurun_out/r02_t8c.err | tail -1 > gpurun_out/r02_team8_agg_k22_bycol.json
python - <<EOF
import json
d=json.load(open("gpurun_out/r02_scale8.json")); print("scale8", d["ms_per_step"], d["value"]); 
for k,v in d.get("team",{}).items(): print(k, v.get("value"), v.get("e2e",{}).get("value"), json.dumps(v.get("phases_ms_per_step")))
d=json.load(open("gpurun_out/r02_team8_agg_k22_bycol.json")); print("bycol", d["ms_per_step"], json.dumps(d.get("phases_ms_per_step")))
EOF
