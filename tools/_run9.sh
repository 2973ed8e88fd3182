set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/r02_s8.err | tail -1 > gpurun_out/r02_scale8_final.json
tail -2 gpurun_out/r02_s8.err
$TR bench.py --gpus 8 --workload chain4 --steps 3 --warmup 1 2>gpurun_out/r02_c8.err | tail -1 > gpurun_out/r02_chain4_8gpu.json
tail -2 gpurun_out/r02_c8.err
python - <<EOF
import json
d=json.load(open("gpurun_out/r02_scale8_final.json")); print("scale8", d["ms_per_step"], d["value"])
for k,v in d.get("team",{}).items(): print(k, v.get("value"), v.get("e2e",{}).get("value"), json.dumps(v.get("phases_ms_per_step")))
d=json.load(open("gpurun_out/r02_chain4_8gpu.json")); print("chain4", d["ms_per_step"], json.dumps(d["config"].get("teams")), d.get("one_proof_per_gpu_ms"), d["config"].get("parity"))
EOF
