set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/r02_s2.err | tail -1 > gpurun_out/r02_scale2_final.json
python tools/parity_big.py > gpurun_out/r02_parity_big.log 2>&1; tail -3 gpurun_out/r02_parity_big.log
python - <<EOF
import json
d=json.load(open("gpurun_out/r02_scale2_final.json")); print("scale2", d["ms_per_step"], d["value"])
for k,v in d.get("team",{}).items(): print(k, v.get("value"), v.get("e2e",{}).get("value"))
EOF
