set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/p2p_bench.py 2>gpurun_out/r02_p2p2.err | tail -1 > gpurun_out/r02_p2p_2gpu.json
cat gpurun_out/r02_p2p_2gpu.json
$TR bench.py --gpus 2 --team --workload agg_k22 --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>gpurun_out/r02_t2.err | tail -1 > gpurun_out/r02_team2_agg_k22.json
python - <<EOF
import json
d=json.load(open("gpurun_out/r02_team2_agg_k22.json")); print(d["ms_per_step"], d["e2e"]["value"], json.dumps(d.get("phases_ms_per_step")))
EOF
tail -5 gpurun_out/r02_t2.err
