set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
WORKLOAD=agg_k20 STEPS=2 $TR --nproc-per-node 3 tools/team_check.py 2>gpurun_out/r02_tc3.err | tail -1 > gpurun_out/r02_team_check_3gpu.json
cat gpurun_out/r02_team_check_3gpu.json; tail -3 gpurun_out/r02_tc3.err
$TR --nproc-per-node 4 bench.py --gpus 4 --team --workload agg_k22 --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>gpurun_out/r02_t4.err | tail -1 > gpurun_out/r02_team4_agg_k22.json
tail -3 gpurun_out/r02_t4.err
