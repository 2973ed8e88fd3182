set -x
timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prover.py tests/test_gpu_team.py -x -q 2>&1 | tail -3
WL=sha_k19 CS=0 TS=0 OCCS=0 python tools/sweep_c.py 2>&1 | grep "^c"
WL=rsa_k17 CS=0 TS=0 OCCS=0 python tools/sweep_c.py 2>&1 | grep "^c"
WL=agg_k20 CS=0 TS=0 OCCS=0 python tools/sweep_c.py 2>&1 | grep "^c"
M=gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum.pct_of_peak_sustained_elapsed
ncu --metrics $M --clock-control none -k regex:k_ntt_ -s 120 -c 120 --csv --log-file gpurun_out/r02_ntt_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --workload sha_k19 > gpurun_out/ncu_n2.log 2>&1
grep -c k_ntt gpurun_out/r02_ntt_launches.csv
