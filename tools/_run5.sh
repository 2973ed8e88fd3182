set -x
timeout 900 python -m pytest tests/test_gpu_team.py tests/test_gpu_ntt.py -x -q 2>&1 | tail -4
