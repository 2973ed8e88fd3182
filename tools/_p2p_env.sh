TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo base; $TR tools/p2p_bench.py 2>/dev/null | tail -1
echo BUFF16M; NCCL_BUFFSIZE=16777216 $TR tools/p2p_bench.py 2>/dev/null | tail -1
echo CHUNK2M_BUFF16M; NCCL_BUFFSIZE=16777216 NCCL_P2P_NVL_CHUNKSIZE=2097152 $TR tools/p2p_bench.py 2>/dev/null | tail -1
echo NCH32; NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 $TR tools/p2p_bench.py 2>/dev/null | tail -1
echo NCH32_BUFF16M; NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 NCCL_BUFFSIZE=16777216 $TR tools/p2p_bench.py 2>/dev/null | tail -1
