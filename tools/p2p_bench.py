"""NCCL point-to-point vs collective bandwidth between the GPUs of one node (sizes of the team-proving exchanges).
torchrun --nproc-per-node N tools/p2p_bench.py  -> one JSON line on rank 0 (GB/s per rank, device-timed, max over ranks)."""
import json
import os
import torch
import torch.distributed as dist


def timed(fn, iters=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    out = {"world": world}
    peer = rank ^ 1
    for mb, nmsg in [(64, 1), (64, 8), (8, 64), (512, 1)]:
        n = mb << 20
        src = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(nmsg)]
        dst = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(nmsg)]

        def pair():
            ops = []
            for i in range(nmsg):
                ops.append(dist.P2POp(dist.isend, src[i], peer))
                ops.append(dist.P2POp(dist.irecv, dst[i], peer))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        ms = timed(pair)
        out["pair_exchange_%dx%dMB_GBps_each_way" % (nmsg, mb)] = round(nmsg * n / ms / 1e6, 1)
        del src, dst
    # all ranks to all ranks, 64 MB per (src, dst) pair
    n = 64 << 20
    a = torch.empty(n * world, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
    ms = timed(lambda: dist.all_to_all_single(b, a))
    out["all_to_all_64MB_per_pair_GBps_recv_per_rank"] = round(n * (world - 1) / ms / 1e6, 1)
    ms = timed(lambda: dist.all_gather_into_tensor(b, a[:n]))
    out["all_gather_64MB_per_rank_GBps_recv_per_rank"] = round(n * (world - 1) / ms / 1e6, 1)
    ms = timed(lambda: dist.broadcast(a[:n * min(world, 4)], 0))
    out["broadcast_%dMB_GBps" % (64 * min(world, 4))] = round(n * min(world, 4) / ms / 1e6, 1)
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
