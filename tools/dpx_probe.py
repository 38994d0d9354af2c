"""Under torchrun (N >= 2): the fused peer-memory exchange alone against ncclAllReduce, over buffer sizes.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dpx_probe.py
One JSON line per size on rank 0: time = fixed latency + bytes / bandwidth; busbw = 2 (W-1)/W bytes / t."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from learned_quantization_b200.training.dp import NcclComm, PeerComm
    sizes = [1 << 18, 1 << 20, 1 << 22, 11186444, 1 << 25]
    peer = PeerComm()
    arena = peer.allocate(max(sizes) * 4)
    flat_all = arena.view(torch.float32)
    nccl = NcclComm()
    other = torch.empty(max(sizes), device="cuda")
    K = 20

    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            fn()
        b.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / K], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for n in sizes:
        flat = flat_all[:n]
        flat.normal_()
        o = other[:n]
        o.normal_()
        t_peer = timed(lambda: peer.allreduce_mean_(flat))
        t_nccl = timed(lambda: nccl.allreduce_mean_(o))
        m4 = n // 4 * 4
        dma = {b: timed(lambda: peer.allreduce_dma_(flat[:m4], 1, b)) * 1e3 for b in (8, 16, 32, 64)}
        if rank == 0:
            bus = lambda ms: 2.0 * (world - 1) / world * n * 4 / (ms * 1e-3) / 1e9
            print(json.dumps({"floats": n, "world": world, "peer_us": t_peer * 1e3, "nccl_us": t_nccl * 1e3,
                              "peer_busbw_GBps": bus(t_peer), "nccl_busbw_GBps": bus(t_nccl), "dma_us_by_blocks": dma,
                              "grid_env": os.environ.get("LQ_DPX_GRID", "")}), flush=True)
    assert not peer.timed_out()
    dist.barrier()
    nccl.close()
    peer.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
