"""GPU tier, needs TWO devices (skipped on a single-GPU box; run with `gpurun --gpus 2`): data parallelism
through our NCCL communicator.  Two ranks, each with half of a batch, must end a training step where one
device with the whole batch ends it (exact mode: allreduce dy, threshold backward on the averaged dy)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LR, STEPS, B = 1e-3, 2, 8


def _net():
    from learned_quantization_b200.training.models import SelfCheckNet
    return SelfCheckNet()


def _prepare(net, x):
    net(x[:1])
    g = torch.Generator(device="cuda").manual_seed(3)
    with torch.no_grad():
        for q in net.quant_layers():
            q.scale.copy_(torch.empty_like(q.scale).uniform_(5e-3, 2e-2, generator=g))


def _data():
    g = torch.Generator().manual_seed(11)
    return torch.rand(B, 10, 10, 3, generator=g), torch.randint(0, 10, (B,), generator=g)


def _worker(rank, world, port, out_path, kind="nccl", graph_steps=0):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    import torch.distributed as dist
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from learned_quantization_b200.custom_loss_terms.custom_components.custom_loss_functions import \
        sparse_categorical_crossentropy as scce
    from learned_quantization_b200.training.dp import NcclComm, PeerComm
    from learned_quantization_b200.training.step import TrainStep
    torch.manual_seed(0)
    x, y = _data()
    x, y = x.cuda(), y.cuda()
    net = _net().cuda()
    _prepare(net, x)
    comm = NcclComm() if kind == "nccl" else PeerComm()
    step = TrainStep(net, scce, optimizer="adam", lr=LR, comm=comm, ds_mode="exact")
    lo, hi = rank * B // world, (rank + 1) * B // world
    if graph_steps:                 # capture() runs 3 eager steps itself; the rest are replays of ONE graph
        step.capture(x[lo:hi].contiguous(), y[lo:hi].contiguous(), warmup=3)
        assert len(step._graph) == (1 if kind == "peer" else 3)
        assert kind != "peer" or (step._early and step._late)      # the exchange runs in two overlapped pieces
        for _ in range(graph_steps - 3):
            step.replay(x[lo:hi], y[lo:hi])
    else:
        for _ in range(STEPS):
            step(x[lo:hi], y[lo:hi])
    torch.cuda.synchronize()
    if kind == "peer":
        assert not comm.timed_out()
    params = [p.detach().cpu().numpy().copy() for p in net.parameters()]
    gathered = [None] * world
    dist.all_gather_object(gathered, params)
    if rank == 0:
        for a, b in zip(gathered[0], gathered[1]):
            assert np.array_equal(a, b), "ranks diverged: exact mode must leave bit-identical parameters everywhere"
        np.savez(out_path, *params)
    comm.close()
    dist.destroy_process_group()


def _single_device(steps):
    from learned_quantization_b200.custom_loss_terms.custom_components.custom_loss_functions import \
        sparse_categorical_crossentropy as scce
    from learned_quantization_b200.training.step import TrainStep
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    x, y = _data()
    x, y = x.cuda(), y.cuda()
    net = _net().cuda()
    _prepare(net, x)
    step = TrainStep(net, scce, optimizer="adam", lr=LR)
    for _ in range(steps):
        step(x, y)
    return [p.detach().cpu().numpy() for p in net.parameters()]


@pytest.mark.parametrize("kind,graph_steps", [("nccl", 0), ("peer", 0), ("peer", 5)])
def test_two_ranks_equal_one_device_on_the_whole_batch(tmp_path, kind, graph_steps):
    """kind "nccl": lq_dp_allreduce + lq_mt_fq_bwd; "peer": the fused peer-memory exchange (lq_dpx_exchange),
    eager and as ONE CUDA graph per step."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "dp.npz")
    port = 29671 + (0 if kind == "nccl" else 1) + graph_steps
    mp.spawn(_worker, args=(2, port, out, kind, graph_steps), nprocs=2, join=True)
    got = np.load(out)
    steps = graph_steps or STEPS
    for i, want in enumerate(_single_device(steps)):          # the same steps on one device with the whole batch
        assert float(np.max(np.abs(got[f"arr_{i}"] - want))) <= 0.02 * LR * steps / STEPS, i
