#!/usr/bin/env python
"""bench.py -- fused fake-quant forward + threshold backward throughput on B200.

Workload (BASELINE.json configs[3] at its headline size): one fp32 tensor of 64 Mi elements,
per-channel scales, layout [outer, C, inner] = [256, 512, 512]; a "step" is one forward
(lq_fq_fwd, 8 B/elem) plus one nested backward (lq_fq_bwd_threshold writing dX, 12 B/elem)
over the whole tensor.  metric = 20 * N / t in GB/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fakequant_fwd_bwd_GBps"
UNIT = "GB/s"
LAYOUT = (256, 512, 512)          # 64 Mi elements, channelwise C = 512 on an HWIO-like tensor
LAMBDA = 1e-3                     # ~3% of elements below threshold: both NQ:103-111 branches live
BYTES_FWD, BYTES_BWD = 8, 12      # algorithmic bytes per element (SURVEY 8d / BASELINE.md 3)
WORKLOAD = "fakequant_fwd_bwd_channelwise_C512_inner512"


def synth(layout, seed, numpy=False):
    """P ~ N(0, 0.05^2), s ~ U[5e-3, 2e-2], dy ~ N(0, 1e-3^2)  (BASELINE.md section 5)."""
    import numpy as np
    n = layout[0] * layout[1] * layout[2]
    rng = np.random.default_rng(seed)
    P = rng.standard_normal(n, dtype=np.float32) * np.float32(0.05)
    dy = rng.standard_normal(n, dtype=np.float32) * np.float32(1e-3)
    s = rng.uniform(5e-3, 2e-2, size=layout[1]).astype(np.float32)
    return P, s, dy


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for name, val in zip(names, r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def ncu_traffic():
    """(bytes, note): dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    summary of an `ncu --set full` capture (profiles/traffic.json, written by tools/refresh_traffic.py) -- reported only
    while the hash of the kernel sources stored with the capture still matches the tree: a stale capture reads null."""
    import hashlib
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)
        if t.get("workload") != WORKLOAD:
            return None, "profiles/traffic.json is for another workload"
        h = hashlib.sha256()
        for name in t["sources"]:
            with open(os.path.join(ROOT, "learned_quantization_b200", "csrc", name), "rb") as f:
                h.update(f.read())
        if h.hexdigest() != t["source_sha256"]:
            return None, f"capture of commit {t.get('commit_at_capture')} predates the current kernel sources: re-run tools/refresh_traffic.py"
        return t["traffic_bytes_per_launch"], f"ncu --set full at commit {t.get('commit_at_capture')}, kernel sources unchanged since"
    except (OSError, ValueError, KeyError) as e:
        return None, f"no usable profiles/traffic.json ({type(e).__name__})"


def bind_to_gpu_numa(local):
    """Pin this process -- and with it the pinned host buffers it allocates from now on (first touch) -- to the NUMA
    node its GPU hangs off: with 8 ranks streaming 0.8 GB per step each through host memory, remote-node buffers halve
    what the root complexes deliver."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return {"gpu": bdf, "numa_node": None, "bound": False}
        cpus = set()
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"gpu": bdf, "numa_node": node, "cpus": len(cpus), "bound": bool(cpus)}
    except (OSError, ValueError, AttributeError) as e:
        return {"numa_node": None, "bound": False, "why": f"{type(e).__name__}: {e}"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------- CPU arms
# The reference's own CPU path is TensorFlow-CPU running the un-fused op chain of NQ:55-118;
# TensorFlow is not installable here, so the timed stand-in is oracle/lq_oracle.c: the same
# chain (one pass + one materialised temporary per reference op), OpenMP over all host cores.
_CO = None


def cpu_oracle():
    global _CO
    if _CO is None:
        from oracle import lq_oracle_c
        _CO = lq_oracle_c.COracle()
    return _CO


def cpu_oracle_step(P, s, dy, layout, lam, out=None):
    return cpu_oracle().fwd_bwd(P, s, dy, layout, lam, out)


def cpu_sample(sample_outer):
    import numpy as np
    layout = (sample_outer, LAYOUT[1], LAYOUT[2])
    P, s, dy = synth(layout, 42)
    out = (np.empty(P.size, np.float32), np.empty(P.size, np.float32))
    return layout, P, s, dy, out


def cpu_describe(layout, n, dt):
    co = cpu_oracle()
    whole = layout[0] == LAYOUT[0]
    return {"unit": UNIT, "cores": co.threads, "kind": "port", "baseline_kind": "c_port", "same_config": whole,
            "sample": f"oracle/lq_oracle.c (reference op chain NQ:55-118, OpenMP x{co.threads}) fwd+bwd on "
                      f"[{layout[0]},{layout[1]},{layout[2]}] = {n} elements "
                      + ("(the whole workload tensor)" if whole else "of the workload") + f", {dt * 1e3:.1f} ms per pass",
            "host_cpus": os.cpu_count()}


def cpu_baseline(sample_outer=32, budget_s=12.0):
    layout, P, s, dy, out = cpu_sample(sample_outer)
    cpu_oracle_step(P, s, dy, layout, LAMBDA, out)          # warm (page faults, thread pool)
    t0 = time.perf_counter()
    reps = 0
    while True:
        cpu_oracle_step(P, s, dy, layout, LAMBDA, out)
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 200:
            break
    dt = (time.perf_counter() - t0) / reps
    cb = cpu_describe(layout, P.size, dt)
    cb["value"] = (BYTES_FWD + BYTES_BWD) * P.size / dt / 1e9
    return cb


def mnist_cpu(budget_s=6.0):
    """BASELINE.json configs[0], the reference's own CPU-runnable case: one training step of the MNIST model
    (Flatten -> CustomDense(128) -> relu -> CustomDense(10) -> softmax, batch 32, Adam 1e-4, per-tensor scales)
    restated on the oracle (oracle/mnist_step.py), end to end on the host: nested variant and the MaxBin
    loss-term variant.  images/s; a reported baseline."""
    import numpy as np
    from oracle.mnist_step import MnistOracle
    rng = np.random.default_rng(42)
    X = (rng.random((32, 784)) * 255).astype(np.float32)     # the reference feeds un-normalised pixels
    Y = rng.integers(0, 10, 32)
    out = {"unit": "images/s", "kind": "port", "batch": 32, "cores": 1,
           "sample": "oracle/mnist_step.py (numpy fp32; fake-quant fwd, nested / loss-term bwd, Keras Adam, "
                     "MinValueConstraint), steps of batch 32 for ~%.0f s per variant" % (budget_s / 2)}
    for name, variant, value in (("nested", "nested", 1e-10), ("custom_loss_maxbin", "maxbin", 1e-9)):
        m = MnistOracle.init(seed=42, orientation="scalar", variant=variant, value=value, lr=1e-4)
        m.step(X, Y)
        t0, reps = time.perf_counter(), 0
        while time.perf_counter() - t0 < budget_s / 2:
            m.step(X, Y)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        out[name] = {"value": 32 / dt, "ms_per_step": dt * 1e3, "steps": reps}
    out["value"] = out["nested"]["value"]
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its children; this arm is one process and uses every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    sample_outer = LAYOUT[0]      # the WHOLE 64 Mi-element tensor per step: same config as the GPU arm
    layout, P, s, dy, out = cpu_sample(sample_outer)
    cpu_oracle_step(P, s, dy, layout, LAMBDA, out)
    t1 = time.perf_counter()
    cpu_oracle_step(P, s, dy, layout, LAMBDA, out)
    one = time.perf_counter() - t1
    # bounded: the whole --steps/--warmup run stays within ~2 minutes on any host
    steps = max(3, min(args.steps, int(90.0 / max(one, 1e-3))))
    warm = max(1, min(args.warmup, 3))
    for _ in range(warm):
        cpu_oracle_step(P, s, dy, layout, LAMBDA, out)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_oracle_step(P, s, dy, layout, LAMBDA, out)
    dt = time.perf_counter() - t0
    val = (BYTES_FWD + BYTES_BWD) * P.size * steps / dt / 1e9
    cb = cpu_describe(layout, P.size, dt / steps)
    cb["value"] = val
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "layout": list(LAYOUT),
                   "lambda": LAMBDA, "step": "the whole 64Mi-element tensor, forward + nested backward",
                   "note": "reference CPU path = TensorFlow-CPU (not installable); timed stand-in is the op-for-op C port"},
        "same_config": True, "baseline_kind": "c_port",
        "elements_per_s": P.size * steps / dt,
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mnist_cpu": mnist_cpu(),
    }))


# ------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import learned_quantization_b200 as lq
    ops = lq.ops
    if args.only_qat:
        q = run_qat(args, rank, world, local, args.qat_config)
        if rank == 0:
            print(json.dumps(q))
        if world > 1:
            dist.destroy_process_group()
        return

    n = LAYOUT[0] * LAYOUT[1] * LAYOUT[2]
    P_h, s_h, dy_h = synth(LAYOUT, 42 + rank)
    # two input sets, alternated between steps; each array is 256 MiB (> 126 MB L2)
    sets = []
    for k in range(2):
        P = torch.from_numpy(P_h).cuda() if k == 0 else torch.from_numpy(P_h).cuda().roll(4099)
        dy = torch.from_numpy(dy_h).cuda() if k == 0 else torch.from_numpy(dy_h).cuda().roll(977)
        sets.append((P, dy))
    s = torch.from_numpy(s_h).cuda()
    y = torch.empty(n, device="cuda")
    dX = torch.empty(n, device="cuda")
    ds = torch.empty(LAYOUT[1], device="cuda")
    d = ops.make_desc(layout=LAYOUT)
    ws = torch.empty(ops.workspace_bytes([d], lq._lib.LQ_PHASE_BWD), dtype=torch.uint8, device="cuda")
    lib, check = lq._lib.lib, lq._lib.check
    stream = torch.cuda.current_stream().cuda_stream
    ninf, pinf = float("-inf"), float("inf")

    def fwd(P):
        check(lib.lq_fq_fwd(P.data_ptr(), s.data_ptr(), y.data_ptr(), None, 0, *LAYOUT, 0, ninf, pinf, stream))

    # N > 1: ONE tensor [256 * world, 512, 512] sharded along `outer`.  Each rank runs the
    # same single-pass kernels on its shard; the per-group statistics (24 * C bytes per rank) are
    # exchanged and folded in shard order by ONE kernel over NVLink peer memory
    # (lq_fq_bwd_threshold_finish_p2p) on a side stream, overlapping the next forward.  ds is identical
    # on every rank and equals the unsharded result; --exchange nccl uses lq_dp_allgather_f64 +
    # lq_fq_bwd_threshold_finish instead, and the two are compared before timing.
    comm = side = p2p = None
    exchange_note = ""
    if world > 1:
        from learned_quantization_b200.training.dp import NcclComm, P2PFinish
        comm = NcclComm()
        if args.exchange == "p2p":
            try:
                p2p = P2PFinish(max_groups=LAYOUT[1])      # raises on every rank if any rank cannot map its peers
            except lq.LqError as e:
                p2p, exchange_note = None, f" (peer-memory path unavailable: {e})"
        side = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        stats = [torch.empty((LAYOUT[1], 3), dtype=torch.float64, device="cuda") for _ in range(2)]
        gathered = [torch.empty((world, LAYOUT[1], 3), dtype=torch.float64, device="cuda") for _ in range(2)]
        ev_stats = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        n_per_group = float(LAYOUT[0] * world * LAYOUT[2])
        step_no = [0]

    def bwd(P, dy):
        if world == 1:
            check(lib.lq_fq_bwd_threshold(dy.data_ptr(), P.data_ptr(), s.data_ptr(), LAMBDA, dX.data_ptr(), ds.data_ptr(),
                                          *LAYOUT, 0, ninf, pinf, ws.data_ptr(), ws.numel(), stream))
            return
        k = step_no[0] % 2
        step_no[0] += 1
        main.wait_event(ev_done[k])                  # the gather of two steps ago has consumed stats[k]
        check(lib.lq_fq_bwd_threshold_stats(dy.data_ptr(), P.data_ptr(), s.data_ptr(), LAMBDA, dX.data_ptr(),
                                            stats[k].data_ptr(), *LAYOUT, 0, ninf, pinf, ws.data_ptr(), ws.numel(), stream))
        ev_stats[k].record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev_stats[k])
            if p2p is not None:
                p2p.finish(stats[k], LAMBDA, n_per_group, ds)
            else:
                comm.allgather_f64(stats[k], gathered[k])
                ops.fq_bwd_threshold_finish(gathered[k], LAMBDA, n_per_group, out=ds)
            ev_done[k].record(side)

    if p2p is not None:          # the fused peer-memory path must reproduce the NCCL gather path bit for bit
        P0, dy0 = sets[0]
        st0 = torch.empty((LAYOUT[1], 3), dtype=torch.float64, device="cuda")
        check(lib.lq_fq_bwd_threshold_stats(dy0.data_ptr(), P0.data_ptr(), s.data_ptr(), LAMBDA, dX.data_ptr(),
                                            st0.data_ptr(), *LAYOUT, 0, ninf, pinf, ws.data_ptr(), ws.numel(), stream))
        g0 = torch.empty((world, LAYOUT[1], 3), dtype=torch.float64, device="cuda")
        comm.allgather_f64(st0, g0)
        ds_nccl = ops.fq_bwd_threshold_finish(g0, LAMBDA, n_per_group)
        ds_p2p = p2p.finish(st0, LAMBDA, n_per_group, torch.empty_like(ds_nccl))
        torch.cuda.synchronize()
        good = torch.tensor([0 if (p2p.timed_out() or not torch.equal(ds_nccl.view(torch.int32), ds_p2p.view(torch.int32))) else 1],
                            device="cuda")
        dist.all_reduce(good, op=dist.ReduceOp.MIN)
        if int(good) == 0:       # never time a path that does not reproduce the reference exchange
            p2p.close()
            p2p, exchange_note = None, " (peer-memory finish disagreed with the NCCL gather path: not used)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        P, dy = sets[i % 2]
        fwd(P); bwd(P, dy)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    # (1) the timed region: EXACTLY K steps between two events, nothing else on the stream
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_start.record()
    for i in range(args.steps):
        P, dy = sets[i % 2]
        fwd(P)
        bwd(P, dy)
    e_end.record()
    barrier()
    total_ms = e_start.elapsed_time(e_end)
    # (2) per-kernel durations for the roofline: K launches of the forward between two events, then K
    # launches of the backward (element kernel + second stage) between two events.  (An event record
    # after every single kernel costs ~3 us of stream time per record and would be charged to the kernel.)
    ef = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ef[0].record()
    for i in range(args.steps):
        fwd(sets[i % 2][0])
    ef[1].record()
    for i in range(args.steps):
        bwd(*sets[i % 2])
    ef[2].record()
    barrier()
    fwd_ms = ef[0].elapsed_time(ef[1]) / args.steps
    bwd_ms = ef[1].elapsed_time(ef[2]) / args.steps
    instr_ms = fwd_ms + bwd_ms

    # ---- end to end through the host-buffer C-ABI entry (pinned host arrays in, host arrays out)
    hP, hs, hdy = (torch.from_numpy(a).pin_memory() for a in (P_h, s_h, dy_h))
    hy = torch.empty(n).pin_memory()
    hds = torch.empty(LAYOUT[1]).pin_memory()
    hws = torch.empty(ops.host_workspace_bytes(LAYOUT), dtype=torch.uint8, device="cuda")
    e2e_steps = max(3, min(args.steps, 10))
    ops.host_fq_fwd_bwd(hP, hs, hdy, LAMBDA, hy, hds, LAYOUT, hws)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ops.host_fq_fwd_bwd(hP, hs, hdy, LAMBDA, hy, hds, LAYOUT, hws)
    barrier()
    e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    # what the box can do for this traffic pattern with every rank at it at once: the step's H2D (P, dy) and D2H (y)
    # bytes as plain pinned copies on two streams, nothing else
    up, down = torch.cuda.Stream(), torch.cuda.Stream()
    dP_, ddy_ = sets[0]

    def copies():
        with torch.cuda.stream(up):
            dP_.copy_(hP.view_as(dP_), non_blocking=True); ddy_.copy_(hdy.view_as(ddy_), non_blocking=True)
        with torch.cuda.stream(down):
            hy.copy_(y.view_as(hy), non_blocking=True)
    copies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        copies()
    barrier()
    pcie_ms = (time.perf_counter() - t0) / 3 * 1e3
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([total_ms, fwd_ms, bwd_ms, e2e_ms, instr_ms, pcie_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, fwd_ms, bwd_ms, e2e_ms, instr_ms, pcie_ms = t.tolist()
    if rank == 0:
        step_ms = total_ms / args.steps
        value = (BYTES_FWD + BYTES_BWD) * n * world / (step_ms * 1e-3) / 1e9
        peak, peak_src = measured_peak()
        achieved = BYTES_BWD * n / (bwd_ms * 1e-3) / 1e9
        traffic, traffic_note = ncu_traffic()
        cb = cpu_baseline() if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "layout": list(LAYOUT),
                       "elements_per_gpu": n, "lambda": LAMBDA, "l2": "inputs_exceed_l2 (3 x 256 MiB arrays per pass, 2 input sets alternated)",
                       "bytes_per_element": {"fwd": BYTES_FWD, "bwd": BYTES_BWD}, "dX": "written (not aliased to dy)",
                       "sharding": "single device" if world == 1 else
                       f"one tensor [{LAYOUT[0] * world},{LAYOUT[1]},{LAYOUT[2]}] sharded along outer over {world} ranks; "
                       f"per-group statistics ({24 * LAYOUT[1]} B per rank) exchanged per step on a side stream by "
                       + ("one fused peer-memory kernel over NVLink (checked against the NCCL gather before timing)"
                          if p2p is not None else "lq_dp_allgather_f64 + lq_fq_bwd_threshold_finish" + exchange_note)},
            "elements_per_s": n * world / (step_ms * 1e-3),
            "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
            "kernel_timing": "value / ms_per_step: K steps (fwd, bwd alternating) between two events; fwd_ms / bwd_ms / "
                             "roofline: a second pass, K forward launches between two events, then K backward launches "
                             "(element kernel + second stage) between two events",
            "fwd_GBps": BYTES_FWD * n / (fwd_ms * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "kernel": "row_kernel<BwdThresholdOp> (+ finalize_warp_kernel, its second stage)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "frac_of_8TBps_spec": achieved / 8000.0,
                         "traffic": traffic, "traffic_source": traffic_note,
                         "algorithmic_bytes_per_launch": BYTES_BWD * n},
            "e2e": {"value": (BYTES_FWD + BYTES_BWD) * n * world / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
                    "ms_per_step": e2e_ms, "h2d_bytes_per_step": 8 * n + 4 * LAYOUT[1],
                    "d2h_bytes_per_step": 4 * n + 4 * LAYOUT[1],
                    "api": "lq_host_fq_fwd_bwd (pinned host P, s, dy in; host y, ds out; dX is dy itself)",
                    "pcie_bound": {"value": (BYTES_FWD + BYTES_BWD) * n * world / (pcie_ms * 1e-3) / 1e9, "unit": UNIT,
                                   "ms_per_step": pcie_ms,
                                   "what": "the same bytes as bare pinned copies (H2D and D2H on two streams), all ranks at once"},
                    "frac_of_pcie_bound": pcie_ms / e2e_ms,
                    "numa": numa},
            "gpu_launches": (3 if world == 1 else 4) * args.steps,
            "clocks": clocks,
        }
        if cb is not None:
            line["cpu_baseline"] = cb
    if comm is not None:
        torch.cuda.synchronize()
        if p2p is not None:
            if p2p.timed_out():
                raise SystemExit("bench.py: a rank timed out in the peer-memory exchange")
            dist.barrier()
            p2p.close()
        comm.close()
        comm = None
    extra = {}
    if not args.no_qat:
        del sets, y, dX, hP, hdy, hy, hws
        torch.cuda.empty_cache()
        if world > 1:
            extra["dp_selfcheck"] = dp_selfcheck(args, rank, world)
            if not extra["dp_selfcheck"].get("ok"):
                args.qat_comm = "nccl"            # never time an exchange that does not reproduce the single device
        for key in qat_keys(world):
            try:
                extra[key] = run_qat(args, rank, world, local, key)
            except Exception as e:                # reported, never allowed to take the headline line down
                extra[key] = {"error": f"{type(e).__name__}: {e}"}
                if world > 1:
                    raise
        if world == 1:
            extra["mnist_cpu"] = mnist_cpu()
            for key, fn in (("layouts_64Mi", lambda: other_layouts_64mi(args.steps)), ("loss_terms_64Mi", loss_terms_64mi)):
                try:
                    extra[key] = fn()
                except Exception as e:            # reported, never allowed to take the headline line down
                    extra[key] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------- QAT training step
# Every training configuration BASELINE.json names, by key of the JSON line:
QAT_CONFIGS = {
    # IMAGENETTE nested, the reference's own batch (64 per GPU)
    "qat": dict(model="imagenette", variant="nested", batch=64),
    # configs[4]: IMAGENETTE nested at 224x224, batch 1024 sharded over 8 GPUs = 128 per GPU
    "qat_b128": dict(model="imagenette", variant="nested", batch=128),
    # configs[2]: IMAGENETTE custom_loss_terms (MaxBin, gamma 1e-9), data parallel
    "qat_custom_loss": dict(model="imagenette", variant="custom_loss", loss="maxbin", batch=64),
    # configs[1]: CIFAR-10 paper_implementation ResNet-18, per-channel LEARNED scale (threshold rule), batch 256, 1 GPU
    "qat_cifar_paper": dict(model="cifar_paper", sgrad="threshold", batch=256),
    # the same model with the reference's stub backward (PI:62-64: ds = zeros)
    "qat_cifar_paper_ste": dict(model="cifar_paper", sgrad="none", batch=256),
    # configs[0] on the GPU (its CPU figure is `mnist_cpu`): MNIST MLP, nested, per-tensor scales, batch 32
    "qat_mnist": dict(model="mnist", variant="nested", batch=32),
}


def qat_keys(world):
    if world == 1:
        return ["qat", "qat_b128", "qat_custom_loss", "qat_cifar_paper", "qat_cifar_paper_ste", "qat_mnist"]
    return ["qat", "qat_b128", "qat_custom_loss"]          # the data-parallel configurations


def make_comm(args, world):
    from learned_quantization_b200.training.dp import NcclComm, PeerComm
    if world == 1:
        return None, "single device"
    if args.qat_comm == "peer":
        return PeerComm(), "lq_dpx_exchange: fused peer-memory all-reduce + threshold backward (one kernel, in the step's graph)"
    return NcclComm(), "lq_dp_allreduce (ncclAllReduce avg) between two graphs + lq_mt_fq_bwd"


def dp_selfcheck(args, rank, world):
    """Before any data-parallel number is timed: two steps of a small quantised model on `world` ranks, each
    with its shard of a batch, must end where ONE device with the whole batch ends (exact mode), and every
    rank must hold bit-identical parameters.  (tests/test_gpu_dp.py does the same under pytest, but the
    round-end test box has a single GPU.)"""
    import torch
    import torch.distributed as dist
    import learned_quantization_b200 as lq
    from learned_quantization_b200.custom_loss_terms.custom_components.custom_loss_functions import \
        sparse_categorical_crossentropy as scce
    from learned_quantization_b200.training.models import SelfCheckNet
    from learned_quantization_b200.training.step import TrainStep

    B, lr, steps = 2 * world, 1e-3, 3          # step 1 exchanges in one go, steps 2-3 overlapped with the backward
    g = torch.Generator().manual_seed(11)
    x, y = torch.rand(B, 10, 10, 3, generator=g).cuda(), torch.randint(0, 10, (B,), generator=g).cuda()

    def build():
        torch.manual_seed(0)
        net = SelfCheckNet().cuda()
        net(x[:1])
        gs = torch.Generator(device="cuda").manual_seed(3)
        with torch.no_grad():
            for q in net.quant_layers():
                q.scale.copy_(torch.empty_like(q.scale).uniform_(5e-3, 2e-2, generator=gs))
        return net

    out = {"kind": args.qat_comm, "steps": steps, "global_batch": B}
    try:
        comm, _ = make_comm(args, world)
        net = build()
        step = TrainStep(net, scce, optimizer="adam", lr=lr, comm=comm, ds_mode="exact")
        lo, hi = rank * B // world, (rank + 1) * B // world
        for _ in range(steps):
            step(x[lo:hi], y[lo:hi])
        torch.cuda.synchronize()
        timed_out = bool(getattr(comm, "timed_out", lambda: False)())
        ref = build()
        one = TrainStep(ref, scce, optimizer="adam", lr=lr)
        for _ in range(steps):
            one(x, y)
        diff = max(float((a.detach() - b.detach()).abs().max()) for a, b in zip(net.parameters(), ref.parameters()))
        flat = torch.cat([p.detach().flatten() for p in net.parameters()]).view(torch.int32)
        mine = torch.stack([flat.sum(dtype=torch.int64), (flat.to(torch.int64) * torch.arange(1, flat.numel() + 1, device="cuda")).sum()])
        allsums = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allsums, mine)
        identical = all(bool(torch.equal(a, allsums[0])) for a in allsums)
        ok = torch.tensor([1 if (diff <= 0.02 * lr and identical and not timed_out) else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        out.update(ok=bool(int(ok)), max_abs_diff_vs_single_device=diff, tolerance=0.02 * lr,
                   overlapped_pieces=[len(getattr(step, "_early", [])), len(getattr(step, "_late", []))],
                   ranks_bit_identical=identical, timed_out=timed_out)
        comm.close()
    except lq.LqError as e:                       # raised on every rank together (PeerComm.allocate is collective)
        out.update(ok=False, error=str(e))
    return out


def run_qat(args, rank, world, local, key="qat"):
    """Second half of BASELINE.json's metric: QAT images/s of one training configuration (QAT_CONFIGS), synthetic data.
    imagenette: ResNet-style model (20 quantised convs + biases = 40 quantised tensors, channelwise scales, Adam 1e-4,
    224x224x3; nested: lambda 1e-11 with the threshold backward, custom_loss: gamma 1e-9 MaxBin with the STE backward).
    cifar_paper: CIFAR-10/paper_implementation ResNet-18 on CustomConv2DLayerNoBias + quantised fc (22 tensors,
    channelwise scales), 32x32x3, batch 256, SGD(0.1, momentum 0.9, nesterov, weight decay 5e-4) (experiment.py:130-141).
    mnist: Flatten -> Dense128 -> relu -> Dense10 -> softmax, per-tensor scales, batch 32, Adam 1e-4.
    Data parallel: batch sharded, gradients exchanged in EXACT mode (training/dp.py).  Weak scaling."""
    import torch
    import torch.distributed as dist
    import learned_quantization_b200 as lq
    from learned_quantization_b200.custom_loss_terms.custom_components.custom_loss_functions import \
        sparse_categorical_crossentropy as scce
    from learned_quantization_b200.training import models
    from learned_quantization_b200.training.step import TrainStep

    cfg = dict(QAT_CONFIGS[key])
    if key == "qat":                                  # the debugging flags act on the default configuration
        cfg.update(batch=args.qat_batch, variant=args.qat_variant, loss=args.qat_loss)
    torch.backends.cudnn.allow_tf32 = True           # TF 2.11 enables TF32 convolutions by default too
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    model, batch = cfg["model"], cfg["batch"]
    cifar, mnist = model == "cifar_paper", model == "mnist"
    shape = (32, 32, 3) if cifar else ((28, 28, 1) if mnist else (224, 224, 3))
    torch.manual_seed(42)
    variant = "paper" if cifar else cfg["variant"]
    threshold = cfg.get("sgrad", "threshold" if variant == "nested" else "none") == "threshold"
    # lambda (threshold rule) / gamma (custom loss): the thesis' operating points (chapter4.tex:106-113,225-226)
    value = (1e-10 if mnist else 1e-11) if variant in ("nested", "paper") else 1e-9
    if cifar:
        from learned_quantization_b200.paper_implementation.custom_components import custom_layers as pi
        pi.set_scale_grad_mode("threshold" if threshold else "none")
        net = models.cifar_resnet18(value if threshold else 0.0, seed=42, orientation="channelwise").cuda().train()
        pi.set_scale_grad_mode("none")
    elif mnist:
        net = models.mnist_mlp(value, seed=42, orientation="scalar", variant=variant).cuda().train()
    else:
        net = models.imagenette_resnet(value, seed=42, orientation="channelwise", variant=variant,
                                       kernel_storage=args.qat_storage).cuda().train()
    x, y = models.synthetic_batch(batch, shape, "cuda", seed=42 + rank)
    if cifar:
        x = x / 255.0
    with torch.no_grad():
        net(x[:2])                                    # builds the layers
    loss_fn = scce
    if variant == "custom_loss":                      # IMAGENETTE/custom_loss_terms/experiment.py:764-793
        import tempfile
        from learned_quantization_b200.custom_loss_terms.custom_components import custom_loss_functions as lf
        from learned_quantization_b200.training.step import quantised_pairs
        hosts = []
        for h, _, _, _ in quantised_pairs(net):
            if h not in hosts:
                hosts.append(h)
        cls = {"maxbin": lf.SCCEMaxBin, "difference": lf.SCCEDifference, "inverse": lf.SCCEInverse}[cfg["loss"]]
        loss_obj = cls(layers=hosts, penalty_rate=value, log_dir=tempfile.mkdtemp(prefix="lq_bench_"))
        loss_obj.log_to_files = False                 # the reference's three tf.print(file://) per step: not timed here
        loss_fn = loss_obj.compute_total_loss
    if cifar:
        loss_fn = lambda yt, logits: scce(yt, torch.softmax(logits, dim=-1))
        opt = dict(optimizer="sgd", lr=0.1, momentum=0.9, nesterov=True, weight_decay=5e-4)
    else:
        opt = dict(optimizer="adam", lr=1e-4)
    comm, comm_note = make_comm(args, world)
    try:
        step = TrainStep(net, loss_fn, comm=comm, ds_mode="exact", overlap=not args.qat_no_overlap,
                         late_blocks=args.qat_late_blocks, **opt)
    except lq.LqError as e:                            # peer memory unavailable: raised on every rank together
        if world == 1 or args.qat_comm != "peer":
            raise
        args.qat_comm = "nccl"
        comm, comm_note = make_comm(args, world)
        comm_note += f" (peer-memory exchange unavailable: {e})"
        step = TrainStep(net, loss_fn, comm=comm, ds_mode="exact", **opt)
    hx = x.cpu().pin_memory()
    hy = y.cpu().pin_memory()
    if args.qat_skip_exchange and world > 1:          # diagnostic only: N independent replicas, nothing exchanged
        def local_only(scale_backward=None):
            with torch.no_grad():
                (scale_backward or step._scale_backward)()
        step._exchange = local_only
        step._early = []
        comm_note = "DIAGNOSTIC: no gradient exchange (independent replicas)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warm = args.qat_steps, max(3, args.qat_warmup)
    if args.qat_eager:
        run = step
        for _ in range(warm):
            step(x, y)
    else:
        step.capture(x, y, warmup=warm)               # whole step in one CUDA graph
        run = step.replay
        for _ in range(3):
            run(x, y)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run(x, y)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1) / steps
    qat_clocks = sampler.stop() if rank == 0 else None
    # end to end: every step's batch comes from pinned host memory and every step's loss is read back.
    # The H2D copy of batch i+1 is issued on a copy stream BEFORE step i is launched, so it overlaps step i
    # (a data loader's prefetch); each step still pays for exactly one batch copy inside the timed region.
    copy_stream = torch.cuda.Stream()
    stage = [(torch.empty_like(x), torch.empty_like(y)) for _ in range(2)]
    main = torch.cuda.current_stream()

    def prefetch(i):
        copy_stream.wait_stream(main)            # the staging buffer's previous consumer has been enqueued on main
        with torch.cuda.stream(copy_stream):
            stage[i % 2][0].copy_(hx, non_blocking=True)
            stage[i % 2][1].copy_(hy, non_blocking=True)

    barrier()
    t0 = time.perf_counter()
    prefetch(0)
    for i in range(steps):
        main.wait_stream(copy_stream)            # batch i has landed
        sx, sy = stage[i % 2]
        if args.qat_eager:
            if i + 1 < steps:
                prefetch(i + 1)
            loss = step(sx, sy)
        else:
            step._sx.copy_(sx); step._sy.copy_(sy)        # device-to-device into the graph's static inputs
            if i + 1 < steps:
                prefetch(i + 1)                           # enqueued before the step: overlaps it
            loss = step._graph_replay_only()
        loss_host = float(loss)
    barrier()
    e2e_ms = (time.perf_counter() - t0) / steps * 1e3
    # N > 1, default configuration only: the same captured step WITHOUT any exchange (N independent replicas running
    # at the same time on the same box) -- what separates the cost of the exchange from what eight busy GPUs in one
    # chassis cost each other (clocks, power), which no exchange design can change
    replica_ms = 0.0
    if world > 1 and key == "qat" and not args.qat_eager and not args.qat_skip_exchange:
        real_exchange = step._exchange
        early = getattr(step, "_early", None)

        def local_only(scale_backward=None):
            with torch.no_grad():
                (scale_backward or step._scale_backward)()
        step._exchange = local_only
        if early is not None:
            step._early = []
        step.capture(x, y, warmup=3)
        for _ in range(3):
            step.replay(x, y)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(steps):
            step.replay(x, y)
        r1.record()
        barrier()
        replica_ms = r0.elapsed_time(r1) / steps
        step._exchange = real_exchange
        if early is not None:
            step._early = early
    # the gradient exchange alone (N > 1): K launches between two events on the step's stream
    exch_ms = 0.0
    if world > 1:
        for _ in range(3):
            step._exchange()
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        for _ in range(steps):
            step._exchange()
        x1.record()
        barrier()
        exch_ms = x0.elapsed_time(x1) / steps
    t = torch.tensor([dev_ms, e2e_ms, exch_ms, replica_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, exch_ms, replica_ms = t.tolist()
    n_q = sum(p.numel() for p in step.bank.params)
    name = {"imagenette": f"imagenette_resnet_{variant}_channelwise_224" + (f"_{cfg['loss']}" if variant == "custom_loss" else ""),
            "cifar_paper": "cifar10_paper_resnet18_channelwise_32", "mnist": "mnist_mlp_nested_per_tensor"}[model]
    peer = getattr(step, "_peer", False)
    out = {"metric": "qat_images_per_s", "value": batch * world / (dev_ms * 1e-3), "unit": "images/s",
           "ms_per_step": dev_ms, "steps": steps, "warmup": warm,
           "e2e": {"value": batch * world / (e2e_ms * 1e-3), "unit": "images/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": hx.numel() * 4 + hy.numel() * 8, "d2h_bytes_per_step": 4},
           "config": {"workload": name, "per_gpu_batch": batch, "global_batch": batch * world,
                      "optimizer": "sgd(0.1, 0.9, nesterov, weight decay 5e-4) fused with MinValueConstraint" if cifar else
                      "adam(1e-4) fused with l2 + MinValueConstraint",
                      ("lambda" if variant in ("nested", "paper") else "gamma"): value if (threshold or variant == "custom_loss") else 0.0,
                      "quantised_tensors": len(step.bank.params), "quantised_elements": n_q,
                      "scale_grad_mode": ("threshold (NQ:62-118) on the gradient averaged over all ranks (exact mode)" if threshold
                                          else "none (STE: dX = dy, ds = zeros, CL:61-62 / PI:62-64)"),
                      "conv_math": "cuDNN TF32 (TF 2.11 default), fp32 storage", "parallelism": f"dp{world}",
                      "gradient_exchange": comm_note,
                      "kernel_storage": "hwio" if (cifar or mnist) else
                      args.qat_storage + (" (logical HWIO, cuDNN channels-last order in memory)" if args.qat_storage == "ohwi" else ""),
                      "launch": "eager" if args.qat_eager else
                      ("one CUDA graph per step" if (world == 1 or peer) else "three CUDA graphs per step around the NCCL call")},
           "allreduce_floats": step.bucket.n_exact,
           "our_launches_per_step": {("lq_mt_fq_fwd_penalty (forward + loss term in one pass)" if getattr(step, "fused_loss", False)
                                      else "lq_mt_fq_fwd"): 1,
                                     "lq_penalty_fwd": 1 if (variant == "custom_loss" and not getattr(step, "fused_loss", False)) else 0,
                                     "lq_penalty_bwd": 1 if variant == "custom_loss" else 0,
                                     "lq_mt_fq_bwd": (2 if threshold else 0) if not peer else 0,
                                     "lq_dpx_exchange": (2 if threshold else 1) if peer else 0,
                                     ("lq_opt_sgd" if cifar else "lq_opt_adam"): -(-len(step.table.params) // 56),
                                     "lq_dp_allreduce": 1 if (world > 1 and not peer) else 0},
           "last_loss": loss_host, "clocks": qat_clocks}
    if replica_ms:
        out["independent_replicas"] = {"ms_per_step": replica_ms, "value": batch * world / (replica_ms * 1e-3),
                                       "dp_over_replicas": replica_ms / dev_ms,
                                       "what": "the same step on the same N GPUs at the same time with the exchange removed"}
    if world > 1:
        nbytes = step.bucket.n_exact * 4
        out["exchange"] = {"ms": exch_ms, "bytes": nbytes,
                           "busbw_GBps": 2.0 * (world - 1) / world * nbytes / (exch_ms * 1e-3) / 1e9,
                           "what": "gradient exchange + scale backward alone, K launches between two events" if peer else
                                   "ncclAllReduce + lq_mt_fq_bwd alone, K launches between two events"}
        if peer and comm.timed_out():
            raise SystemExit("bench.py: a rank timed out in the peer-memory gradient exchange")
    if comm is not None:
        barrier()
        comm.close()
    return out


# ------------------------------------------------------------------------- layout sweep
def sweep_layouts():
    """BASELINE configs[3] / SURVEY 8d: per-tensor and per-channel layouts, 1 Mi .. 1 Gi elements."""
    out = []
    for lg in (20, 22, 24, 26, 28, 30):
        n = 1 << lg
        out.append(("per_tensor", (1, 1, n)))
        for c in (64, 256, 512, 2048):
            if n // (512 * c) >= 1:
                out.append((f"hwio_like_C{c}", (n // (512 * c), c, 512)))
            out.append((f"rows_C{c}", (1, c, n // c)))
            out.append((f"interleaved_C{c}", (n // c, c, 1)))
    return [(name, layout, "trained", LAMBDA) for name, layout in out] + sweep_variants()


def sweep_variants():
    """SURVEY 8d's other operating points, at the headline size only: scales at their initial value
    100*eps (codes ~ +-2e4, nothing below the threshold) and the reference's own lambdas (0, 1e-11, 1e-10:
    every element above the threshold, the all-above branch NQ:77-79 / 103-105)."""
    n = 1 << 26
    base = [("per_tensor", (1, 1, n)), ("hwio_like_C512", (n // (512 * 512), 512, 512)), ("interleaved_C512", (n // 512, 512, 1))]
    out = [(name, layout, "init", LAMBDA) for name, layout in base]
    for lam in (0.0, 1e-11, 1e-10):
        out += [(name, layout, "trained", lam) for name, layout in base]
    return out


def time_layout(name, layout, scales, lam, steps, warmup, gen, peak):
    """fwd / bwd device times of one layout (median over five batches of `steps` launches between two events, or a CUDA
    graph of `steps` launches for the launch-bound sizes); arrays rotate so that consecutive launches exceed L2."""
    import torch
    import learned_quantization_b200 as lq
    ops, lib, check = lq.ops, lq._lib.lib, lq._lib.check
    ninf, pinf = float("-inf"), float("inf")
    n = layout[0] * layout[1] * layout[2]
    nsets = max(2, min(16, -(-(600 << 20) // (12 * n))))
    if scales == "trained":
        s = torch.empty(layout[1], device="cuda").uniform_(5e-3, 2e-2, generator=gen)
    else:
        s = torch.full((layout[1],), 1.1920928955078125e-05, device="cuda")        # NQ:156-158
    Ps = [torch.empty(n, device="cuda").normal_(0, 0.05, generator=gen) for _ in range(nsets)]
    dys = [torch.empty(n, device="cuda").normal_(0, 1e-3, generator=gen) for _ in range(nsets)]
    outs = [torch.empty(n, device="cuda") for _ in range(nsets)]
    ds = torch.empty(layout[1], device="cuda")
    ws = torch.empty(max(ops.workspace_bytes([ops.make_desc(layout=layout)], lq._lib.LQ_PHASE_BWD), 256),
                     dtype=torch.uint8, device="cuda")

    def fwd(i):
        check(lib.lq_fq_fwd(Ps[i].data_ptr(), s.data_ptr(), outs[i].data_ptr(), None, 0, *layout, 0, ninf, pinf,
                            torch.cuda.current_stream().cuda_stream))

    def bwd(i):
        check(lib.lq_fq_bwd_threshold(dys[i].data_ptr(), Ps[i].data_ptr(), s.data_ptr(), lam, outs[i].data_ptr(),
                                      ds.data_ptr(), *layout, 0, ninf, pinf, ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream().cuda_stream))

    res = {}
    graphed = n <= (1 << 24)      # launch-bound sizes: time a CUDA graph of `steps` launches
    for tag, fn in (("fwd", fwd), ("bwd", bwd)):
        for i in range(max(warmup, 3)):
            fn(i % nsets)
        torch.cuda.synchronize()
        if graphed:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(steps):
                    fn(i % nsets)
            g.replay()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); g.replay(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) / steps)
            res[tag] = sorted(ts)[len(ts) // 2]
            del g
        else:
            # like the headline: `steps` launches between two events (an event record costs ~3 us of stream
            # time, which a per-launch bracket would charge to every launch); median of five such batches
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
            k = 0
            for a, b in evs:
                a.record()
                for _ in range(steps):
                    fn(k % nsets); k += 1
                b.record()
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) / steps for a, b in evs)
            res[tag] = ts[len(ts) // 2]
    gb = lambda bpe, ms: bpe * n / (ms * 1e-3) / 1e9
    out = {"layout_name": name, "layout": list(layout), "elements": n, "scales": scales, "lambda": lam,
           "rotating_sets": nsets,
           "timing": f"CUDA graph of {steps} launches" if graphed else f"{steps} launches between two CUDA events, median of 5",
           "fwd_us": res["fwd"] * 1e3, "bwd_us": res["bwd"] * 1e3,
           "fwd_GBps": gb(BYTES_FWD, res["fwd"]), "bwd_GBps": gb(BYTES_BWD, res["bwd"]),
           "fwd_bwd_GBps": (BYTES_FWD + BYTES_BWD) * n / ((res["fwd"] + res["bwd"]) * 1e-3) / 1e9,
           "frac_of_measured_peak": (BYTES_FWD + BYTES_BWD) * n / ((res["fwd"] + res["bwd"]) * 1e-3) / 1e9 / peak}
    del Ps, dys, outs
    torch.cuda.empty_cache()
    return out


def run_sweep(args):
    """Not the headline: one JSON line per (layout, size) with fwd/bwd device times, written to stdout."""
    import torch
    torch.cuda.set_device(0)
    peak, _ = measured_peak()
    gen = torch.Generator(device="cuda").manual_seed(42)
    only = set(args.only.split(",")) if args.only else None
    for name, layout, scales, lam in sweep_layouts():
        n = layout[0] * layout[1] * layout[2]
        if n > args.max_elems or (only and name not in only):
            continue
        print(json.dumps(time_layout(name, layout, scales, lam, args.steps, args.warmup, gen, peak)), flush=True)


def other_layouts_64mi(steps):
    """Extra key of the default line (N = 1): the other two traversal regimes at the headline size -- per-tensor scale
    (ROW, one group) and the interleaved layout (COL: register forward, TMA-staged backward) -- beside the headline's
    channelwise [256, 512, 512].  Same timing method as the headline."""
    import torch
    peak, _ = measured_peak()
    gen = torch.Generator(device="cuda").manual_seed(42)
    n = 1 << 26
    out = {}
    for name, layout in (("per_tensor", (1, 1, n)), ("interleaved_C512", (n // 512, 512, 1))):
        r = time_layout(name, layout, "trained", LAMBDA, steps, 3, gen, peak)
        out[name] = {k: r[k] for k in ("layout", "fwd_us", "bwd_us", "fwd_bwd_GBps", "frac_of_measured_peak", "timing")}
    return out


def loss_terms_64mi():
    """Extra key of the default line (N = 1): the custom-loss reductions (LF:75-116 / 161-195) on the three 64 Mi
    layouts, forward (4 B/elem: read P) and backward (8 B/elem: read P, write dX); GPU time per call from a CUDA graph
    of ten calls, fraction of the measured copy peak."""
    import torch
    import learned_quantization_b200 as lq
    from learned_quantization_b200 import _lib
    ops = lq.ops
    peak, _ = measured_peak()
    gen = torch.Generator(device="cuda").manual_seed(3)

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / reps * 1e3)
        return sorted(ts)[1]

    out = {}
    n = 1 << 26
    for name, shape, sshape in (("channelwise_256x512x512", (256, 512, 512), (1, 512, 1)), ("per_tensor", (n,), (1,)),
                                ("interleaved_131072x512", (131072, 512), (1, 512))):
        P = [torch.empty(shape, device="cuda").normal_(0, 0.05, generator=gen)]
        s = [torch.empty(sshape, device="cuda").uniform_(5e-3, 2e-2, generator=gen)]
        dX, dS = [torch.empty_like(P[0])], [torch.empty_like(s[0])]
        for kname, kind in (("maxbin", _lib.LQ_PEN_MAXBIN), ("difference", _lib.LQ_PEN_DIFFERENCE)):
            _, ws = ops.penalty_fwd(kind, P, s)
            f_us = timed(lambda: ops.penalty_fwd(kind, P, s))
            b_us = timed(lambda: ops.penalty_bwd(kind, P, s, ws, 1.0, dXs=dX, dss=dS))
            out[f"{name}/{kname}"] = {"fwd_us": f_us, "bwd_us": b_us,
                                      "fwd_frac_of_measured_peak": 4 * n / (f_us * 1e-6) / 1e9 / peak,
                                      "bwd_frac_of_measured_peak": 8 * n / (b_us * 1e-6) / 1e9 / peak}
        del P, dX
        torch.cuda.empty_cache()
    return out


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL writes its version
    banner to fd 1 when NCCL_DEBUG is set in the environment): point fd 1 at stderr for the run and
    keep a private handle on the real stdout for the JSON lines."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w", buffering=1)
    os.dup2(2, 1)
    sys.stdout = real


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-qat", action="store_true", help="skip the QAT images/s section of the line")
    ap.add_argument("--only-qat", action="store_true", help="debug: run only the QAT section and print it")
    ap.add_argument("--qat-config", default="qat", choices=sorted(QAT_CONFIGS),
                    help="with --only-qat: which training configuration to time")
    ap.add_argument("--qat-skip-exchange", action="store_true", help="diagnostic: N > 1 without any gradient exchange")
    ap.add_argument("--qat-no-overlap", action="store_true", help="N > 1: exchange all gradients after the backward")
    ap.add_argument("--qat-late-blocks", type=int, default=16,
                    help="N > 1: grid of the exchange that is started while the backward is still running")
    ap.add_argument("--qat-comm", default="peer", choices=["peer", "nccl"],
                    help="N > 1: gradient exchange of the QAT step (fused peer-memory kernel, or ncclAllReduce)")
    ap.add_argument("--qat-batch", type=int, default=64, help="per-GPU batch of the QAT step (reference: 64)")
    ap.add_argument("--qat-steps", type=int, default=20)
    ap.add_argument("--qat-warmup", type=int, default=5)
    ap.add_argument("--qat-variant", default="nested", choices=["nested", "custom_loss"])
    ap.add_argument("--qat-loss", default="maxbin", choices=["maxbin", "difference", "inverse"])
    ap.add_argument("--qat-storage", default="ohwi", choices=["hwio", "ohwi"],
                    help="memory order of the conv kernels (logical shape is HWIO either way)")
    ap.add_argument("--qat-eager", action="store_true", help="do not capture the QAT step in a CUDA graph")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the shards' statistics are exchanged (fused peer-memory kernel, or NCCL all-gather)")
    ap.add_argument("--sweep", action="store_true", help="layout/size sweep (one JSON line each) instead of the headline")
    ap.add_argument("--max-elems", type=int, default=1 << 28)
    ap.add_argument("--only", default="", help="comma-separated layout names for --sweep")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.sweep:
        run_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
