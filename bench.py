#!/usr/bin/env python
"""bench.py -- PFPN head action-samples/s (fwd+bwd) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm

One "step" = one pass of the fused head (log_prob + entropy + PPO surrogate, forward
and backward) over one minibatch of B=65536 states
per GPU, A=36, P=35 -- the configuration BASELINE.json's target is quoted on (c2's
B=4096 working set, 42 MB, is L2-resident and therefore only a parity / latency case).
Prints ONE JSON line on rank 0.

Besides the headline the line carries (all measured in this run, outside the timed region):
  roofline      the K1 call (head_kernel + its [A,P] finalize) against the measured HBM peak
  e2e           the same metric through the host-buffer API, H2D / D2H inside the timed region
  cpu_baseline  the oracle (torch-CPU restatement of the TF-1.14 graph) at 1 thread and on all cores
  xcheck        N > 1: the peer-memory exchanges against NCCL, replicas bit-identical (asserted)
  dppo_update   BASELINE c4: the DPPO minibatch update, B_total = 65536 sharded over the N GPUs
  extra         BASELINE c2 / c3 / c5: the named secondary shapes at this run's N
"""
from __future__ import annotations

import argparse
import hashlib
import importlib.util
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_PER_GPU, A, P = 65536, 36, 35
SEED = 28949
METRIC = "pfpn_head_action_samples_per_s_fwd_bwd"
UNIT = "action-samples/s"
ALG_BYTES_PER_STATE = 8 * A * P + 4 * A + 16  # logits in, dlogits out, action, adv+lp_old, lp+ent
WORKLOAD = f"PFPN head fwd+bwd (log_prob+entropy+PPO surrogate), B={B_PER_GPU}/GPU, A={A}, P={P}, fp32"
K1_SOURCE = os.path.join(ROOT, "pfpn_b200", "csrc", "head_logprob.cu")


def static_config(world: int) -> dict:
    """The workload description, identical in both arms."""
    return {"workload": WORKLOAD, "B_per_gpu": B_PER_GPU, "A": A, "P": P, "n_gpus": world,
            "l2": "inputs larger than L2: 330 MB logits in + 330 MB dlogits out per step vs 126 MB L2"}


def load_synth():
    """pfpn_b200/synth.py loaded BY PATH: importing the package would dlopen libpfpn_b200.so, which the reference arm
    must not do (it times the CPU restatement only)."""
    spec = importlib.util.spec_from_file_location("_pfpn_synth", os.path.join(ROOT, "pfpn_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def file_sha256(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of the head kernel at this exact shape
    (profiles/head_kernel_ncu.json, written by tools/make_profiles.py together with the sha256 of the kernel source it was
    captured on).  A capture of a DIFFERENT kernel source is stale: traffic is then null and the reason is stated."""
    try:
        with open(os.path.join(ROOT, "profiles", "head_kernel_ncu.json")) as f:
            d = json.load(f)
        cur = file_sha256(K1_SOURCE)
        if d.get("source_sha256") != cur:
            return None, f"stale: captured on head_logprob.cu {str(d.get('source_sha256'))[:12]}, current {cur[:12]}"
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]), f"ncu --set full, {d.get('round')}, head_logprob.cu {cur[:12]}"
    except Exception as e:  # noqa: BLE001
        return None, f"no capture ({e.__class__.__name__})"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "10"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------- CPU legs ----
def cpu_reference_inputs(b: int):
    synth = load_synth()
    d = synth.head_inputs(b, A, P, seed=SEED, far_frac=0.0)
    d["lp_old"] = torch.zeros(b)
    return d


def cpu_reference_pass(d):
    """One pass of the oracle: the op-order-faithful torch-CPU fp32 restatement of the reference's TF graph for this path
    (every [B,A,P] intermediate materialised as the graph does; TF 1.14 cannot be installed).  Checker code -- it is only
    TIMED here, as the baseline."""
    from oracle import head as oracle_head
    return oracle_head.ppo_head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], d["adv"], d["lp_old"],
                                        dtype=torch.float32)


def cpu_reference_rate(d, reps: int, threads: int):
    torch.set_num_threads(threads)
    cpu_reference_pass(d)  # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        cpu_reference_pass(d)
    dt = (time.perf_counter() - t0) / reps
    return d["logits"].shape[0] / dt, dt


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """The reference arm: the reference's CPU implementation of the path (TF 1.14 is not installable: the oracle port) on
    all host threads, the FULL workload of the GPU arm per step (65536 states).  Imports neither pfpn_b200 nor its .so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_cores()
    torch.set_num_threads(threads)
    d = cpu_reference_inputs(B_PER_GPU)
    for _ in range(max(0, args.warmup)):
        cpu_reference_pass(d)
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        cpu_reference_pass(d)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    rate = B_PER_GPU / dt
    sample = (f"oracle (torch-CPU fp32 op-order restatement of the TF-1.14 graph; TF itself cannot be installed), the full "
              f"{B_PER_GPU} states per step, {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": static_config(args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "loaded_pfpn_so": any("libpfpn_b200" in l for l in open("/proc/self/maps")) if os.path.exists("/proc/self/maps") else None,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- helpers ----
def timed(fn, n, stream, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record(stream)
    for i in range(n):
        fn()
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ts[n // 2], sum(ts) / n


def tensor_hash(t: torch.Tensor) -> torch.Tensor:
    bits = t.detach().contiguous().reshape(-1).view(torch.int32).to(torch.int64)
    w = torch.arange(1, bits.numel() + 1, device=bits.device, dtype=torch.int64) % 65521
    return torch.stack([bits.sum(), (bits * w).sum()])


def numa_note(local_rank: int) -> str:
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        rows = [l for l in out.splitlines() if l.startswith(f"GPU{local_rank}\t") or l.startswith(f"GPU{local_rank} ")]
        hdr = [l for l in out.splitlines() if "NUMA Affinity" in l]
        if rows and hdr:
            cols = hdr[0].split("\t")
            vals = rows[0].split("\t")
            pick = {c.strip(): v.strip() for c, v in zip(cols, vals) if c.strip() in ("CPU Affinity", "NUMA Affinity")}
            return ", ".join(f"{k} {v}" for k, v in pick.items())
    except Exception:
        pass
    return "unknown"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dppo", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--dppo-steps", type=int, default=10)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pfpn_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import ctypes as C
    from pfpn_b200 import _cabi, head, synth
    from pfpn_b200.host import HostHeadPipeline

    host_binding, all_cpus = None, os.sched_getaffinity(0)
    if os.environ.get("PFPN_BIND_NUMA", "1") != "0":
        from pfpn_b200.host import bind_host_to_gpu
        host_binding = bind_host_to_gpu(local_rank)  # before any pinned allocation: first touch decides the NUMA node

    def maxr(x: float) -> float:
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- synthetic minibatch shard, resident in HBM ---------------------------------
    B = B_PER_GPU
    g = torch.Generator(device="cuda")
    g.manual_seed(SEED + rank)
    logits = torch.randn(B, A, P, device=dev, generator=g) * 2.0
    loc, logstd = synth.particle_grid(A, P, torch.Generator().manual_seed(SEED))
    loc, logstd = loc.to(dev), logstd.to(dev)
    idx = torch.multinomial(torch.softmax(logits.view(B * A, P), -1), 1, generator=g).view(B, A)
    value = (loc.expand(B, A, P).gather(2, idx[..., None]) +
             logstd.exp().expand(B, A, P).gather(2, idx[..., None]) * torch.randn(B, A, 1, device=dev, generator=g))[..., 0].contiguous()
    adv = torch.randn(B, device=dev, generator=g)
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    lp, ent, dlogits, dloc, dlogstd, loss, stats = f(B), f(B), f(B, A, P), f(A, P), f(A, P), f(1), f(2)
    ws = torch.empty(head.head_workspace_bytes(A, P), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    a = _cabi.HeadArgs()
    a.logits, a.loc, a.logstd, a.value = logits.data_ptr(), loc.data_ptr(), logstd.data_ptr(), value.data_ptr()
    a.B, a.A, a.P, a.mode, a.flags = B, A, P, _cabi.HEAD_FWD, 0
    a.lp, a.ent = lp.data_ptr(), ent.data_ptr()
    _cabi.check(_cabi.pfpn_head_logprob(a, ws.data_ptr(), ws.numel(), stream.cuda_stream))
    lp_old = (lp + 0.05 * torch.randn(B, device=dev, generator=g)).contiguous()
    a.mode = _cabi.HEAD_PPO
    a.adv, a.lp_old, a.adv_stats = adv.data_ptr(), lp_old.data_ptr(), stats.data_ptr()
    a.eps_clip, a.loss_scale = 0.2, 1.0 / (B * world)
    a.dlogits, a.dloc, a.dlogstd, a.loss = dlogits.data_ptr(), dloc.data_ptr(), dlogstd.data_ptr(), loss.data_ptr()
    flat_small = torch.empty(2, A, P, device=dev)
    use_peer = world > 1 and os.environ.get("PFPN_FUSED_ALLREDUCE", "1") != "0"
    gather = None
    if use_peer:
        from pfpn_b200.peer import PeerGather
        try:
            gather = PeerGather(2 * A * P, dev)
        except (RuntimeError, ValueError):  # collective outcome: every rank falls back together
            use_peer = False

    def step(mode="full"):
        """mode (N > 1, diagnostic breakdown only): "none" = no exchange at all, "push" = K1 pushes but nobody sums."""
        _cabi.check(_cabi.pfpn_adv_stats(adv.data_ptr(), B, stats.data_ptr(), stream.cuda_stream))
        if use_peer and mode != "none":
            # the path's only exchange: [2, A, P] particle gradients.  K1's finalize kernel PUSHES them into
            # every rank's gather buffer as 8-byte {value, sequence} packets (no fence / flag round trip behind the data).
            # The sum is produced ONE exchange late (it feeds the optimizer, not the next head launch) by the finalize
            # kernel of the NEXT step: the packets arrived a step ago, nothing waits, no kernel of its own.
            if mode == "push":
                pa = gather.push_args()
                gather.consumed = gather.pushed  # (diagnostic: the rows are never read)
            else:
                pa = gather.push_args(consume_into=flat_small.view(-1))  # the same launch sums the PREVIOUS exchange
            _cabi.check(_cabi.pfpn_head_logprob_push(a, ws.data_ptr(), ws.numel(), pa, stream.cuda_stream))
        else:
            _cabi.check(_cabi.pfpn_head_logprob(a, ws.data_ptr(), ws.numel(), stream.cuda_stream))
            if world > 1 and mode != "none":
                flat_small[0].copy_(dloc)
                flat_small[1].copy_(dlogstd)
                dist.all_reduce(flat_small)
    # adv_stats, K1, K1 finalize (which also pushes and sums the previous exchange at N > 1); NCCL fallback: + 2 copies + all-reduce
    launches_per_step = 3 if (use_peer or world == 1) else 6

    def drain():
        while use_peer and gather.pending > 0:  # the last exchange(s): consumed inside the timed region
            gather.reduce(flat_small.view(-1), 1.0, stream.cuda_stream)

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for i in range(args.steps):
        step()
        if i == args.steps - 1:
            drain()  # (inside the timed region: every exchange is consumed before the closing event)
        evs[i + 1].record(stream)
    barrier()
    per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps))
    clocks = sampler.stop() if sampler else None  # (sampled during the timed region only: the poller must not run under the other legs)
    local_ms = evs[0].elapsed_time(evs[-1])
    total_ms = maxr(local_ms)
    rank_ms, breakdown = None, None
    if world > 1:
        # where the weak-scaling loss comes from: every rank's own step time (the line reports the MAX), and the same loop
        # without the exchange / with the push only (diagnostic, outside the timed region above)
        t_all = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(t_all, torch.tensor([local_ms / args.steps], device=dev, dtype=torch.float64))
        rank_ms = [round(float(t.item()), 5) for t in t_all]
        breakdown = {}
        for mode in ("none", "push"):
            for _ in range(5):
                step(mode)
            barrier()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record(stream)
            for _ in range(args.steps):
                step(mode)
            b1.record(stream)
            barrier()
            breakdown[f"ms_per_step_{mode}"] = round(maxr(b0.elapsed_time(b1)) / args.steps, 5)
        breakdown["note"] = "max over ranks; none = K1 alone on every rank (GPU-to-GPU spread only), push = + the peer stores and flags"
    value_rate = world * B * args.steps / (total_ms * 1e-3)

    # ---- N > 1: the exchange against NCCL, in the driver's record (outside every timed region) -------------------
    xcheck = None
    if world > 1:
        xcheck = {}
        step()
        torch.cuda.synchronize()
        drain()
        torch.cuda.synchronize()
        mine = torch.stack([dloc, dlogstd]).clone() if not use_peer else gather.row(gather.pushed, rank).reshape(2, A, P).clone()
        allc = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        ordered = allc[0].clone()
        for r in range(1, world):
            ordered += allc[r]  # rank-ordered fp32 sum: what the kernel promises, bit for bit
        nccl = mine.clone()
        dist.all_reduce(nccl)
        xcheck["exchange"] = "push (K1 finalize -> peers' gather rows, 8-byte {value, sequence} packets), summed by the next finalize / pfpn_peer_gather_sum_packets" if use_peer else "NCCL all_reduce"
        xcheck["peer_sum_bit_equal_rank_ordered"] = bool(torch.equal(flat_small, ordered)) if use_peer else None
        xcheck["peer_sum_max_abs_vs_nccl"] = float((flat_small - nccl).abs().max())
        xcheck["peer_sum_ref_max_abs"] = float(nccl.abs().max())
        h = tensor_hash(flat_small)
        hs = [torch.empty_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        xcheck["peer_sum_replicas_bit_identical"] = all(torch.equal(x, hs[0]) for x in hs)

    # ---- dominant kernel alone (head_kernel + its [A,P] finalize), for the roofline ---
    kev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
    _cabi.check(_cabi.pfpn_adv_stats(adv.data_ptr(), B, stats.data_ptr(), stream.cuda_stream))
    a.dloc, a.dlogstd = dloc.data_ptr(), dlogstd.data_ptr()
    for i in range(args.steps):
        kev[2 * i].record(stream)
        _cabi.check(_cabi.pfpn_head_logprob(a, ws.data_ptr(), ws.numel(), stream.cuda_stream))
        kev[2 * i + 1].record(stream)
    torch.cuda.synchronize()
    kms = sorted(kev[2 * i].elapsed_time(kev[2 * i + 1]) for i in range(args.steps))
    k_avg_ms = sum(kms) / len(kms)

    # ---- end to end through the host-buffer API -----------------------------------------
    pipe = HostHeadPipeline(B, A, P, dev)
    pin = lambda t_: t_.detach().cpu().pin_memory()
    h = dict(logits=pin(logits), loc=pin(loc), logstd=pin(logstd), value=pin(value), adv=pin(adv), lp_old=pin(lp_old))
    for _ in range(3):
        pipe.run(**h)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.e2e_steps):
        out = pipe.run(**h)
    e1.record(stream)
    barrier()
    e2e_ms_local = e0.elapsed_time(e1)
    e2e_ms = maxr(e2e_ms_local)
    e2e_rate = world * B * args.e2e_steps / (e2e_ms * 1e-3)
    # sanity: the host path and the resident path agree
    assert torch.allclose(out["lp"], lp.cpu(), rtol=0, atol=0), "host pipeline lp mismatch"
    e2e_step_s = e2e_ms_local * 1e-3 / args.e2e_steps
    os.sched_setaffinity(0, all_cpus)  # (the CPU baseline below uses every core the process was given)

    # ---- the DPPO minibatch update around the head (BASELINE c4): B_total = 65536 sharded ------
    dppo = None
    if not args.no_dppo:
        from pfpn_b200.learner import SyncReplicasAdam, shard_bounds
        from pfpn_b200.network import ParticleFilteringClipPPONetwork
        lo, hi = shard_bounds(B_PER_GPU, rank, world)
        Bs = hi - lo
        net = ParticleFilteringClipPPONetwork(True, [197], [A], action_lower_bound=[-1.0] * A, action_upper_bound=[1.0] * A,
                                              particles=P, resample=-1, resample_interval=368, normalize_state=True,
                                              clip_state=5.0, normalize_advantage=True, device=dev, seed=SEED).init()
        opt = SyncReplicasAdam(lr=1e-4, norm_clip=1.0)
        st_ = torch.randn(Bs, 197, device=dev, generator=g)
        ac_, lp_, v_ = net.run_batch(st_)
        lpo_ = lp_ + 0.05 * torch.randn(Bs, device=dev, generator=g)
        adv_ = torch.randn(Bs, device=dev, generator=g)

        from pfpn_b200.learner import GraphedUpdate

        def time_updates(fn, n):
            barrier()
            u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            u0.record(stream)
            for _ in range(n):
                fn()
            u1.record(stream)
            barrier()
            return maxr(u0.elapsed_time(u1) / n)

        def upd_eager():
            net.compute_gradients(st_, ac_, v_, lpo_, adv_)
            opt.apply_gradients(net)
        for _ in range(3):
            upd_eager()
        # the shard size decides the default mode (learner.GraphedUpdate): small shards replay ONE captured CUDA graph,
        # large ones stay eager.  The default mode is timed FIRST (the second measurement of a back-to-back pair runs
        # ~10 % slower at 65536 states: the part settles at its power cap), the other one is reported beside it.
        graph_max = int(os.environ.get("PFPN_GRAPH_MAX_BATCH", "24576"))
        um_graph = None
        if Bs <= graph_max:
            gu = GraphedUpdate(net, opt, Bs, warmup=0)
            gu._set(st_, ac_, v_, lpo_, adv_)
            for _ in range(3):
                gu.run()  # capture + first replays
            um_graph = time_updates(gu.run, args.dppo_steps)
            um_eager = time_updates(upd_eager, args.dppo_steps)
            um = um_graph
        else:
            um_eager = time_updates(upd_eager, args.dppo_steps)
            um = um_eager
        dppo = {"workload": f"DPPO minibatch update, B_total={B_PER_GPU} sharded over {world} GPU(s): 197-1024-512 actor+critic trunk (tcgen05 3xTF32 GEMMs, critic on a parallel graph branch), PFPN head, local clip -> staged bucket -> rank-ordered mean over NVLink peer memory -> Adam: 3 launches" + ("; whole update replayed as ONE CUDA graph" if um_graph is not None else "; eager multi-stream issue (shard above the graph-replay threshold)"),
                "ms_per_update": um, "samples_per_s": B_PER_GPU / (um * 1e-3), "scaling": "strong",
                "trunk_tflops": 12.6e6 * B_PER_GPU / (um * 1e-3) / 1e12,
                "ms_per_update_eager": um_eager, "ms_per_update_graph": um_graph, "optimizer_chain_launches": getattr(opt, "launches_last_step", None),
                "exchange": opt._mode + (f", {'two' if world >= int(os.environ.get('PFPN_PEER_TWO_PHASE_MIN', '6')) else 'one'}-phase peer kernel" if world > 1 else "")}
        if world > 1:
            # one more update from IDENTICAL state through each exchange implementation: the round-1 chain with an NCCL
            # all-reduce, and the 3-launch step with the one-phase and the two-phase (reduce-scatter + all-gather) kernel
            snap_n, snap_o = net.state_dict(), opt.state_dict()
            results = {}
            for name, fused, sync_step, two_phase_min in (("nccl", False, "0", "99"), ("peer_one_phase", True, "1", "99"),
                                                          ("peer_two_phase", True, "1", "2")):
                if fused and not opt.fused_peer:
                    continue
                net.load_state_dict(snap_n)
                o2 = SyncReplicasAdam(lr=1e-4, norm_clip=1.0, fused_peer=fused)
                o2._lazy(net)
                o2.load_state_dict(snap_o)
                if fused:
                    o2._peers = opt._peers  # same peer-mapped buffers (their call counter carries on)
                os.environ["PFPN_PEER_TWO_PHASE_MIN"] = two_phase_min
                os.environ["PFPN_SYNC_STEP"] = sync_step
                net.compute_gradients(st_, ac_, v_, lpo_, adv_)
                o2.apply_gradients(net)
                torch.cuda.synchronize()
                results[name] = (net.params.clone(), o2.m.clone(), o2.v.clone(), net.state_mean.clone())
            os.environ.pop("PFPN_PEER_TWO_PHASE_MIN", None)
            os.environ.pop("PFPN_SYNC_STEP", None)
            ref_p = results["nccl"][0]
            for name, (p_, m_, v2_, sm_) in results.items():
                hh = torch.cat([tensor_hash(p_), tensor_hash(m_), tensor_hash(v2_), tensor_hash(sm_)])
                hs = [torch.empty_like(hh) for _ in range(world)]
                dist.all_gather(hs, hh)
                xcheck[f"{name}_replica_hash_equal"] = all(torch.equal(x, hs[0]) for x in hs)
                if name != "nccl":
                    d_ = (p_ - ref_p).abs().max() / (ref_p - snap_n["params"]).abs().max().clamp_min(1e-30)
                    xcheck[f"{name}_vs_nccl_rel_update_diff"] = float(d_)
                    xcheck[f"{name}_vs_nccl_state_mean_max_abs"] = float((sm_ - results["nccl"][3]).abs().max())
    # ---- BASELINE c2 / c3 / c5 at this run's N ---------------------------------------------------------------------
    extra = None
    if not args.no_extra:
        extra = run_extra(dev, rank, world, stream, maxr)

    if xcheck is not None:
        ok = xcheck["peer_sum_replicas_bit_identical"] and xcheck["peer_sum_max_abs_vs_nccl"] <= 1e-6 * max(1.0, xcheck["peer_sum_ref_max_abs"])
        if xcheck.get("peer_sum_bit_equal_rank_ordered") is not None:
            ok = ok and xcheck["peer_sum_bit_equal_rank_ordered"]
        for k, v in xcheck.items():
            if k.endswith("_replica_hash_equal"):
                ok = ok and v
            if k.endswith("_vs_nccl_rel_update_diff"):
                ok = ok and v < 1e-3  # (Adam normalises by |g|: an ulp of the averaged gradient moves near-zero entries)
        xcheck["ok"] = bool(ok)

    if rank == 0:
        peak, peak_src = peaks()
        achieved = ALG_BYTES_PER_STATE * B / (k_avg_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic()
        cfg = static_config(world)
        cfg["parallelism"] = f"dp{world} (states sharded, [2,A,P] particle-gradient exchange" + \
            (" pushed by K1's finalize kernel over NVLink peer memory)" if use_peer else ", NCCL all-reduce)" if world > 1 else ")")
        line = {
            "metric": METRIC, "value": value_rate, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "step_ms_median": per[len(per) // 2], "rank_ms_per_step": rank_ms, "exchange_breakdown": breakdown,
            "clocks": clocks,
            "e2e": {"value": e2e_rate, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
                    "d2h_bytes_per_step": pipe.d2h_bytes, "steps": args.e2e_steps,
                    "api": "pfpn_b200.host.HostHeadPipeline.run (pinned host buffers, 3-stream chunked)",
                    "rank0_h2d_GBps": pipe.h2d_bytes / e2e_step_s / 1e9, "rank0_d2h_GBps": pipe.d2h_bytes / e2e_step_s / 1e9,
                    "bound": "PCIe: both directions stream concurrently for the whole step; the 0.17 ms kernel is ~2 % of it",
                    "rank0_host_affinity": numa_note(local_rank), "rank0_host_binding": host_binding},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": "pfpn::head_kernel<.., KM=PPO> (+head_finalize)",
                         "alg_bytes_per_state": ALG_BYTES_PER_STATE, "kernel_ms_avg": k_avg_ms,
                         "kernel_ms_median": kms[len(kms) // 2]},
        }
        if dppo is not None:
            line["dppo_update"] = dppo
        if xcheck is not None:
            line["xcheck"] = xcheck
        if extra is not None:
            line["extra"] = extra
        if not args.no_cpu_baseline and world == 1:
            cores = host_cores()
            b_s = 8192
            d = cpu_reference_inputs(b_s)
            r1, dt1 = cpu_reference_rate(d, 2, 1)
            rN, dtN = cpu_reference_rate(d, 4, cores)
            line["cpu_baseline"] = {"value": rN, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"oracle torch-CPU fp32 op-order restatement, {b_s} of {B} states per pass, 4 passes "
                                              f"({dtN*1e3:.0f} ms each) on {cores} threads; 2 passes ({dt1*1e3:.0f} ms each) on 1 thread",
                                    "single_thread": {"value": r1, "unit": UNIT, "cores": 1,
                                                      "note": "benchmark.sh:18 pins OPENBLAS_NUM_THREADS=1 per worker"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if xcheck is not None and not xcheck["ok"]:
        raise SystemExit(f"xcheck failed: {xcheck}")


def run_extra(dev, rank, world, stream, maxr):
    """The secondary shapes BASELINE.json names, at this run's N (each rank its contiguous shard, no collective):
    c2 head at B=4096 (L2-resident: a latency case), c3 resampler sweep P=10/35/100 (replicas only, rank 0's time),
    c5 SAC head at B=1M total, A=36, P=100 against the HBM roofline of SURVEY 8(d)."""
    import numpy as np
    from pfpn_b200 import _cabi, head, resampling, sampling, synth
    peak, _ = peaks()
    out = {}
    # ---- c2 ----
    B2, A2, P2 = 4096, 36, 35
    d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synth.head_inputs(B2, A2, P2).items()}
    o = head.head_call(_cabi.HEAD_FWD, d["logits"], d["loc"], d["logstd"], d["value"])
    lp_old = o["lp"] + d["lp_noise"]
    st2 = head.adv_stats(d["adv"])
    buf = {}
    ms, _ = timed(lambda: head.head_call(_cabi.HEAD_PPO, d["logits"], d["loc"], d["logstd"], d["value"], adv=d["adv"],
                                         lp_old=lp_old, adv_stats_t=st2, out=buf), 30, stream)
    ms_f, _ = timed(lambda: head.head_call(_cabi.HEAD_FWD, d["logits"], d["loc"], d["logstd"], d["value"], out=buf), 30, stream)
    ms_s, _ = timed(lambda: sampling.sample_plain(d["logits"], d["loc"], d["logstd"], seed=1, offset=2), 30, stream)
    out["c2_head_B4096_per_gpu"] = {"ppo_fwd_bwd_us": maxr(ms) * 1e3, "fwd_only_us": maxr(ms_f) * 1e3, "plain_sample_us": maxr(ms_s) * 1e3,
                                    "Mstates_s_all_gpus": world * B2 / maxr(ms) / 1e3,
                                    "note": "41.9 MB working set is L2-resident: a latency case, not a roofline case"}
    # ---- rollout side at the headline shape: one fused pass vs the three-kernel form (K2 + K1 forward + K4) ----
    Br = B_PER_GPU
    gr = torch.Generator(device="cuda")
    gr.manual_seed(SEED + 7 + rank)
    lgr = torch.randn(Br, A, P, device=dev, generator=gr) * 2.0
    locr, lsr = (x.to(dev) for x in synth.particle_grid(A, P, torch.Generator().manual_seed(SEED)))
    mxa, sma = torch.zeros(A, P, device=dev), torch.zeros(A, P, device=dev)
    t_s, _ = timed(lambda: sampling.sample_plain(lgr, locr, lsr, seed=1, offset=2), 10, stream)
    act_r, _ = sampling.sample_plain(lgr, locr, lsr, seed=1, offset=2)
    bufr = {}
    t_f, _ = timed(lambda: head.head_call(_cabi.HEAD_FWD, lgr, locr, lsr, act_r, out=bufr), 10, stream)
    t_k, _ = timed(lambda: sampling.stats_update(lgr, mxa, sma), 10, stream)
    t_z, _ = timed(lambda: sampling.rollout_fused(lgr, locr, lsr, seed=1, offset=2, max_active=mxa, sum_active=sma), 10, stream)
    ro_bytes = (4 * A * P + 8 * A + 8) * Br
    out["rollout_head_B65536_per_gpu"] = {
        "fused_one_pass_us": maxr(t_z) * 1e3, "frac_of_hbm_peak": ro_bytes / (maxr(t_z) * 1e-3) / 1e9 / peak,
        "alg_bytes_per_state": 4 * A * P + 8 * A + 8,
        "three_kernel_form_us": {"sample": maxr(t_s) * 1e3, "log_prob_fwd": maxr(t_f) * 1e3, "stats": maxr(t_k) * 1e3},
        "note": "sample + log_prob + activity statistics of a batched rollout step (ppo.py:56-62, a2c.py:346-365)"}
    del lgr
    # ---- c3 ----
    res = {}
    for P3 in (10, 35, 100):
        rng = np.random.default_rng(33406 + P3)
        A3, H = 36, 512
        lg = rng.normal(0, 3, (512, A3, P3))
        pr = np.exp(lg - lg.max(-1, keepdims=True))
        pr /= pr.sum(-1, keepdims=True)
        mx, sm = pr.max(0).astype(np.float32), pr.sum(0).astype(np.float32)
        dead = rng.random((A3, P3)) < 0.1
        mx[dead] = 1e-6
        t = lambda a_: torch.tensor(a_, dtype=torch.float32, device=dev)
        base = dict(mx=t(mx), sm=t(sm), loc=t(np.linspace(-1, 1, P3)[None].repeat(A3, 0)), ls=t(np.full((A3, P3), np.log(2 / (P3 - 1)))),
                    b=t(rng.normal(0, 1, A3 * P3)), W=t(rng.normal(0, .01, (H, A3 * P3))))
        work = {k: v.clone() for k, v in base.items()}

        def copies():
            for k in work:
                work[k].copy_(base[k])

        def call():
            copies()
            resampling.resample_(work["mx"], work["sm"], work["loc"], work["ls"], work["b"], work["W"], seed=3, offset=5)
        us = (timed(call, 20, stream)[0] - timed(copies, 20, stream)[0]) * 1e3
        res[f"P{P3}"] = {"dead": int(dead.sum()), "us_per_resample": us}
    out["c3_resample_A36_H512"] = dict(res, note="replicas only (every rank runs it identically); latency-bound, roofline fraction not meaningful")
    # ---- c5 ----
    A5, P5, BT = 36, 100, 1_000_000
    B5 = BT // world
    g5 = torch.Generator(device="cuda")
    g5.manual_seed(12831 + rank)
    logits5 = torch.randn(B5, A5, P5, device=dev, generator=g5) * 2.0
    from pfpn_b200.network import initial_particles
    loc5, ls5 = (x.to(dev) for x in initial_particles(A5, P5, True))
    g_s = torch.randn(B5, A5, device=dev, generator=g5)
    g_lp = torch.full((B5,), 1.0 / BT, device=dev)
    smp, s_pre, _ = sampling.rsample_fwd(logits5, loc5, ls5, seed=7, offset=0)
    buf5 = dict(dlogits=torch.empty_like(logits5))
    t_f, _ = timed(lambda: sampling.rsample_fwd(logits5, loc5, ls5, seed=7, offset=0), 5, stream, warm=1)
    t_l, _ = timed(lambda: head.head_call(_cabi.HEAD_GRAD, logits5, loc5, ls5, s_pre, tanh=True, g_lp=g_lp, want_dvalue=True, out=buf5), 5, stream, warm=1)
    t_b, _ = timed(lambda: sampling.rsample_bwd(logits5, loc5, ls5, g_s, buf5["dvalue"], seed=7, offset=0), 3, stream, warm=1)
    tot = maxr(t_f + t_l + t_b)
    alg_state = 8 * A5 * P5 + 12 * A5 + 8
    roof_states = peak * 1e9 / alg_state
    c5 = {"B_total": BT, "B_per_gpu": B5, "A": A5, "P": P5,
          "rsample_fwd_ms": maxr(t_f), "tanh_logprob_fwd_bwd_ms": maxr(t_l), "rsample_bwd_ms": maxr(t_b),
          "variant": "split (boundary-faithful: rsample fwd, tanh log_prob fwd+bwd with dvalue, rsample bwd = 3 launches)",
          "Mstates_s_all_gpus": BT / tot / 1e3,
          "frac_of_8d_roofline": (BT / (tot * 1e-3)) / (world * roof_states),
          "roofline_Mstates_s_per_gpu": roof_states / 1e6, "alg_bytes_per_state": alg_state}
    try:
        fz = {}
        t_z, _ = timed(lambda: sampling.sac_head_fused(logits5, loc5, ls5, g_s, g_lp, seed=7, offset=0, out=fz), 3, stream, warm=1)
        tz = maxr(t_z)
        c5["fused_fwd_bwd_ms"] = tz
        c5["fused_Mstates_s_all_gpus"] = BT / tz / 1e3
        c5["fused_frac_of_8d_roofline"] = (BT / (tz * 1e-3)) / (world * roof_states)
    except AttributeError:
        pass
    out["c5_sac_head"] = c5
    del logits5, buf5
    torch.cuda.empty_cache()
    # ---- SAC-PFPN learner step at the reference's batch size (deepmimic_sac_base.py:8), eager and as one replayed graph ----
    t_e = t_g = -1.0
    sac_err, sac_replays = None, 0
    # every rank its own learner: a one-rank group each (new_group is collective: all ranks create all groups, same order)
    solo = None
    if world > 1:
        solo = [torch.distributed.new_group([r]) for r in range(world)][rank]
    try:  # (local work only inside the try: the collectives below must run on every rank whatever happens here)
        from pfpn_b200.sac import GraphedSACUpdate, ParticleFilteringSACNetwork, SACOptimizer, ReplayRing
        Ss, Bs_ = 197, 256
        snet = ParticleFilteringSACNetwork(True, [Ss], [A], action_lower_bound=[-1.0] * A, action_upper_bound=[1.0] * A,
                                           particles=35, resample=-1, resample_interval=12000, normalize_state=True,
                                           clip_state=5.0, device=dev, seed=SEED).init()
        sopt = SACOptimizer(group=solo)
        ring = ReplayRing(200_000, Ss, A, device=dev, seed=SEED + rank)
        n0 = 50_000
        ring.append(torch.randn(n0, Ss, device=dev, generator=gr), torch.rand(n0, A, device=dev, generator=gr) * 2 - 1,
                    torch.randn(n0, device=dev, generator=gr), torch.ones(n0, device=dev), torch.randn(n0, Ss, device=dev, generator=gr))

        def sac_eager():
            snet.compute_gradients(*ring.sample(Bs_))
            sopt.apply_gradients(snet)
        t_e, _ = timed(sac_eager, 20, stream)
        sgu = GraphedSACUpdate(snet, sopt, Bs_, warmup=1)
        t_g, _ = timed(lambda: sgu.run(*ring.sample(Bs_)), 20, stream, warm=4)
        sac_replays = sgu.replays
        del snet, sopt, ring, sgu
    except Exception as e:  # noqa: BLE001 -- an extra must never take the headline down
        sac_err = repr(e)[:200]
    t_e, t_g = maxr(t_e), maxr(t_g)
    out["sac_step_B256_P35"] = {"eager_ms": t_e, "graph_ms": t_g, "graph_replays": sac_replays,
                                "samples_s_per_gpu_graph": (256 / (t_g * 1e-3)) if t_g > 0 else None, "error": sac_err,
                                "note": "replicas only (every rank its own independent learner); replay sample -> two actor "
                                        "forwards, six critic evaluations, head fwd/bwd, joint clip, two Adams, target sync"}
    # ---- K6: the trunk's tensor-core GEMM at the headline batch, against the MEASURED tensor peak -------------------
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mp = json.load(f)
        bf16_burst, bf16_sust = float(mp["bf16_tflops"]), float(mp.get("bf16_tflops_sustained", mp["bf16_tflops"]))
        peak_src = "MEASURED_PEAKS.json: cuBLAS bf16 / 2 (tf32 runs at half the bf16 rate)"
    except Exception:  # noqa: BLE001
        bf16_burst, bf16_sust, peak_src = 2250.0, 2250.0, "fallback: nominal 2.25 PFLOP/s bf16 / 2"
    Mg, Ng, Kg = B_PER_GPU, 512, 1024
    Ag = torch.randn(Mg, Kg, device=dev, generator=gr)
    Wg = torch.randn(Kg, Ng, device=dev, generator=gr) * 0.05
    Wlo = torch.empty_like(Wg)
    bg = torch.zeros(Ng, device=dev)
    Cg = torch.empty(Mg, Ng, device=dev)
    _cabi.check(_cabi.pfpn_split_lo(Wg.data_ptr(), Wlo.data_ptr(), Wg.numel(), stream.cuda_stream))
    t_g, _ = timed(lambda: _cabi.check(_cabi.pfpn_tc_gemm_nn_lo(Ag.data_ptr(), Kg, Wg.data_ptr(), Wlo.data_ptr(), Ng, Cg.data_ptr(), Ng,
                                                                  bg.data_ptr(), None, Ng, Mg, Ng, Kg, 2, stream.cuda_stream)), 20, stream)
    t_g = maxr(t_g)
    fp32_tf = 2.0 * Mg * Ng * Kg / (t_g * 1e-3) / 1e12
    out["k6_trunk_gemm"] = {"shape": f"M={Mg} (per GPU), N={Ng}, K={Kg}, bias+relu6 (layer 2 forward)", "ms": t_g,
                            "fp32_equivalent_TFLOPs": fp32_tf, "tf32_mma_TFLOPs": 3.0 * fp32_tf,
                            "tf32_peak_burst": bf16_burst / 2, "tf32_peak_sustained": bf16_sust / 2,
                            "frac_of_burst_peak": 3.0 * fp32_tf / (bf16_burst / 2), "frac_of_sustained_peak": 3.0 * fp32_tf / (bf16_sust / 2),
                            "peak_source": peak_src,
                            "note": "3xTF32 error-compensated: three tcgen05.mma (kind::tf32, cta_group::2) per fp32 product; "
                                    "bound = tensor pipe under the power cap"}
    del Ag, Wg, Wlo, Cg
    return out


if __name__ == "__main__":
    main()
