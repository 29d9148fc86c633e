#!/usr/bin/env python
"""bench.py -- PFPN head action-samples/s (fwd+bwd) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm

One "step" = one pass of the fused head (log_prob + entropy + PPO surrogate, forward
and backward) over one minibatch of B=65536 states
per GPU, A=36, P=35 -- the configuration BASELINE.json's target is quoted on (c2's
B=4096 working set, 42 MB, is L2-resident and therefore only a parity case).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
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


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of the head kernel at this
    exact shape (profiles/head_kernel_ncu.json, written by tools/make_profiles.py); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "head_kernel_ncu.json")) as f:
            d = json.load(f)
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
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


def cpu_reference_rate(b_sample: int, reps: int, threads: int):
    """Times the oracle (op-order-faithful torch-CPU fp32 restatement of the reference
    graph; TF 1.14 cannot be installed) on `b_sample` states.  Checker code, timed only
    as the baseline."""
    from oracle import head as oracle_head
    from pfpn_b200 import synth
    torch.set_num_threads(threads)
    d = synth.head_inputs(b_sample, A, P, seed=SEED, far_frac=0.0)
    lp_old = torch.zeros(b_sample)
    oracle_head.ppo_head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], d["adv"], lp_old,
                                 dtype=torch.float32)  # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle_head.ppo_head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], d["adv"], lp_old,
                                     dtype=torch.float32)
    dt = (time.perf_counter() - t0) / reps
    return b_sample / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    b_sample = 8192
    # warm-up steps are real oracle passes too
    for _ in range(max(0, args.warmup - 1)):
        cpu_reference_rate(b_sample, 1, threads)
    rate, dt = cpu_reference_rate(b_sample, max(1, args.steps), threads)
    sample = f"oracle (torch-CPU fp32 op-order restatement of the TF-1.14 graph), {b_sample} of {B_PER_GPU} states per step, {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "B_per_gpu": B_PER_GPU, "A": A, "P": P},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dppo", action="store_true")
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
    if use_peer:
        from pfpn_b200.peer import PeerSum
        psum = PeerSum(2 * A * P, dev)

    def step():
        _cabi.check(_cabi.pfpn_adv_stats(adv.data_ptr(), B, stats.data_ptr(), stream.cuda_stream))
        if use_peer:
            # the path's only exchange: [2, A, P] particle gradients.  K1's finalize kernel writes
            # them straight into this rank's peer-visible staging slot; one kernel signals, waits and sums.
            slot = psum.slot()
            a.dloc, a.dlogstd = slot.data_ptr(), slot.data_ptr() + 4 * A * P
        _cabi.check(_cabi.pfpn_head_logprob(a, ws.data_ptr(), ws.numel(), stream.cuda_stream))
        if use_peer:
            psum.reduce(flat_small.view(-1), 1.0, stream.cuda_stream)
        elif world > 1:
            flat_small[0].copy_(dloc)
            flat_small[1].copy_(dlogstd)
            dist.all_reduce(flat_small)
    launches_per_step = 3 + (1 if use_peer else 0)

    def barrier():
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
        evs[i + 1].record(stream)
    barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value_rate = world * B * args.steps / (total_ms * 1e-3)

    # ---- dominant kernel alone (head_kernel + its [A,P] finalize), for the roofline ---
    kev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
    _cabi.check(_cabi.pfpn_adv_stats(adv.data_ptr(), B, stats.data_ptr(), stream.cuda_stream))
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
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_rate = world * B * args.e2e_steps / (float(e2e_ms.item()) * 1e-3)
    # sanity: the host path and the resident path agree
    assert torch.allclose(out["lp"], lp.cpu(), rtol=0, atol=0), "host pipeline lp mismatch"

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

        def upd():
            net.compute_gradients(st_, ac_, v_, lpo_, adv_)
            opt.apply_gradients(net)
        for _ in range(3):
            upd()
        barrier()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record(stream)
        for _ in range(args.dppo_steps):
            upd()
        u1.record(stream)
        barrier()
        um = torch.tensor([u0.elapsed_time(u1) / args.dppo_steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(um, op=dist.ReduceOp.MAX)
        dppo = {"workload": f"DPPO minibatch update, B_total={B_PER_GPU} sharded over {world} GPU(s): 197-1024-512 actor+critic trunk (tcgen05 3xTF32 GEMMs), PFPN head, local clip, all-reduce of the 8.4 MB bucket fused into Adam over NVLink peer memory (N>1)",
                "ms_per_update": float(um.item()), "samples_per_s": B_PER_GPU / (float(um.item()) * 1e-3), "scaling": "strong",
                "trunk_tflops": 12.6e6 * B_PER_GPU / (float(um.item()) * 1e-3) / 1e12}
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peak, peak_src = peaks()
        achieved = ALG_BYTES_PER_STATE * B / (k_avg_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value_rate, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "B_per_gpu": B, "A": A, "P": P,
                       "l2": "inputs larger than L2: 330 MB logits in + 330 MB dlogits out per step vs 126 MB L2",
                       "step_ms_median": per[len(per) // 2], "parallelism": f"dp{world} (states sharded, [2,A,P] particle-gradient exchange" + (" over NVLink peer memory, one kernel)" if use_peer else ", NCCL all-reduce)" if world > 1 else ")")},
            "clocks": clocks,
            "e2e": {"value": e2e_rate, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
                    "d2h_bytes_per_step": pipe.d2h_bytes, "steps": args.e2e_steps,
                    "api": "pfpn_b200.host.HostHeadPipeline.run (pinned host buffers, 3-stream chunked)"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "peak_source": peak_src, "kernel": "pfpn::head_kernel<.., KM=PPO> (+head_finalize)",
                         "alg_bytes_per_state": ALG_BYTES_PER_STATE, "kernel_ms_avg": k_avg_ms,
                         "kernel_ms_median": kms[len(kms) // 2]},
        }
        if dppo is not None:
            line["dppo_update"] = dppo
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            rate, dt = cpu_reference_rate(4096, 3, cores)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"oracle torch-CPU fp32 op-order restatement, 4096 states x 3 reps ({dt*1e3:.0f} ms each), {cores} threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
