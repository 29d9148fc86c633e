"""profiles/<round>_scaling.md from the bench lines of the N = 1, 2, 4, 8 runs (gpurun_out/bench_n<N>.json)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
lines = {}
for n in (1, 2, 4, 8):
    p = os.path.join(ROOT, "gpurun_out", f"bench_n{n}.json")
    if os.path.exists(p):
        ls = [l for l in open(p).read().splitlines() if l.startswith("{")]
        if ls:
            lines[n] = json.loads(ls[-1])
if not lines:
    raise SystemExit("no bench_n*.json under gpurun_out/")
json.dump(lines, open(os.path.join(ROOT, "profiles", f"{rnd}_bench_lines.json"), "w"), indent=1)
base = lines.get(1)
out = [f"# {rnd}: scaling of `bench.py` over 1 / 2 / 4 / 8 B200 (builder-run; the driver's SCALE record is the judged one)\n",
       "Head step = adv-stats + fused head fwd+bwd (PPO) over 65536 states PER GPU (weak scaling); the `[2,A,P]` exchange is pushed by "
       "K1's finalize kernel as {value, sequence} packets and summed by the next step's finalize kernel.  DPPO update = c4, 65536 states in total sharded over the GPUs "
       "(strong scaling).  c5 = fused SAC head, 10^6 states in total.\n",
       "| N | head ms/step | head samples/s | weak eff. | K1 roofline frac | DPPO ms/update | strong eff. | c5 fused M states/s | c5 frac of 8(d) roofline | e2e samples/s | xcheck |",
       "|---|---|---|---|---|---|---|---|---|---|---|"]
for n, d in sorted(lines.items()):
    we = base["ms_per_step"] / d["ms_per_step"] if base else float("nan")
    dp = d.get("dppo_update", {})
    se = (base["dppo_update"]["ms_per_update"] / (n * dp["ms_per_update"])) if base and dp else float("nan")
    c5 = d.get("extra", {}).get("c5_sac_head", {})
    xc = d.get("xcheck")
    out.append(f"| {n} | {d['ms_per_step']:.4f} | {d['value']/1e6:.1f} M | {we:.3f} | {d['roofline']['frac']:.3f} | "
               f"{dp.get('ms_per_update', float('nan')):.3f} ({dp.get('exchange', '')}) | {se:.2f} | "
               f"{c5.get('fused_Mstates_s_all_gpus', float('nan')):.1f} | {c5.get('fused_frac_of_8d_roofline', float('nan')):.3f} | "
               f"{d['e2e']['value']/1e6:.2f} M | {'ok' if (xc or {}).get('ok') else ('-' if xc is None else 'FAIL')} |")
open(os.path.join(ROOT, "profiles", f"{rnd}_scaling.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
