"""Turn the ncu outputs of a gpurun call into the tracked summaries under profiles/.

  python tools/make_profiles.py r01          # reads gpurun_out/{launches.csv, head_full.ncu-rep}
Writes profiles/<round>_launches.csv (per-launch device time of `bench.py --steps 2 --warmup 1`),
profiles/<round>_launch_shares.md, profiles/<round>_head_kernel.md and profiles/head_kernel_ncu.json
(the dram bytes bench.py reports as roofline.traffic)."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles"); os.makedirs(out, exist_ok=True)
src = os.path.join(ROOT, "gpurun_out")

lp = os.path.join(src, "launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp, errors="replace")) if r and not r[0].startswith("==")]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]; ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    iu = h.index("Metric Unit")
    per = collections.OrderedDict(); launches = []
    for r in rows[hdr + 1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        t = float(r[iv].replace(",", "")); u = r[iu]
        t_us = t / 1000.0 if u in ("ns", "nsecond") else (t * 1000.0 if u in ("ms", "msecond") else t)
        name = r[ik].split("(")[0]
        launches.append((name, t_us)); per.setdefault(name, []).append(t_us)
    with open(os.path.join(out, f"{rnd}_launches.csv"), "w") as f:
        f.write("launch,kernel,gpu_time_us\n")
        for i, (n, t) in enumerate(launches):
            f.write(f"{i},{n},{t:.3f}\n")
    tot = sum(t for _, t in launches)
    with open(os.path.join(out, f"{rnd}_launch_shares.md"), "w") as f:
        f.write(f"# {rnd}: ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline`\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n\n"
                "| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for n, ts in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{n[:90]}` | {len(ts)} | {sum(ts):.1f} | {100*sum(ts)/tot:.1f}% |\n")
    print("launch list:", len(launches), "launches")

rep = os.path.join(src, "head_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units, vals = rows[0], rows[1], rows[2]
    get = lambda k: vals[hdr.index(k)] if k in hdr else None
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"] + \
           [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")]
    def tobytes(k):
        v = float(get(k).replace(",", "")); u = units[hdr.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    js = {"kernel": get("Kernel Name"), "dram_bytes_read": tobytes("dram__bytes_read.sum"), "dram_bytes_write": tobytes("dram__bytes_write.sum"),
          "gpu_time_us_under_ncu": float(get("gpu__time_duration.sum").replace(",", "")), "round": rnd,
          # bench.py nulls roofline.traffic when the kernel source no longer matches the one this capture was taken on
          "source_sha256": __import__("hashlib").sha256(open(os.path.join(ROOT, "pfpn_b200", "csrc", "head_logprob.cu"), "rb").read()).hexdigest(),
          "command": "ncu --set full --clock-control none --import-source on -k regex:head_kernel -s 6 -c 1 python tools/time_head.py  (B=65536, A=36, P=35, PPO fwd+bwd)"}
    json.dump(js, open(os.path.join(out, "head_kernel_ncu.json"), "w"), indent=1)
    with open(os.path.join(out, f"{rnd}_head_kernel.md"), "w") as f:
        f.write(f"# {rnd}: `ncu --set full` of the K1 head kernel (B=65536, A=36, P=35, PPO fwd+bwd)\n\n| metric | unit | value |\n|---|---|---|\n")
        for k in keys:
            if k in hdr:
                f.write(f"| {k} | {units[hdr.index(k)]} | {get(k)} |\n")
        alg = 10240 * 65536
        f.write(f"\nalgorithmic bytes per launch = 10240 B x 65536 = {alg/1e6:.1f} MB; measured DRAM traffic = "
                f"{(js['dram_bytes_read'] + js['dram_bytes_write'])/1e6:.1f} MB ({(js['dram_bytes_read'] + js['dram_bytes_write'])/alg:.3f}x)\n")
    srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    tmp = os.path.join(src, "head_full_src.csv"); open(tmp, "w").write(srcp)
    mix = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_opmix.py"), tmp], capture_output=True, text=True).stdout
    open(os.path.join(out, f"{rnd}_head_kernel.md"), "a").write("\n## SASS opcode mix (warp-instructions executed) and stall samples\n\n```\n" + mix + "```\n")
    print("head kernel summary written")

# ---- K6 tcgen05 GEMM (tc_full.ncu-rep: one forward-layer launch, M=65536, N=512, K=1024) -----------------------------
rep = os.path.join(src, "tc_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units, vals = rows[0], rows[1], rows[2]
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"] + \
           [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")]
    with open(os.path.join(out, f"{rnd}_tc_gemm_kernel.md"), "w") as f:
        f.write(f"# {rnd}: `ncu --set full` of the K6 tcgen05 3xTF32 GEMM (forward layer, M=65536, N=512, K=1024, bias+relu6)\n\n"
                "| metric | unit | value |\n|---|---|---|\n")
        for k in keys:
            if k in hdr:
                f.write(f"| {k} | {units[hdr.index(k)]} | {vals[hdr.index(k)]} |\n")
    srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    tmp = os.path.join(src, "tc_full_src.csv"); open(tmp, "w").write(srcp)
    mix = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_opmix.py"), tmp], capture_output=True, text=True).stdout
    open(os.path.join(out, f"{rnd}_tc_gemm_kernel.md"), "a").write("\n## SASS opcode mix (UTCHMMA = tcgen05.mma, UTMALDG = TMA load) and stall samples\n\n```\n" + mix + "```\n")
    print("tc gemm summary written")
# ---- generic summaries: K3f (sac_full.ncu-rep) and K2f (rollout_full.ncu-rep) ------------------------------------------
def summarize(rep_name, out_name, title, alg_bytes=None):
    rep = os.path.join(src, rep_name)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units, vals = rows[0], rows[1], rows[2]
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"] + \
           [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")]
    with open(os.path.join(out, out_name), "w") as f:
        f.write(f"# {rnd}: `ncu --set full` of {title}\n\n| metric | unit | value |\n|---|---|---|\n")
        for k in keys:
            if k in hdr:
                f.write(f"| {k} | {units[hdr.index(k)]} | {vals[hdr.index(k)]} |\n")
        if alg_bytes:
            f.write(f"\nalgorithmic bytes per launch = {alg_bytes/1e6:.1f} MB\n")
    srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    tmp = os.path.join(src, rep_name.replace(".ncu-rep", "_src.csv")); open(tmp, "w").write(srcp)
    mix = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_opmix.py"), tmp], capture_output=True, text=True).stdout
    open(os.path.join(out, out_name), "a").write("\n## SASS opcode mix (warp-instructions executed) and stall samples\n\n```\n" + mix + "```\n")
    print(out_name, "written")


summarize("sac_full.ncu-rep", f"{rnd}_sac_head_kernel.md", "K3f, the fused SAC head kernel (B=65536, A=36, P=100, fwd+bwd)", 29240 * 65536)
summarize("rollout_full.ncu-rep", f"{rnd}_rollout_kernel.md", "K2f, the fused rollout kernel (B=65536, A=36, P=35)", 5336 * 65536)
for extra in ("configs_r01.json",):
    pth = os.path.join(src, extra)
    if os.path.exists(pth):
        open(os.path.join(out, f"{rnd}_configs.json"), "w").write(open(pth).read())
