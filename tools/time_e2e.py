"""Tuning aid: host-buffer end-to-end throughput of HostHeadPipeline for several chunk sizes / stream counts."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfpn_b200.host import HostHeadPipeline
from pfpn_b200 import synth
dev = torch.device("cuda:0")
B, A, P = 65536, 36, 35
pin = lambda t: t.contiguous().pin_memory()
g = torch.Generator().manual_seed(1)
logits = pin(torch.randn(B, A, P, generator=g) * 2); loc, logstd = synth.particle_grid(A, P, g); loc, logstd = pin(loc), pin(logstd)
value = pin(torch.rand(B, A, generator=g) * 2 - 1); adv = pin(torch.randn(B, generator=g)); lp_old = pin(torch.randn(B, generator=g) * 0.1 - 20)
for chunk, ns in [(8192, 3), (6144, 3), (12288, 3), (16384, 3), (8192, 4), (4096, 4), (16384, 2)]:
    pipe = HostHeadPipeline(B, A, P, dev, chunk=chunk, n_streams=ns)
    for _ in range(2): pipe.run(logits, loc, logstd, value, adv, lp_old)
    t0 = time.perf_counter(); n = 6
    for _ in range(n): pipe.run(logits, loc, logstd, value, adv, lp_old)
    ms = (time.perf_counter() - t0) * 1e3 / n
    print(json.dumps({"chunk": chunk, "streams": ns, "ms": round(ms, 3), "Msamples_s": round(B / ms / 1e3, 2),
                      "GBps_each_way": round(pipe.h2d_bytes / ms / 1e6, 1)}))
    del pipe
