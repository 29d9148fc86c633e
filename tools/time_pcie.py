"""Raw PCIe ceiling of the box: pinned H2D alone, D2H alone, both concurrently (330 MB each), vs the e2e pipeline."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda:0"); n = 65536 * 36 * 35
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, device=dev); d_out = torch.randn(n, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=6):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return round(n * 4 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
for _ in range(2): run(True, True)
print(json.dumps({"h2d_alone_GBps": run(True, False), "d2h_alone_GBps": run(False, True), "both_GBps_each_way": run(True, True)}))
