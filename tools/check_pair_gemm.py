"""CTA-pair GEMM check: several shapes vs an fp64 reference, and timing at the DPPO shapes (PFPN_TC_PAIR=0/1 by env)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfpn_b200 import _cabi
from pfpn_b200.head import _stream_ptr
dev = torch.device("cuda:0"); st = _stream_ptr()
g = torch.Generator(device="cuda"); g.manual_seed(3)
def run(M, N, K, mn, lo, epi):
    A = torch.randn(M, K, device=dev, generator=g)
    Bt = torch.randn(N, K, device=dev, generator=g) * 0.05
    Bs = Bt.t().contiguous() if mn else Bt          # nn: B stored [K, N]
    b = torch.randn(N, device=dev, generator=g); H = torch.rand(M, N, device=dev, generator=g) * 8 - 1
    C = torch.full((M, N), float("nan"), device=dev)
    Blo = torch.empty_like(Bs); _cabi.check(_cabi.pfpn_split_lo(Bs.data_ptr(), Blo.data_ptr(), Bs.numel(), st))
    ldb = N if mn else K
    bias = b.data_ptr() if epi in (1, 2) else None
    hm = H.data_ptr() if epi == 3 else None
    if lo:
        f = _cabi.pfpn_tc_gemm_nn_lo if mn else _cabi.pfpn_tc_gemm_nt_lo
        call = lambda: _cabi.check(f(A.data_ptr(), K, Bs.data_ptr(), Blo.data_ptr(), ldb, C.data_ptr(), N, bias, hm, N, M, N, K, epi, st))
    else:
        f = _cabi.pfpn_tc_gemm_nn if mn else _cabi.pfpn_tc_gemm_nt
        call = lambda: _cabi.check(f(A.data_ptr(), K, Bs.data_ptr(), ldb, C.data_ptr(), N, bias, hm, N, M, N, K, epi, st))
    call(); torch.cuda.synchronize()
    ref = A.double() @ Bt.double().t()
    if epi in (1, 2): ref = ref + b.double()
    if epi == 2: ref = ref.clamp(0, 6)
    if epi == 3: ref = torch.where((H > 0) & (H < 6), ref, torch.zeros_like(ref))
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    return err, call
res = []
for (M, N, K) in [(256, 128, 32), (256, 128, 64), (384, 256, 96), (1000, 36 * 35, 512), (8192, 1024, 200), (4100, 512, 1024)]:
    for mn in (False, True):
        for lo in (True, False):
            for epi in (0, 2, 3):
                if N % 4 or K % 4: continue
                err, _ = run(M, N, K, mn, lo, epi)
                res.append((M, N, K, mn, lo, epi, err))
                if not (err < 5e-6): print("BAD", M, N, K, mn, lo, epi, err, flush=True)
print("max err", max(r[-1] for r in res), "cases", len(res), flush=True)
def t(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = {"pair": os.environ.get("PFPN_TC_PAIR", "1")}
for (M, N, K, mn, epi) in [(65536, 1024, 200, True, 2), (65536, 512, 1024, True, 2), (65536, 1260, 512, True, 1), (65536, 1024, 512, False, 3),
                           (65536, 512, 1260, False, 3), (8192, 1024, 200, True, 2), (8192, 512, 1024, True, 2), (8192, 1024, 512, False, 3)]:
    err, call = run(M, N, K, mn, True, epi)
    ms = t(call)
    out[f"{M}x{N}x{K}{'nn' if mn else 'nt'}"] = [round(ms, 4), round(2.0 * M * N * K / ms / 1e9, 1), float(f"{err:.2e}")]
print(json.dumps(out), flush=True)
