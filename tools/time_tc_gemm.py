import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfpn_b200 import _cabi
from pfpn_b200.head import _stream_ptr
dev = torch.device("cuda:0")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
M = 65536
for (N, K) in [(1024, 200), (512, 1024), (1260, 512), (1024, 512), (512, 1260)]:
    A = torch.randn(M, K, device=dev); Bt = torch.randn(N, K, device=dev) * 0.05; W = Bt.t().contiguous()
    b = torch.randn(N, device=dev); C = torch.empty(M, N, device=dev); C2 = torch.empty(M, N, device=dev)
    st = _stream_ptr()
    Blo = torch.empty_like(Bt); _cabi.check(_cabi.pfpn_split_lo(Bt.data_ptr(), Blo.data_ptr(), Bt.numel(), st))
    # the production form: weight low halves pre-split once per optimizer step, persistent kernel
    tc = lambda: _cabi.check(_cabi.pfpn_tc_gemm_nt_lo(A.data_ptr(), K, Bt.data_ptr(), Blo.data_ptr(), K, C.data_ptr(), N, b.data_ptr(), None, 0, M, N, K, 2, st))
    ff = lambda: _cabi.check(_cabi.pfpn_mlp_linear_fwd(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), C2.data_ptr(), N, M, K, N, 1, st))
    ms_tc, ms_ff = t(tc), t(ff)
    ref = (A.double() @ Bt.double().t() + b.double()).clamp(0, 6)
    e_tc = float((C.double() - ref).abs().max() / ref.abs().max()); e_ff = float((C2.double() - ref).abs().max() / ref.abs().max())
    fl = 2.0 * M * N * K
    print(json.dumps({"M": M, "N": N, "K": K, "tc_ms": round(ms_tc, 3), "tc_TFLOPs": round(fl / ms_tc / 1e9, 1), "ffma_ms": round(ms_ff, 3),
                      "ffma_TFLOPs": round(fl / ms_ff / 1e9, 1), "err_tc": e_tc, "err_ffma": e_ff}))
# input-gradient (Relu6Grad mask epilogue) and weight-gradient shapes of the DPPO trunk
import ctypes as CT
for (N, K) in [(1024, 512), (512, 1260)]:
    dY = torch.randn(M, K, device=dev) * 0.01; W = torch.randn(N, K, device=dev) * 0.05; H = torch.rand(M, N, device=dev) * 8 - 1
    dX = torch.empty(M, N, device=dev); st = _stream_ptr()
    f = lambda: _cabi.check(_cabi.pfpn_tc_gemm_nt(dY.data_ptr(), K, W.data_ptr(), K, dX.data_ptr(), N, None, H.data_ptr(), N, M, N, K, 3, st))
    ms = t(f)
    print(json.dumps({"kind": "dX(mask)", "M": M, "N": N, "K": K, "tc_ms": round(ms, 3), "tc_TFLOPs": round(2.0 * M * N * K / ms / 1e9, 1)}))
for (Kin, N) in [(200, 1024), (1024, 512), (512, 1260)]:
    X = torch.randn(M, Kin, device=dev); dY = torch.randn(M, N, device=dev) * 0.01; dW = torch.empty(Kin, N, device=dev); db = torch.empty(N, device=dev)
    n = CT.c_size_t(0); _cabi.check(_cabi.pfpn_tc_wgrad_workspace_bytes(M, Kin, N, CT.byref(n)))
    ws = torch.empty(n.value, dtype=torch.uint8, device=dev); st = _stream_ptr()
    f = lambda: _cabi.check(_cabi.pfpn_tc_linear_bwd_weight(X.data_ptr(), Kin, dY.data_ptr(), N, dW.data_ptr(), db.data_ptr(), M, Kin, N, ws.data_ptr(), ws.numel(), st))
    ms = t(f)
    print(json.dumps({"kind": "wgrad", "M": M, "Kin": Kin, "N": N, "tc_ms": round(ms, 3), "tc_TFLOPs": round(2.0 * M * N * Kin / ms / 1e9, 1)}))
