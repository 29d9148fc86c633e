"""tcgen05 3xTF32 GEMM (K6 tensor-core path) against fp64: must stay within the same 1e-5 norm-wise
tolerance as the FFMA anchor, otherwise it may not be used (north_star)."""
import pytest
import torch

from pfpn_b200 import _cabi
from pfpn_b200.head import _stream_ptr

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (256, 128, 64), (300, 1024, 200), (1000, 512, 1024), (517, 1260, 512),
                                   (4096, 1024, 200), (77, 132, 36)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_tc_gemm_nt_matches_fp64(cuda_dev, M, N, K, epi):
    g = torch.Generator().manual_seed(M * 7 + N + K + epi)
    A = torch.randn(M, K, generator=g)
    Bt = torch.randn(N, K, generator=g) * 0.05
    bias = torch.randn(N, generator=g) * 0.1
    H = torch.randn(M, N, generator=g) * 4
    cu = lambda t: t.to(cuda_dev).contiguous()
    Ad, Bd, bd, Hd = cu(A), cu(Bt), cu(bias), cu(H)
    C = torch.full((M, N), float("nan"), device=cuda_dev)
    _cabi.check(_cabi.pfpn_tc_gemm_nt(Ad.data_ptr(), K, Bd.data_ptr(), K, C.data_ptr(), N, bd.data_ptr(), Hd.data_ptr(), N,
                                      M, N, K, epi, _stream_ptr()))
    ref = A.double() @ Bt.double().T
    if epi in (1, 2):
        ref = ref + bias.double()
    if epi == 2:
        ref = ref.clamp(0, 6)
    if epi == 3:
        ref = ref * ((H > 0) & (H < 6)).double()
    assert torch.isfinite(C).all()
    assert rel(C, ref) < 1e-5


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 1024, 200), (1000, 512, 1024), (517, 1260, 512), (77, 132, 36)])
@pytest.mark.parametrize("epi", [1, 2])
def test_tc_gemm_nn_weights_as_stored(cuda_dev, M, N, K, epi):
    """Forward layer act(X W + b) with W [in, out] read as an MN-major operand (no transposed copy)."""
    g = torch.Generator().manual_seed(M * 5 + N + K + epi)
    X = torch.randn(M, K, generator=g)
    W = torch.randn(K, N, generator=g) * 0.05
    bias = torch.randn(N, generator=g) * 0.1
    cu = lambda t: t.to(cuda_dev).contiguous()
    Xd, Wd, bd = cu(X), cu(W), cu(bias)
    C = torch.full((M, N), float("nan"), device=cuda_dev)
    _cabi.check(_cabi.pfpn_tc_gemm_nn(Xd.data_ptr(), K, Wd.data_ptr(), N, C.data_ptr(), N, bd.data_ptr(), None, 0,
                                      M, N, K, epi, _stream_ptr()))
    ref = X.double() @ W.double() + bias.double()
    if epi == 2:
        ref = ref.clamp(0, 6)
    assert torch.isfinite(C).all()
    assert rel(C, ref) < 1e-5


@pytest.mark.parametrize("M,K,N", [(2048, 200, 1024), (5000, 1024, 512), (4099, 512, 1260), (513, 36, 132), (65536, 200, 128)])
def test_tc_wgrad_operands_as_stored(cuda_dev, M, K, N):
    """dW = X^T dY from row-major X [M, K] and dY [M, N] (both MN-major operands), split-K over the batch."""
    g = torch.Generator().manual_seed(M + K + N)
    X = torch.randn(M, K, generator=g)
    dY = torch.randn(M, N, generator=g) * 0.01
    cu = lambda t: t.to(cuda_dev).contiguous()
    Xd, Yd = cu(X), cu(dY)
    import ctypes as C
    n = C.c_size_t(0)
    _cabi.check(_cabi.pfpn_tc_wgrad_workspace_bytes(M, K, N, C.byref(n)))
    ws = torch.empty(max(n.value, 16), dtype=torch.uint8, device=cuda_dev)
    dW = torch.full((K, N), float("nan"), device=cuda_dev)
    db = torch.full((N,), float("nan"), device=cuda_dev)
    _cabi.check(_cabi.pfpn_tc_linear_bwd_weight(Xd.data_ptr(), K, Yd.data_ptr(), N, dW.data_ptr(), db.data_ptr(), M, K, N,
                                                ws.data_ptr(), ws.numel(), _stream_ptr()))
    ref = X.double().T @ dY.double()
    assert torch.isfinite(dW).all()
    assert rel(dW, ref) < 1e-5
    # bias gradient accumulated from the staged dY tiles (fc_layer's `b`, networks/ops.py:108-116)
    assert rel(db, dY.double().sum(0)) < 1e-5
    # deterministic: a second launch reproduces both bit for bit
    dW2, db2 = torch.empty_like(dW), torch.empty_like(db)
    _cabi.check(_cabi.pfpn_tc_linear_bwd_weight(Xd.data_ptr(), K, Yd.data_ptr(), N, dW2.data_ptr(), db2.data_ptr(), M, K, N,
                                                ws.data_ptr(), ws.numel(), _stream_ptr()))
    assert torch.equal(dW, dW2) and torch.equal(db, db2)


@pytest.mark.parametrize("M,N,K", [(777, 512, 1024), (300, 1260, 512)])
def test_presplit_weight_operand_is_bit_identical(cuda_dev, M, N, K):
    """pfpn_tc_gemm_{nn,nt}_lo (B_lo = B - tf32(B) precomputed by pfpn_split_lo, fetched by TMA) vs the on-the-fly splitter."""
    from pfpn_b200 import _cabi
    from pfpn_b200.head import _stream_ptr
    g = torch.Generator().manual_seed(M)
    X = torch.randn(M, K, generator=g).to(cuda_dev)
    W = (torch.randn(K, N, generator=g) * 0.05).to(cuda_dev)
    b = torch.randn(N, generator=g).to(cuda_dev)
    dY = torch.randn(M, N, generator=g).to(cuda_dev)
    Wlo = torch.empty_like(W)
    st = _stream_ptr()
    _cabi.check(_cabi.pfpn_split_lo(W.data_ptr(), Wlo.data_ptr(), W.numel(), st))
    hi = (W.view(torch.int32) & -8192).view(torch.float32)
    assert torch.equal(Wlo, W - hi)
    Y0, Y1 = torch.empty(M, N, device=cuda_dev), torch.empty(M, N, device=cuda_dev)
    _cabi.check(_cabi.pfpn_tc_gemm_nn(X.data_ptr(), K, W.data_ptr(), N, Y0.data_ptr(), N, b.data_ptr(), None, 0, M, N, K, 2, st))
    _cabi.check(_cabi.pfpn_tc_gemm_nn_lo(X.data_ptr(), K, W.data_ptr(), Wlo.data_ptr(), N, Y1.data_ptr(), N, b.data_ptr(), None, 0,
                                         M, N, K, 2, st))
    assert torch.equal(Y0, Y1)
    D0, D1 = torch.empty(M, K, device=cuda_dev), torch.empty(M, K, device=cuda_dev)
    _cabi.check(_cabi.pfpn_tc_gemm_nt(dY.data_ptr(), N, W.data_ptr(), N, D0.data_ptr(), K, None, X.data_ptr(), K, M, K, N, 3, st))
    _cabi.check(_cabi.pfpn_tc_gemm_nt_lo(dY.data_ptr(), N, W.data_ptr(), Wlo.data_ptr(), N, D1.data_ptr(), K, None, X.data_ptr(), K,
                                         M, K, N, 3, st))
    assert torch.equal(D0, D1)


_FORMS = r"""
import hashlib, sys, torch
from pfpn_b200 import _cabi
from pfpn_b200.head import _stream_ptr
dev = torch.device("cuda:0"); st = _stream_ptr(); h = hashlib.sha256()
for (M, N, K) in [(256, 128, 32), (1000, 1260, 512), (4100, 512, 1024), (8192, 1024, 200), (129, 132, 36)]:
    g = torch.Generator().manual_seed(M + N)
    X = torch.randn(M, K, generator=g).to(dev); W = (torch.randn(K, N, generator=g) * 0.05).to(dev)
    b = torch.randn(N, generator=g).to(dev); dY = torch.randn(M, N, generator=g).to(dev)
    Wlo = torch.empty_like(W); _cabi.check(_cabi.pfpn_split_lo(W.data_ptr(), Wlo.data_ptr(), W.numel(), st))
    Y = torch.empty(M, N, device=dev); D = torch.empty(M, K, device=dev); D2 = torch.empty(M, K, device=dev)
    _cabi.check(_cabi.pfpn_tc_gemm_nn_lo(X.data_ptr(), K, W.data_ptr(), Wlo.data_ptr(), N, Y.data_ptr(), N, b.data_ptr(), None, 0, M, N, K, 2, st))
    _cabi.check(_cabi.pfpn_tc_gemm_nt_lo(dY.data_ptr(), N, W.data_ptr(), Wlo.data_ptr(), N, D.data_ptr(), K, None, X.data_ptr(), K, M, K, N, 3, st))
    _cabi.check(_cabi.pfpn_tc_gemm_nt(dY.data_ptr(), N, W.data_ptr(), N, D2.data_ptr(), K, None, X.data_ptr(), K, M, K, N, 0, st))
    torch.cuda.synchronize()
    for t in (Y, D, D2):
        h.update(t.cpu().numpy().tobytes())
print("HASH", h.hexdigest())
"""


def test_cta_pair_form_is_bit_identical_to_the_one_cta_forms(cuda_dev):
    """The cta_group::2 kernel (default for M > 128) against the persistent one-CTA kernel and the one-tile kernel: the same
    products reach the same accumulators in the same order, so every output bit must agree (ragged M, N, K included).
    The form is chosen once per process (PFPN_TC_PAIR / PFPN_TC_PERSISTENT), hence three child processes."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hashes = []
    for env in ({"PFPN_TC_PAIR": "1"}, {"PFPN_TC_PAIR": "0"}, {"PFPN_TC_PAIR": "0", "PFPN_TC_PERSISTENT": "0"}):
        e = dict(os.environ, **env)
        e["PYTHONPATH"] = root + os.pathsep + e.get("PYTHONPATH", "")
        r = subprocess.run([sys.executable, "-c", _FORMS], capture_output=True, text=True, timeout=300, env=e, cwd=root)
        assert r.returncode == 0, r.stderr[-2000:]
        hashes.append([l for l in r.stdout.splitlines() if l.startswith("HASH")][-1])
    assert hashes[0] == hashes[1] == hashes[2], hashes
