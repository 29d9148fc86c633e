"""Debug aid: which element of an MN-major B operand does the tensor core read for (n, k)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfpn_b200 import _cabi
from pfpn_b200.head import _stream_ptr
dev = torch.device("cuda:0")
M, N, K = 128, 128, 32
W = (torch.arange(K).float()[:, None] * 1000 + torch.arange(N).float()[None, :]).to(dev).contiguous()
bias = torch.zeros(N, device=dev)
for k1 in (0, 1, 7, 8, 31):
    X = torch.zeros(M, K, device=dev); X[:, k1] = 1.0
    C = torch.full((M, N), float("nan"), device=dev)
    _cabi.check(_cabi.pfpn_tc_gemm_nn(X.data_ptr(), K, W.data_ptr(), N, C.data_ptr(), N, bias.data_ptr(), None, 0, M, N, K, 1, _stream_ptr()))
    torch.cuda.synchronize()
    print("k1", k1, "row0[:12]", C[0, :12].tolist(), "row0[30:36]", C[0, 30:36].tolist(), "row5[:4]", C[5, :4].tolist())
# wgrad: dW[k, n] = sum_m X[m,k] dY[m,n]; X one-hot in m
import ctypes as CT
Mb, Kw, Nw = 64, 128, 128
for m1 in (0, 1, 9, 33):
    X = torch.zeros(Mb, Kw, device=dev); X[m1, :] = torch.arange(Kw, device=dev).float() + 1
    dY = torch.zeros(Mb, Nw, device=dev); dY[m1, :] = (torch.arange(Nw, device=dev).float() + 1) * 0.001
    dW = torch.full((Kw, Nw), float("nan"), device=dev)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
    _cabi.check(_cabi.pfpn_tc_linear_bwd_weight(X.data_ptr(), Kw, dY.data_ptr(), Nw, dW.data_ptr(), None, Mb, Kw, Nw, ws.data_ptr(), ws.numel(), _stream_ptr()))
    torch.cuda.synchronize()
    ref = X.T @ dY
    print("m1", m1, "err", float((dW - ref).abs().max()), "dW[0,:4]", dW[0, :4].tolist(), "dW[3,:4]", dW[3, :4].tolist(), "ref[3,:4]", ref[3, :4].tolist())
