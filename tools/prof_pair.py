import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfpn_b200 import _cabi
from pfpn_b200.head import _stream_ptr
dev = torch.device("cuda:0"); st = _stream_ptr()
M, N, K = 65536, 512, 1024
A = torch.randn(M, K, device=dev); Bs = torch.randn(K, N, device=dev) * 0.05; b = torch.randn(N, device=dev); C = torch.empty(M, N, device=dev)
Blo = torch.empty_like(Bs); _cabi.check(_cabi.pfpn_split_lo(Bs.data_ptr(), Blo.data_ptr(), Bs.numel(), st))
for _ in range(4):
    _cabi.check(_cabi.pfpn_tc_gemm_nn_lo(A.data_ptr(), K, Bs.data_ptr(), Blo.data_ptr(), N, C.data_ptr(), N, b.data_ptr(), None, N, M, N, K, 2, st))
torch.cuda.synchronize(); print("ok")
