import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dinov2_b200 import engine as E
E.LIB_PATH = os.path.abspath(sys.argv[1])
M = 64 * 1370
for name, N, K in (("proj", 1024, 1024), ("fc2", 1024, 4096)):
    A = (torch.randn(M, K, device="cuda") * 0.5).half(); W = (torch.randn(N, K, device="cuda") * 0.05).half()
    bias = torch.randn(N, device="cuda") * 0.1; ls = torch.rand(N, device="cuda") + 0.3
    out = torch.zeros(M, N, device="cuda"); gam = torch.randn(N, device="cuda"); bet = torch.randn(N, device="cuda")
    ln = torch.empty(M, N, device="cuda", dtype=torch.half)
    nblk = (M + 127) // 128
    cnt = torch.zeros(2 * nblk + 2 + 8, device="cuda", dtype=torch.int32)
    for _ in range(3):
        E.kernel_gemm_resid_ln(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), ls.data_ptr(), out.data_ptr(), gam.data_ptr(), bet.data_ptr(), 1e-6, ln.data_ptr(), cnt.data_ptr())
    torch.cuda.synchronize()
    t = cnt[2 * nblk + 2: 2 * nblk + 7].tolist()
    print(f"{name}: CTA0 worker 8: ticket {t[0]} wait {t[1]} rows {t[2]} done {t[3]} cycles over {t[4]} slices -> per slice: ticket {t[0]/max(t[4],1):.0f} wait {t[1]/max(t[4],1):.0f} rows {t[2]/max(t[4],1):.0f} done {t[3]/max(t[4],1):.0f}")
