import ctypes as C, os, sys, itertools
import torch
here = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(here, "libprobe.so"))
lib.probe_run.restype = C.c_int
lib.probe_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
torch.manual_seed(0)
X = torch.randn(512, 64, device="cuda").bfloat16()
W = torch.randn(128, 64, device="cuda").bfloat16()
out = torch.zeros(128, 64, device="cuda")
Xf, Wf = X.float(), W.float()
print("K-major A (fprop halo):  row0 sbo base_off -> max rel err")
for sbo in (1024, 1280, 2048, 2304):
    for row0 in (0, 1, 2, 3, 9, 19):
        for bo in sorted({0, row0 % 8}):
            out.zero_()
            rc = lib.probe_run(X.data_ptr(), W.data_ptr(), row0, sbo, bo, 0, out.data_ptr())
            m = torch.arange(128, device="cuda")
            rows = row0 + (m // 8) * (sbo // 128) + m % 8
            ref = Xf[rows] @ Wf[:64].t()
            err = ((out - ref).abs().max() / ref.abs().max()).item()
            print(f"  sbo={sbo:5d} row0={row0:2d} bo={bo}  rc={rc} err={err:.3e} {'OK' if err < 1e-2 else ''}")
print("MN-major A (wgrad halo): rows are K; D[m][n] = sum_k X[krow(k)][m] * W[n][k], m < 64 valid")
for sbo in (1024, 1280, 2048, 2304):
    for row0 in (0, 1, 2, 3, 9):
        for bo in sorted({0, row0 % 8}):
            out.zero_()
            rc = lib.probe_run(X.data_ptr(), W.data_ptr(), row0, sbo, bo, 1, out.data_ptr())
            k = torch.arange(64, device="cuda")
            krows = row0 + (k // 8) * (sbo // 128) + k % 8
            A = Xf[krows]                      # [K=64][M=64]
            ref = A.t() @ Wf[:64].t()           # [M=64][N=64]
            err = ((out[:64] - ref).abs().max() / ref.abs().max()).item()
            print(f"  sbo={sbo:5d} row0={row0:2d} bo={bo}  rc={rc} err={err:.3e} {'OK' if err < 1e-2 else ''}")
