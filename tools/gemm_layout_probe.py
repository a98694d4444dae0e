"""Times gatk_gemm_batched (NN, products shape) with A interleaved [N, H*F] vs head-major [H, N, F], and C likewise."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygat_b200 import _lib
N, H, F, D = 2449029, 8, 100, 64
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
W = torch.randn(F, H * D, device=dev) * 0.1
def run(name, A, lda, a_bs, C, ldc, c_bs, M=N, Nn=D, K=F, tb=0, B=W, ldb=H * D, b_bs=D):
    wsb = _lib.query("gatk_gemm_batched_workspace_bytes", 0, tb, M, Nn, K, H)
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
    def call():
        _lib.call("gatk_gemm_batched", 0, tb, M, Nn, K, H, A.data_ptr(), lda, a_bs, B.data_ptr(), ldb, b_bs, C.data_ptr(), ldc, c_bs, 0,
                  ws.data_ptr(), wsb, st)
    for _ in range(3): call()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = (M * K * H + M * Nn * H) * 4 / 1e9
    print(f"{name:40s} {ms:7.3f} ms  {gb / ms:7.1f} GB/s", flush=True)
A = torch.randn(N, H * F, device=dev)
C = torch.empty(N, H * D, device=dev)
run("project A[N,H*F] C[N,H*D]", A, H * F, F, C, H * D, D)
run("project A[H,N,F] C[N,H*D]", A, F, N * F, C, H * D, D)
run("project A[H,N,F] C[H,N,D]", A, F, N * F, C, D, N * D)
run("project A[N,H*F] C[H,N,D]", A, H * F, F, C, D, N * D)
# dxagg: A = dhp [N, H*D], B = W^T (nt), C = dxagg [N, H*F]
G = torch.randn(N, H * D, device=dev)
X = torch.empty(N, H * F, device=dev)
run("dxagg A[N,H*D] C[N,H*F]", G, H * D, D, X, H * F, F, Nn=F, K=D, tb=1)
run("dxagg A[N,H*D] C[H,N,F]", G, H * D, D, X, F, N * F, Nn=F, K=D, tb=1)
run("dxagg A[H,N,D] C[H,N,F]", G, D, N * D, X, F, N * F, Nn=F, K=D, tb=1)
X2 = torch.empty(N, H * 104, device=dev)
run("dxagg A[N,H*D] C[N,H*104]", G, H * D, D, X2, H * 104, 104, Nn=F, K=D, tb=1)
A2 = torch.randn(N, H * 104, device=dev)
run("project A[N,H*104] C[N,H*D]", A2, H * 104, 104, C, H * D, D)
