import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from textflux_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
M, N, K = 2560, 9216, 3072
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
W = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
b = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
def run(cg, iters=10):
    def f():
        _lib.check(lib.tfx_op_linear(A.data_ptr(), K, W.data_ptr(), b.data_ptr(), out.data_ptr(), N, M, N, K, 0, None, None, cg, st))
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): f()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / iters
    return 2.0 * M * N * K / ms / 1e9
print("cg", sys.argv[1], "TF/s", run(int(sys.argv[1])))
