"""Library comparison bars (NOT on the product path): cuBLAS DGEMM and cuSOLVER potrf via torch, fp64."""
import json, time, torch
torch.backends.cuda.preferred_linalg_library("cusolver")
dev = torch.device("cuda:0")
def ev_time(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = ev_time(lambda: torch.matmul(a, b))
    out[f"dgemm_{n}_tflops"] = 2 * n**3 / ms * 1e-9
    spd = a @ a.T + n * torch.eye(n, dtype=torch.float64, device=dev)
    ms = ev_time(lambda: torch.linalg.cholesky(spd))
    out[f"potrf_{n}_ms"] = ms
    out[f"potrf_{n}_tflops"] = n**3 / 3 / ms * 1e-9
    del a, b, spd
# batched potrf 8 x 4096
n = 4096
a = torch.randn(8, n, n, dtype=torch.float64, device=dev)
spd = a @ a.transpose(1, 2) + n * torch.eye(n, dtype=torch.float64, device=dev)
ms = ev_time(lambda: torch.linalg.cholesky(spd), reps=3, warm=1)
out["potrf_batched8_4096_ms"] = ms
print(json.dumps(out))
