"""Measure the FP64 GEMM peak (cuBLAS DGEMM through torch.matmul) and a STREAM-style copy on this B200.

Same protocol as MEASURED_PEAKS.json's bf16 entry: 8192^3, best of 10 (burst) and back to back for 4 s
(sustained).  Writes gpurun_out/fp64_peak.json; the numbers are copied into profiles/ by hand.
"""
import json, time, os, torch
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
n = 8192
a = torch.rand(n, n, dtype=torch.float64, device=dev)
b = torch.rand(n, n, dtype=torch.float64, device=dev)
c = torch.empty(n, n, dtype=torch.float64, device=dev)
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / (best * 1e-3) / 1e12
t0 = time.time(); cnt = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(5):
        torch.matmul(a, b, out=c); cnt += 1
    torch.cuda.synchronize()
e1.record(); e1.synchronize()
sus = 2 * n**3 * cnt / (e0.elapsed_time(e1) * 1e-3) / 1e12
# skinny DGEMM resembling the first contraction (M=90000*8, K=300, N=50)
A = torch.rand(720000, 300, dtype=torch.float64, device=dev)
B = torch.rand(300, 50, dtype=torch.float64, device=dev)
C = torch.empty(720000, 50, dtype=torch.float64, device=dev)
for _ in range(3): torch.matmul(A, B, out=C)
torch.cuda.synchronize()
bs = 1e9
for _ in range(10):
    e0.record(); torch.matmul(A, B, out=C); e1.record(); e1.synchronize()
    bs = min(bs, e0.elapsed_time(e1))
skinny = 2 * 720000 * 300 * 50 / (bs * 1e-3) / 1e12
# copy bandwidth
x = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev); y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
bc = 1e9
for _ in range(10):
    e0.record(); y.copy_(x); e1.record(); e1.synchronize()
    bc = min(bc, e0.elapsed_time(e1))
gbs = 2 * x.numel() * 2 / (bc * 1e-3) / 1e9
res = {"fp64_tflops_burst": burst, "fp64_tflops_sustained": sus, "fp64_skinny_720000x300x50_tflops": skinny,
       "hbm_copy_gbs": gbs, "gpu": torch.cuda.get_device_name(0), "how": "torch.matmul fp64 8192^3 best of 10 / 4 s loop"}
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/fp64_peak.json", "w"), indent=1)
