"""Measures FP64 roofline denominators on the GPU box (MEASURED_PEAKS.json has none):
cuBLAS DGEMM through torch.matmul (burst best-of-10 and a 3 s sustained loop) and the
DFMA / DMMA issue-bound microbenchmark tools/fp64_peak. Writes gpurun_out/fp64_peaks.json."""
import json, os, subprocess, time
import torch

out = {}
exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fp64_peak")
try:
    out["micro"] = json.loads(subprocess.check_output([exe]).decode())
except Exception as e:  # noqa
    out["micro_error"] = str(e)
dev = torch.device("cuda:0")
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["dgemm_%d_tflops_burst" % n] = 2.0 * n ** 3 / best * 1e-9
    t0 = time.time(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); reps = 0
    while time.time() - t0 < 3.0:
        for _ in range(5):
            torch.matmul(a, b); reps += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    out["dgemm_%d_tflops_sustained" % n] = 2.0 * n ** 3 * reps / e0.elapsed_time(e1) * 1e-9
# skinny shapes like ours: [440 x 512] x [512 x B]
for (m, k, nn) in ((440, 512, 16384), (440, 440, 16384), (250, 353, 16384), (1515, 1515, 8192)):
    a = torch.randn(m, k, dtype=torch.float64, device=dev)
    b = torch.randn(k, nn, dtype=torch.float64, device=dev)
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["dgemm_%dx%dx%d_tflops" % (m, k, nn)] = 2.0 * m * k * nn / best * 1e-9
out["clocks"] = subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"]).decode().strip()
out["nproc"] = os.cpu_count()
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/fp64_peaks.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
