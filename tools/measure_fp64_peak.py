"""Measure the FP64 roofline denominator on the box: cuBLAS DGEMM 8192^3 via torch.matmul
(burst = best of 10, sustained = back-to-back for ~4 s), same method MEASURED_PEAKS.json
uses for bf16.  Also runs tools/fp64_micro (DFMA vs DMMA issue rates).  Writes
gpurun_out/fp64_peaks.json; the committed copy lives in profiles/FP64_PEAKS.json."""
import json, os, subprocess, sys, time
import torch

def main():
    out = {}
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["dgemm_tflops_burst"] = 2 * n**3 / (best * 1e-3) / 1e12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); k = 0
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(5):
            torch.matmul(a, b, out=c); k += 1
        torch.cuda.synchronize()
    e1.record(); e1.synchronize()
    out["dgemm_tflops_sustained"] = 2 * n**3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
    out["gpu_name"] = torch.cuda.get_device_name(0)
    out["how"] = "torch.matmul fp64 8192^3 (cuBLAS DGEMM): best of 10 (burst), back-to-back 4 s (sustained)"
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fp64_micro")
    if os.path.exists(exe):
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        try:
            out["micro"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:  # noqa
            out["micro_error"] = (r.stdout + r.stderr)[-500:]
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/fp64_peaks.json", "w"), indent=1)
    print(json.dumps(out))

if __name__ == "__main__":
    main()
