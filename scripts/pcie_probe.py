"""Dev: pinned D2H / H2D bandwidth ceiling of the box and the e2e pipeline at several chunk sizes."""
import os, sys, time, json, subprocess
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda", 0)
n = 1 << 27  # 1 GiB of doubles
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device=dev)
for name, src, dst in (("D2H", d, h), ("H2D", h, d)):
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(f"{name} pinned 1 GiB: {n*8/dt/1e9:.1f} GB/s", flush=True)
# two concurrent D2H streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h[: n // 2].copy_(d[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2): h[n // 2:].copy_(d[n // 2:], non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print(f"D2H two streams: {n*8/dt/1e9:.1f} GB/s", flush=True)
del h, d
for chunk in (2048, 8192, 32768):
    out = subprocess.run([sys.executable, "bench.py", "--steps", "5", "--warmup", "3", "--no-cpu", "--no-nuts", "--host-chunk", str(chunk)],
                         capture_output=True, text=True).stdout.strip().splitlines()[-1]
    e = json.loads(out)["e2e"]
    print(f"host_chunk={chunk}: e2e {e['value']:.0f} traj/s, {e['ms_per_step']:.1f} ms/step, {e['d2h_bytes_per_step']/e['ms_per_step']/1e6:.1f} GB/s D2H", flush=True)
