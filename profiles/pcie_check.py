"""Floor of the e2e leg: pinned H2D / D2H bandwidth of this box for the byte counts of one configs[1] tick."""
import torch, time
dev = torch.device("cuda", 0)
h2d_bytes, d2h_bytes = 481_282_964, 137_504_920
hin = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
hout = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
din = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
dout = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
def both():
    h2d(); d2h()
def chunked():
    k = 16 * 9
    step = h2d_bytes // k
    with torch.cuda.stream(s1):
        for i in range(k): din[i * step:(i + 1) * step].copy_(hin[i * step:(i + 1) * step], non_blocking=True)
a, b, c, d = t(h2d), t(d2h), t(both), t(chunked)
print(f"H2D {h2d_bytes/1e6:.0f} MB: {a:.2f} ms ({h2d_bytes/a/1e6:.1f} GB/s); D2H {d2h_bytes/1e6:.0f} MB: {b:.2f} ms ({d2h_bytes/b/1e6:.1f} GB/s); "
      f"both directions at once: {c:.2f} ms; H2D in 144 pieces: {d:.2f} ms")
