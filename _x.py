import math, time, torch
import radiorust_b200 as rr
ctx = rr.Context(0)
def lowpass(c): return lambda b, f: 1.0 if abs(f) <= c else 0.0
def run(name, S, sr, n, c, flt):
    ch = rr.Chain(ctx, [rr.Filter.new(lowpass(100000.0))], flt, n_streams=S)
    L = n * c
    dt = torch.complex64 if flt == "f32" else torch.complex128
    x = torch.randn(S, L, dtype=dt, device="cuda"); y = torch.empty(S, L, dtype=dt, device="cuda")
    for _ in range(3): ch.push_device(sr, n, c, x.data_ptr(), L, y.data_ptr(), L, L)
    ch.sync(); t0 = time.perf_counter()
    for _ in range(5): ch.push_device(sr, n, c, x.data_ptr(), L, y.data_ptr(), L, L)
    ch.sync(); d = (time.perf_counter() - t0) / 5
    print(f"{name}: {d*1e3:.3f} ms/push {S*L/d/1e9:.1f} GS/s plan={ch.plan}", flush=True)
    ch.close()
run("C4 filter f32 2^17", 256, 10e6, 65536, 4, "f32")
run("C5 filter f64 2^20", 16, 2.4e6, 1 << 19, 4, "f64")
run("f32 2^15", 1024, 10e6, 16384, 4, "f32")
run("f64 2^14", 512, 10e6, 8192, 4, "f64")
