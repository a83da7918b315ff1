"""Read-bandwidth calibration on the current GPU: torch reductions / copies over buffers >> L2."""
import torch, time
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
n = 1 << 30
x = torch.randn(n, device="cuda", dtype=torch.bfloat16)
y = torch.empty_like(x)
xf = x.view(torch.float32)
xi = x.view(torch.int32)
print("copy r+w  GB/s", 2 * x.numel() * 2 / t(lambda: y.copy_(x)) / 1e9)
print("sum bf16  GB/s", x.numel() * 2 / t(lambda: x.sum()) / 1e9)
print("sum f32   GB/s", x.numel() * 2 / t(lambda: xf.sum()) / 1e9)
print("max int32 GB/s", x.numel() * 2 / t(lambda: xi.max()) / 1e9)
print("fill (w)  GB/s", x.numel() * 2 / t(lambda: y.zero_()) / 1e9)
