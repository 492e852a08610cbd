import torch, time, numpy as np
torch.cuda.set_device(0)
dev = torch.empty(2_560_000, dtype=torch.uint8, device="cuda")
big = torch.empty(300 << 20, dtype=torch.uint8).pin_memory()
small = torch.empty(2_560_000, dtype=torch.uint8).pin_memory()
s = torch.cuda.Stream()
def d2h(dst, reps=5, pre=None):
    ts = []
    for _ in range(reps):
        if pre: pre(dst)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            a.record(); dst.copy_(dev, non_blocking=True); b.record()
        b.synchronize(); ts.append(a.elapsed_time(b))
    return ["%.3f" % t for t in ts]
def h2d(src, reps=5, pre=None):
    ts = []
    for _ in range(reps):
        if pre: pre(src)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s):
            a.record(); dev.copy_(src, non_blocking=True); b.record()
        b.synchronize(); ts.append(a.elapsed_time(b))
    return ["%.3f" % t for t in ts]
sl = big[64 << 20:(64 << 20) + 2_560_000]
n = sl.numpy()
print("d2h small untouched", d2h(small))
print("d2h slice untouched", d2h(sl))
print("d2h slice after cpu read ", d2h(sl, pre=lambda t: t.numpy().sum()))
print("d2h slice after cpu write", d2h(sl, pre=lambda t: t.numpy().fill(3)))
print("d2h small after cpu read ", d2h(small, pre=lambda t: t.numpy().sum()))
print("d2h small after cpu write", d2h(small, pre=lambda t: t.numpy().fill(3)))
print("h2d small untouched", h2d(small))
print("h2d small after cpu write", h2d(small, pre=lambda t: t.numpy().fill(5)))
print("h2d slice after cpu write", h2d(sl, pre=lambda t: t.numpy().fill(5)))
out = np.empty(2_560_000, dtype=np.uint8)
print("d2h slice after cpu copy-out (memcpy read)", d2h(sl, pre=lambda t: np.copyto(out, t.numpy())))
