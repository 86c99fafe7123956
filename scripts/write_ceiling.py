import torch, time
dev = torch.device("cuda:0")
x = torch.empty(16, 64, 320, 320, device=dev)
y = torch.empty_like(x)
big = torch.empty(1 << 28, device=dev)   # 1 GiB
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
nb = x.numel() * 4
for name, fn, bytes_ in [("zero_ 419MB", lambda: x.zero_(), nb), ("fill_ 419MB", lambda: x.fill_(1.5), nb),
                         ("copy_ 419MB (r+w)", lambda: y.copy_(x), 2 * nb), ("zero_ 1GiB", lambda: big.zero_(), big.numel() * 4),
                         ("fill_ 1GiB", lambda: big.fill_(2.0), big.numel() * 4)]:
    ms = t(fn)
    print(f"{name:22s} {ms*1e3:8.1f} us  {bytes_/ms/1e6:8.1f} GB/s")
