import torch, time
n = 121*10000
d = torch.randn(n, dtype=torch.float64, device='cuda')
h = torch.empty(n, dtype=torch.float64, pin_memory=True)
s2 = [torch.cuda.Stream(), torch.cuda.Stream()]
def one():
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
def two():
    half = n//2
    ev = torch.cuda.Event(); ev.record()
    for i, s in enumerate(s2):
        s.wait_event(ev)
        with torch.cuda.stream(s):
            h[i*half:(i+1)*half].copy_(d[i*half:(i+1)*half], non_blocking=True)
    torch.cuda.synchronize()
for f in (one, two, one, two):
    for _ in range(5): f()
    t = time.perf_counter()
    for _ in range(50): f()
    dt = (time.perf_counter()-t)/50
    print(f.__name__, '%.3f ms  %.1f GB/s' % (dt*1e3, n*8/dt/1e9))
# small copies latency
d1 = torch.randn(50000, dtype=torch.float64, device='cuda'); h1 = torch.empty(50000, dtype=torch.float64, pin_memory=True)
for _ in range(5): h1.copy_(d1, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(100): h1.copy_(d1, non_blocking=True); torch.cuda.synchronize()
print('0.4MB d2h+sync %.3f ms' % ((time.perf_counter()-t)/100*1e3))
hh = torch.empty((13,10000), dtype=torch.float64, pin_memory=True)
t = time.perf_counter()
for _ in range(100): dd = hh.to('cuda', non_blocking=True); torch.cuda.synchronize()
print('1MB h2d+sync %.3f ms' % ((time.perf_counter()-t)/100*1e3))
t = time.perf_counter()
for _ in range(100): x = torch.empty((11,11,10000), dtype=torch.float64, pin_memory=True)
print('pinned empty %.3f ms' % ((time.perf_counter()-t)/100*1e3))
