import torch, time
dev = torch.device('cuda', 0)
for mb in (1, 4, 64):
    n = mb * 2**20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print(mb, "MB", name, "%.1f us  %.1f GB/s" % (dt * 1e6, n / dt / 1e9))
    def both():
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    both(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): both()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(mb, "MB duplex", "%.1f us  %.1f GB/s per direction" % (dt * 1e6, n / dt / 1e9))
