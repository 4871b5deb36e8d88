import time, torch
x = [torch.randn(101, 100, 25, device="cuda") for _ in range(1000)]   # 1000 blocks like the README run (1.01 GB)
big = [torch.randn(50, 65536, 100, device="cuda") for _ in range(4)]   # 4 blocks of 1.3 GB
for name, blocks in (("1000 small blocks", x), ("4 large blocks", big)):
    shape = (sum(len(b) for b in blocks), *blocks[0].shape[1:])
    for pin in (True, False, True, False):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        host = torch.empty(shape, dtype=torch.float32, pin_memory=pin)
        t1 = time.perf_counter()
        off = 0
        for b in blocks:
            host[off:off + len(b)].copy_(b, non_blocking=True); off += len(b)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"{name}: pin={pin}: alloc {t1-t0:.3f} s, copy {t2-t1:.3f} s, total {t2-t0:.3f} s, {host.numel()*4/1e9:.2f} GB")
        del host
    t0 = time.perf_counter(); dev = torch.cat(blocks, 0); h = dev.cpu(); torch.cuda.synchronize()
    print(f"{name}: cat on device + .cpu(): {time.perf_counter()-t0:.3f} s")
    del dev, h
