"""How much of a freshly written tensor does the next kernel still find in L2?  write X MB (copy), then read it (sum): the read's
effective bandwidth vs X.  (tuning aid for the sub-batching question)"""
import torch

dev = "cuda"
for mb in (5, 10, 21, 42, 63, 84, 126, 168, 336):
    n = mb * (1 << 20) // 4
    src = torch.randn(n, device=dev)
    dst = torch.empty_like(src)
    big = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    res = []
    for mode in ("warm", "cold"):
        ts = []
        for _ in range(8):
            dst.copy_(src)
            if mode == "cold":
                big.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            s = dst.sum()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        res.append(ts[len(ts) // 2])
    print(f"{mb:4d} MB: read after write {res[0]*1e3:7.1f} us ({mb/1024/res[0]*1e3:6.2f} TB/s)   after an L2 flush {res[1]*1e3:7.1f} us ({mb/1024/res[1]*1e3:6.2f} TB/s)")
