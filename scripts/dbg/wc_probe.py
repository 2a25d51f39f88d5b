"""H2D bandwidth from default pinned vs write-combined pinned host memory, alone and with a concurrent D2H stream."""
import ctypes, torch
from cuda import cudart
n = 64 << 20
dev = torch.device("cuda")
def wc_tensor(nbytes):
    err, ptr = cudart.cudaHostAlloc(nbytes, cudart.cudaHostAllocWriteCombined)
    assert int(err) == 0, err
    buf = (ctypes.c_char * nbytes).from_address(int(ptr))
    return torch.frombuffer(buf, dtype=torch.uint8)
srcs = {"pinned": torch.empty(n, dtype=torch.uint8).pin_memory(), "write-combined": wc_tensor(n)}
for k, t in srcs.items():
    t.fill_(1)
    print(k, "is_pinned", t.is_pinned())
dst = torch.empty(n, dtype=torch.uint8, device=dev)
back_src = torch.empty(n, dtype=torch.uint8, device=dev)
back_dst = torch.empty(n, dtype=torch.uint8).pin_memory()
s2 = torch.cuda.Stream()
for k, t in srcs.items():
    for duplex in (False, True):
        for _ in range(2):
            dst.copy_(t, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dst.copy_(t, non_blocking=True)
            if duplex:
                with torch.cuda.stream(s2):
                    back_dst.copy_(back_src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        print(f"{k:15s} duplex={duplex}: H2D {10 * n / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
