"""PCIe probe: pinned D2H / H2D bandwidth alone and concurrently, at the sizes of the config-2 host path
(trajectory 760 MB device->host, inputs 137 MB host->device).  Prints GB/s; used to state the floor of `e2e`."""
import torch

dev = torch.device("cuda:0")
d2h_bytes, h2d_bytes = 760 << 20, 137 << 20
d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
h_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a.record()
        fn()
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def d2h():
    s1.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        h_out.copy_(d_out, non_blocking=True)


def h2d():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        d_in.copy_(h_in, non_blocking=True)


def both():
    d2h(); h2d()


def chunked(k=20):
    s1.wait_stream(torch.cuda.current_stream())
    n = d2h_bytes // k
    with torch.cuda.stream(s1):
        for c in range(k):
            h_out[c * n:(c + 1) * n].copy_(d_out[c * n:(c + 1) * n], non_blocking=True)


t = timed(d2h); print(f"D2H alone   {d2h_bytes / t / 1e6:7.1f} GB/s  {t:.3f} ms")
t = timed(h2d); print(f"H2D alone   {h2d_bytes / t / 1e6:7.1f} GB/s  {t:.3f} ms")
t = timed(both); print(f"both        {(d2h_bytes + h2d_bytes) / t / 1e6:7.1f} GB/s total  {t:.3f} ms")
t = timed(chunked); print(f"D2H 20 chunks {d2h_bytes / t / 1e6:7.1f} GB/s  {t:.3f} ms")
