"""HBM bandwidth by read/write mix on this B200 (context for the rollout kernel, which streams
16 % reads / 84 % writes): torch fill (write-only), copy (50/50) and sum (read-only) over 1 GiB."""
import json
import torch
n = 1 << 28
a = torch.empty(n, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
def t(fn, bytes_):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); best = min(best, s.elapsed_time(e))
    return bytes_ / best / 1e6
print(json.dumps({"fill_write_only_GBs": t(lambda: a.fill_(1.0), 4 * n), "copy_GBs": t(lambda: b.copy_(a), 8 * n),
                  "sum_read_only_GBs": t(lambda: a.sum(), 4 * n)}))
