"""Synthetic probe of the trajectory rollout's memory pattern (see write_pattern.cu)."""
import ctypes as C, json, os, torch
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libwrite_pattern.so"))
lib.probe_write.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
lib.probe_write_in.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
T, n = 100, 65536
out = torch.empty(T * n * 28, dtype=torch.float32, device="cuda")
act = torch.randint(0, 17, (T, n, 4), dtype=torch.int32, device="cuda")
dem = torch.randint(0, 9, (T + 3, n), dtype=torch.int32, device="cuda")
def timeit(run):
    for _ in range(3): run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): run()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / 10
st = lambda: torch.cuda.current_stream().cuda_stream
for delay in (0, 100, 200, 300):
    row = {"dependent_fmas_per_period": delay}
    row["compute_only_ms"] = round(timeit(lambda: lib.probe_write(out.data_ptr(), n, T, 2, delay, 0, st())), 4)
    row["stores_ms"] = round(timeit(lambda: lib.probe_write(out.data_ptr(), n, T, 0, delay, 0, st())), 4)
    row["stores+inputs_pd1_ms"] = round(timeit(lambda: lib.probe_write_in(out.data_ptr(), act.data_ptr(), dem.data_ptr(), n, T, 3, delay, st())), 4)
    row["stores+inputs_pd4_ms"] = round(timeit(lambda: lib.probe_write_in(out.data_ptr(), act.data_ptr(), dem.data_ptr(), n, T, 4, delay, st())), 4)
    print(json.dumps(row))
