"""Times the UNMODIFIED reference's Python env stepping loop in the build container (context for
DESIGN.md / BASELINE notes; the reference cannot travel to the GPU box).  Workload = BASELINE.md 4:
complex scenario, centralized, global observation, random init, fixed action [4,4,4,4], 100-step
episodes, per-env np.random.seed, P processes x 16 envs stepped round-robin.
Usage: python oracle/time_reference.py [P]"""
import multiprocessing as mp
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(args):
    wid, n_envs, episodes = args
    import numpy as np
    from oracle import ref_harness as H
    warnings.simplefilter("ignore")
    envs = [H.make_env("uniform", H.NODE_NAMES, True, True, 100) for _ in range(n_envs)]
    a = np.array([4, 4, 4, 4])
    t0 = time.perf_counter()
    steps = 0
    for ep in range(episodes):
        for i, e in enumerate(envs):
            np.random.seed(1000 * wid + 16 * ep + i)
            e.reset()
        for _ in range(100):
            for e in envs:
                e.step(a)
                steps += 1
    return steps, time.perf_counter() - t0


if __name__ == "__main__":
    P = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    for procs in sorted({1, P}):
        with mp.Pool(procs) as pool:
            res = pool.map(worker, [(w, 16, 3) for w in range(procs)])
        total = sum(r[0] for r in res)
        slowest = max(r[1] for r in res)
        print(f"reference python: {procs} process(es): {total / slowest:,.0f} env-steps/s aggregate "
              f"({total / slowest / procs:,.0f} per core), {total} steps, slowest worker {slowest:.1f}s")
