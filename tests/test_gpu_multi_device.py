"""Shard invariance across REAL devices ("N envs on 1 GPU == same N on 8 GPUs"): one global batch, cut by
env index over all visible GPUs, gives bit for bit what the whole batch gives on one GPU.  Needs at least two GPUs
(skipped on the single-GPU box of the round-end test run; `bench.py --gpus N` makes the same check under torchrun at
every N > 1 and reports it as `shard_invariance`)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_index_sharding_over_devices_equals_one_device():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two or more GPUs")
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.distributed import shard_range
    from multi_echelon_drl_b200.simulator import BatchedSupplyNetwork
    from multi_echelon_drl_b200.utils.demands import bulk_episode_inputs
    G, n, T = torch.cuda.device_count(), 10007, 100
    spec = uniform_spec(("retailer", "wholesaler", "distributor", "manufacturer"), True, T, True)
    rs = np.random.RandomState(5)
    init, demand = bulk_episode_inputs(spec, n, rs)
    acts = rs.randint(0, 17, (T, n, 4)).astype(np.int32)

    def run(dev, lo, hi):
        with torch.cuda.device(dev):
            sim = BatchedSupplyNetwork(spec, hi - lo, f"cuda:{dev}", "int32")
            sim.load_episode(init[lo:hi], demand[lo:hi])
            out = sim.episode(T, actions=torch.as_tensor(acts[:, lo:hi]), record_obs=True, record_reward=True)
            return out["traj_obs"].cpu(), out["traj_reward"].cpu(), sim.ep_return.cpu()

    whole = run(0, 0, n)
    parts = [run(g, *shard_range(n, g, G)) for g in range(G)]
    assert torch.equal(torch.cat([p[0] for p in parts], dim=1), whole[0])
    assert torch.equal(torch.cat([p[1] for p in parts], dim=1), whole[1])
    assert torch.equal(torch.cat([p[2] for p in parts]), whole[2])


def _mailboxes(lib, devices, world):
    """One mailbox per device of this process, every device allowed to write to every other one."""
    import ctypes as C
    from multi_echelon_drl_b200 import _lib
    mails = []
    for d in devices:
        with torch.cuda.device(d):
            m, h = C.c_void_p(), (C.c_ubyte * 64)()
            _lib.check(lib.sn_mail_alloc(world, C.byref(m), h))
            mails.append(m)
    for a in devices:
        for b in devices:
            _lib.check(lib.sn_enable_peer_access(a, b))
    return mails


def test_stats_allreduce_single_rank_is_the_identity_over_many_calls():
    """sn_stats_allreduce with one rank: the vector goes through the rank's own mailbox (both slot parities, growing
    sequence tags) and comes back unchanged, in place and out of place."""
    import ctypes as C
    from multi_echelon_drl_b200 import _lib
    lib = _lib.load()
    mails = _mailboxes(lib, [0], 1)
    ptrs = torch.tensor([mails[0].value], dtype=torch.int64, device="cuda:0")
    g = torch.Generator(device="cuda:0"); g.manual_seed(0)
    for seq in range(1, 8):
        x = torch.randn(16, dtype=torch.float64, device="cuda:0", generator=g) * 1e6
        y = torch.empty_like(x)
        _lib.check(lib.sn_stats_allreduce(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(ptrs.data_ptr()), 0, 1,
                                          seq, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        assert torch.equal(x, y)
        keep = x.clone()
        _lib.check(lib.sn_stats_allreduce(C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(ptrs.data_ptr()), 0, 1,
                                          seq + 100, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        assert torch.equal(x, keep)
    torch.cuda.synchronize()
    _lib.check(lib.sn_mail_free(mails[0]))


def test_stats_allreduce_over_peer_memory_equals_the_sum_in_rank_order():
    """All visible GPUs as the ranks of one process: every rank's kernel stores into every mailbox over NVLink and sums
    what arrives; all ranks get the same bits, equal to the float64 sum taken in rank order, call after call."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two or more GPUs")
    import ctypes as C
    from multi_echelon_drl_b200 import _lib
    lib = _lib.load()
    G = min(torch.cuda.device_count(), 8)
    mails = _mailboxes(lib, list(range(G)), G)
    ptrs = [torch.tensor([m.value for m in mails], dtype=torch.int64, device=f"cuda:{d}") for d in range(G)]
    rs = np.random.RandomState(1)
    for seq in range(1, 12):
        host = rs.randn(G, 16) * 10.0 ** rs.randint(-3, 9, (G, 16))
        xs = [torch.as_tensor(host[d]).to(f"cuda:{d}") for d in range(G)]
        ys = [torch.empty_like(x) for x in xs]
        for d in range(G):                                   # every rank launches before anybody synchronises
            with torch.cuda.device(d):
                _lib.check(lib.sn_stats_allreduce(C.c_void_p(xs[d].data_ptr()), C.c_void_p(ys[d].data_ptr()),
                                                  C.c_void_p(ptrs[d].data_ptr()), d, G, seq,
                                                  C.c_void_p(torch.cuda.current_stream(d).cuda_stream)))
        want = np.zeros(16)
        for d in range(G):
            want = want + host[d]
        for d in range(G):
            torch.cuda.synchronize(d)
            assert np.array_equal(ys[d].cpu().numpy(), want), (seq, d)
    for d in range(G):
        with torch.cuda.device(d):
            _lib.check(lib.sn_mail_free(mails[d]))


def test_peer_stats_allreduce_class_on_a_one_rank_group(tmp_path):
    """distributed.PeerStatsAllReduce end to end on the one GPU of the test box: mailbox allocation, the handle exchange
    through the process group (one rank), calls in and out of place with a growing sequence number on the statistics
    vector of a real batch, argument checks, close()."""
    import torch.distributed as dist
    from multi_echelon_drl_b200 import register_envs as R
    from multi_echelon_drl_b200.distributed import PeerStatsAllReduce, summarize
    from multi_echelon_drl_b200.simulator import BasePolicyParams
    dist.init_process_group("gloo", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        ar = PeerStatsAllReduce(torch.device("cuda", 0))
        env = R.make("BeerGameUniformMultiFacilityFullInfo-v0", n_envs=512, dtype="int32", seed=2)
        sim = env.sim
        env._next_episode()
        sim.track_stats = True
        sim.stats.zero_()
        sim.episode(policy=BasePolicyParams([19, 20, 20, 14], 0, 16, "a"), write_state=False)
        total = ar(sim.stats, torch.empty_like(sim.stats))
        assert torch.equal(total, sim.stats) and summarize(total)["env_steps"] == 512 * 100
        keep = sim.stats.clone()
        for _ in range(5):
            ar(sim.stats)                                   # in place
        assert torch.equal(sim.stats, keep) and ar.seq == 6
        with pytest.raises(ValueError):
            ar(torch.zeros(3, dtype=torch.float64, device="cuda:0"))
        with pytest.raises(ValueError):
            ar(torch.zeros(16, dtype=torch.float32, device="cuda:0"))
        ar.close()
        ar.close()                                          # idempotent
    finally:
        dist.destroy_process_group()
