"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/supplynet.h
declares; struct layouts of the ctypes binding match the header as compiled by gcc; host-only
entry points (validation) behave; nothing here launches a kernel."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

HDR = os.path.join(ROOT, "include", "supplynet.h")


@pytest.fixture(scope="module")
def lib():
    from multi_echelon_drl_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        subprocess.run(["make", "-C", os.path.dirname(_lib.LIB_PATH)], check=True)
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    from multi_echelon_drl_b200 import _lib
    text = open(HDR).read()
    declared = set(re.findall(r"\b(sn_[a-z0-9_]+)\s*\(", text))
    assert declared >= {"sn_abi_version", "sn_last_error", "sn_obs_dim", "sn_spec_validate", "sn_reset", "sn_step",
                        "sn_observe", "sn_heuristic", "sn_rollout"}
    for name in declared:
        assert hasattr(lib, name), f"libsupplynet.so does not export {name}"
    assert set(_lib.SIGNATURES) == declared, "ctypes binding and header disagree on the entry points"
    assert lib.sn_abi_version() == 11


def test_struct_layout_matches_header(tmp_path):
    from multi_echelon_drl_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "supplynet.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(sn_spec),offsetof(sn_spec,agents),offsetof(sn_spec,holding_cost),offsetof(sn_spec,dpa_ub),'
                   'sizeof(sn_policy),offsetof(sn_policy,rule),offsetof(sn_policy,lb),offsetof(sn_spec,n_items),'
                   'offsetof(sn_spec,item_backorder_cost));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    S, P = _lib.SnSpec, _lib.SnPolicy
    assert got == [C.sizeof(S), S.agents.offset, S.holding_cost.offset, S.dpa_ub.offset, C.sizeof(P), P.rule.offset,
                   P.lb.offset, S.n_items.offset, S.item_backorder_cost.offset]


def test_step_io_layout_matches_header(tmp_path):
    """struct sn_step_io (the argument block of sn_step_ex) as gcc lays it out vs the ctypes mirror."""
    from multi_echelon_drl_b200 import _lib
    fields = ["state", "period", "ctl", "env_reward", "stats", "vn_scratch", "vn_gamma", "vn_nobs", "vn_clip_obs", "vn_epsilon"]
    src = tmp_path / "layout_io.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "supplynet.h"\nint main(){printf("%zu", sizeof(sn_step_io));'
                   + "".join(f'printf(" %zu", offsetof(sn_step_io, {f}));' for f in fields)
                   + 'printf(" %d %d %d\\n", SN_CTL_WORDS, SN_CTL_SEQ, SN_VN_ONEPASS_SCRATCH_DOUBLES);return 0;}\n')
    exe = tmp_path / "layout_io"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    S = _lib.SnStepIo
    assert got == [C.sizeof(S)] + [getattr(S, f).offset for f in fields] + [_lib.SN_CTL_WORDS, _lib.SN_CTL_SEQ,
                                                                           _lib.SN_VN_ONEPASS_SCRATCH_DOUBLES]


def test_spec_validation_and_error_mapping(lib):
    from multi_echelon_drl_b200 import _lib, uniform_spec
    spec = uniform_spec(["retailer", "wholesaler", "distributor", "manufacturer"], True)
    c = spec.to_c()
    for dt in (_lib.SN_I32, _lib.SN_F32, _lib.SN_F64):
        assert lib.sn_spec_validate(C.byref(c), dt) == 0
    assert lib.sn_obs_dim(C.byref(c)) == 28
    c2 = uniform_spec(["wholesaler"], False).to_c()
    assert lib.sn_obs_dim(C.byref(c2)) == 7
    # non-contiguous agent set: the reference crashes at supplynetwork.py:199; we reject up front
    c2.n_agents, c2.agents[0], c2.agents[1] = 2, 0, 2
    assert lib.sn_spec_validate(C.byref(c2), _lib.SN_I32) == _lib.SN_ERR_INVALID
    with pytest.raises(ValueError, match="contiguous"):
        _lib.check(lib.sn_spec_validate(C.byref(c2), _lib.SN_I32))
    # lead times: 1..4 with info + shipment <= 5 (generic kernels, 57 state words); else unsupported
    assert lib.sn_state_words(C.byref(c)) == 40 and lib.sn_init_words(C.byref(c)) == 19
    c3 = spec.to_c()
    c3.info_leadtime[0] = 3
    assert lib.sn_spec_validate(C.byref(c3), _lib.SN_I32) == 0
    assert lib.sn_state_words(C.byref(c3)) == 57 and lib.sn_init_words(C.byref(c3)) == 36
    c3.info_leadtime[0] = 4
    with pytest.raises(NotImplementedError):
        _lib.check(lib.sn_spec_validate(C.byref(c3), _lib.SN_I32))
    c3.info_leadtime[0], c3.ship_leadtime[0] = 0, 2
    with pytest.raises(NotImplementedError):
        _lib.check(lib.sn_spec_validate(C.byref(c3), _lib.SN_I32))
    # fractional constants with the integer simulator -> TypeError
    c4 = spec.to_c()
    c4.coplayer_target[1] = 19.5
    with pytest.raises(TypeError):
        _lib.check(lib.sn_spec_validate(C.byref(c4), _lib.SN_I32))
    assert lib.sn_spec_validate(C.byref(c4), _lib.SN_F32) == 0
    assert lib.sn_spec_validate(None, _lib.SN_I32) == _lib.SN_ERR_INVALID
    assert lib.sn_spec_validate(C.byref(c), 7) == _lib.SN_ERR_INVALID


def test_argument_checks_precede_any_launch(lib):
    """Bad arguments are rejected on the host (no GPU needed to see the error)."""
    from multi_echelon_drl_b200 import _lib, uniform_spec
    c = uniform_spec(None, True).to_c()
    rc = lib.sn_step(C.byref(c), _lib.SN_I32, 8, 8, 100, None, None, None, None, None, None, None, None, None, None)
    assert rc == _lib.SN_ERR_TERMINAL and b"terminal" in lib.sn_last_error()      # envs.py:88-89
    assert lib.sn_step(C.byref(c), _lib.SN_I32, 8, 6, 0, None, None, None, None, None, None, None, None, None, None) == _lib.SN_ERR_INVALID
    assert lib.sn_rollout(C.byref(c), _lib.SN_I32, 8, 8, 90, 20, None, None, 0, None, None, None, None, None, None, None,
                          None, None) == _lib.SN_ERR_INVALID
    assert lib.sn_reset(C.byref(c), _lib.SN_I32, 0, 0, None, None, None, None, None) == _lib.SN_ERR_INVALID


def test_product_fails_loudly_without_cuda():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multi_echelon_drl_b200 import uniform_spec
    from multi_echelon_drl_b200.simulator import BasePolicyParams, BatchedSupplyNetwork, heuristic_actions
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchedSupplyNetwork(uniform_spec(), 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        heuristic_actions(BasePolicyParams([19, 20, 20, 14]), torch.zeros(2, 28))


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (checked on the source text)."""
    pkg = os.path.join(ROOT, "multi_echelon_drl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "beergame_c" not in txt and "beergame_np" not in txt, f


def test_every_entry_point_reports_through_sn_last_error(lib):
    """Argument errors of the VecNormalize and table-fill entry points carry a message too (one error channel for the
    whole library, csrc/supplynet_error.h)."""
    from multi_echelon_drl_b200 import _lib
    rc = lib.sn_vecnorm_apply(10, 99, None, None, None, None, None, None, 1.0, 1.0, 1e-8, 0, None, None)
    assert rc == _lib.SN_ERR_INVALID and b"obs_dim=99" in lib.sn_last_error()
    rc = lib.sn_vecnorm_update(0, 28, None, None, 0.99, None, None, None, None, None, None, None)
    assert rc == _lib.SN_ERR_INVALID and b"sn_vecnorm_update" in lib.sn_last_error()
    rc = lib.sn_fill_uniform(_lib.SN_I32, None, 10, 0, 5, 1, 0, None)
    assert rc == _lib.SN_ERR_INVALID and b"sn_fill_uniform" in lib.sn_last_error()
    with pytest.raises(ValueError, match="sn_fill_uniform"):
        _lib.check(rc)


def test_specialised_tree_library_serves_its_own_network_only():
    """specialize.build -> csrc/spec/libsupplynet_tree_<key>.so (nvcc, about 7 s): same ABI version and symbols for the
    period step / rollouts; its sn_spec_validate accepts exactly the network it was compiled for (host-only calls)."""
    import shutil
    if shutil.which("nvcc") is None:
        pytest.skip("needs nvcc")
    from multi_echelon_drl_b200 import _lib, specialize, uniform_spec
    from multi_echelon_drl_b200.spec import tree_spec
    star = tree_spec([dict(name="r1", demand=(0, 8), target=19), dict(name="r2", demand=(0, 6), target=12), dict(name="r3", demand=(0, 3), target=9),
                      dict(name="hub", target=40)],
                     [("ext", "hub", 2, 3), ("hub", "r3", 1, 1), ("hub", "r1", 2, 2), ("hub", "r2", 2, 1)], ["r1", "r2", "r3", "hub"], True)
    t = specialize.tables(star)
    assert t["nf"] == 4 and t["kf"] == 4 and t["sup"][:4] == [3, 3, 3, -1] and t["rank"][:4] == [1, 2, 0, 0]
    assert t["max_rank"] == 2 and t["lat"][3] in (0, 1, 2) and t["li"][:4] == [2, 2, 1, 2] and t["ls"][:4] == [2, 1, 1, 3]
    path = specialize.build(star, "int32")
    assert os.path.isfile(path) and os.path.basename(path) == f"libsupplynet_tree_{specialize.key(star, 'int32')}.so"
    lib = C.CDLL(path)
    for name in ("sn_abi_version", "sn_spec_validate", "sn_step_ex", "sn_rollout", "sn_episode", "sn_last_error"):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = _lib.SIGNATURES[name]
    assert lib.sn_abi_version() == _lib.ABI_VERSION
    assert lib.sn_spec_validate(C.byref(star.to_c()), _lib.SN_I32) == 0
    # other agents / observation mode: same tables, still served; another topology, other lead times, a chain: refused
    assert lib.sn_spec_validate(C.byref(tree_spec([dict(name="r1", demand=(0, 8), target=19), dict(name="r2", demand=(0, 6), target=12),
                                                     dict(name="r3", demand=(0, 3), target=9), dict(name="hub", target=40)],
                                                    [("ext", "hub", 2, 3), ("hub", "r3", 1, 1), ("hub", "r1", 2, 2), ("hub", "r2", 2, 1)],
                                                    ["hub"], False).to_c()), _lib.SN_I32) == 0
    other = tree_spec([dict(name="r1", demand=(0, 8), target=19), dict(name="r2", demand=(0, 6), target=12), dict(name="r3", demand=(0, 3), target=9),
                       dict(name="hub", target=40)],
                      [("ext", "hub", 2, 3), ("hub", "r1", 2, 2), ("hub", "r3", 1, 1), ("hub", "r2", 2, 1)], ["r1", "r2", "r3", "hub"], True)
    assert lib.sn_spec_validate(C.byref(other.to_c()), _lib.SN_I32) == _lib.SN_ERR_UNSUPPORTED      # other service order
    chain = uniform_spec(["retailer", "wholesaler", "distributor", "manufacturer"], True)
    assert lib.sn_spec_validate(C.byref(chain.to_c()), _lib.SN_I32) == _lib.SN_ERR_UNSUPPORTED
    assert lib.sn_spec_validate(C.byref(star.to_c()), _lib.SN_F32) == _lib.SN_OK     # validation is dtype-agnostic ...
    # ... the compute entry points are not: this library was built for int32 only
    rc = lib.sn_rollout(C.byref(star.to_c()), _lib.SN_F32, 8, 8, 0, 1, C.c_void_p(8), C.c_void_p(8), _lib.SN_POLICY_BASE_STOCK, None,
                        C.byref(_lib.SnPolicy()), None, None, None, None, None, None, None)
    assert rc == _lib.SN_ERR_UNSUPPORTED and b"dtype" in lib.sn_last_error()
    assert specialize.key(star, "int32") != specialize.key(other, "int32") != specialize.key(star, "float32")


def test_sn_specialize_auto_builds_the_missing_library_on_first_load(monkeypatch):
    """SN_SPECIALIZE=auto: specialize.load compiles the library of a tree network that has none yet (nvcc on the path)
    and serves it; without the switch a missing library means the general kernels (None)."""
    import shutil
    if shutil.which("nvcc") is None:
        pytest.skip("needs nvcc")
    from multi_echelon_drl_b200 import specialize
    from multi_echelon_drl_b200.spec import tree_spec
    tree = tree_spec([dict(name="a", demand=(0, 7), target=11), dict(name="b", target=21), dict(name="c", demand=(0, 3), target=9)],
                     [("ext", "b", 1, 1), ("b", "c", 1, 2), ("b", "a", 2, 1)], ["a", "b", "c"], True)
    path = specialize.lib_path(tree, "int32")
    if os.path.isfile(path):
        os.remove(path)
    specialize._loaded.pop(path, None)
    monkeypatch.delenv("SN_SPECIALIZE", raising=False)
    assert specialize.load(tree, "int32") is None
    specialize._loaded.pop(path, None)
    monkeypatch.setenv("SN_SPECIALIZE", "auto")
    try:
        lib = specialize.load(tree, "int32")
        assert lib is not None and os.path.isfile(path) and lib.sn_abi_version() == specialize._lib.ABI_VERSION
    finally:
        specialize._loaded.pop(path, None)
        if os.path.isfile(path):
            os.remove(path)          # not one of the libraries build() ships


def test_sn_step_ex_and_pack_rows_reject_inconsistent_arguments_before_any_launch(lib):
    """Host-side validation of the round-2 entry points (no GPU needed: every call fails before its first CUDA call):
    the combinations of VecNormalize extras that make no sense, multi-item rewards on single-item batches, the terminal
    period, unknown table types - each with its own message through sn_last_error."""
    from multi_echelon_drl_b200 import _lib, uniform_spec
    from multi_echelon_drl_b200.spec import tree_spec
    spec = uniform_spec(["retailer", "wholesaler", "distributor", "manufacturer"], True).to_c()
    fake = 1 << 20                                           # an aligned, never dereferenced address

    def call(dtype=_lib.SN_F32, n=64, ld=64, spec=spec, **kw):
        io = _lib.SnStepIo()
        io.state, io.actions, io.demand, io.period = fake, fake, fake, 0
        for k, v in kw.items():
            setattr(io, k, v)
        rc = lib.sn_step_ex(C.byref(spec), dtype, n, ld, C.byref(io), None)
        return rc, lib.sn_last_error().decode()

    assert lib.sn_step_ex(C.byref(spec), _lib.SN_F32, 64, 64, None, None) == _lib.SN_ERR_INVALID
    rc, msg = call(period=100)
    assert rc == _lib.SN_ERR_TERMINAL and "terminal" in msg                       # envs.py:88-89
    rc, msg = call(period=-1)
    assert rc == _lib.SN_ERR_INVALID and "period=-1" in msg
    rc, msg = call(actions=fake + 4)
    assert rc == _lib.SN_ERR_INVALID and "16-byte" in msg
    rc, msg = call(env_reward=fake)
    assert rc == _lib.SN_ERR_UNSUPPORTED and "n_items=1" in msg
    rc, msg = call(vn_scratch=fake)
    assert rc == _lib.SN_ERR_INVALID and "vn_obs_rms" in msg
    rc, msg = call(vn_obs_rms=fake)
    assert rc == _lib.SN_ERR_INVALID and "vn_scratch (update), vn_nobs (normalise)" in msg
    rc, msg = call(vn_obs_rms=fake, vn_scratch=fake, vn_returns=fake)
    assert rc == _lib.SN_ERR_INVALID and "vn_ret_rms" in msg
    rc, msg = call(vn_obs_rms=fake, vn_nobs=fake, vn_ret_rms=fake, vn_returns=fake)
    assert rc == _lib.SN_ERR_INVALID and "needs vn_scratch" in msg
    rc, msg = call(vn_obs_rms=fake, vn_scratch=fake, vn_nrew=fake)
    assert rc == _lib.SN_ERR_INVALID and "vn_nrew needs vn_nobs" in msg
    rc, msg = call(vn_obs_rms=fake, vn_scratch=fake)
    assert rc == _lib.SN_ERR_INVALID and "need ctl and obs" in msg
    rc, msg = call(dtype=_lib.SN_F64, vn_obs_rms=fake, vn_nobs=fake, ctl=fake, obs=fake)
    assert rc == _lib.SN_ERR_UNSUPPORTED and "SN_I32 / SN_F32" in msg
    tree = tree_spec([dict(name="a", demand=(0, 7), target=11), dict(name="b", target=21)], [("ext", "b", 1, 1), ("b", "a", 2, 1)],
                     ["a", "b"], True).to_c()
    rc, msg = call(spec=tree, vn_obs_rms=fake, vn_scratch=fake, ctl=fake, obs=fake)
    assert rc == _lib.SN_ERR_UNSUPPORTED and "single-item beer-game chain" in msg
    rc, msg = call(n=0)
    assert rc == _lib.SN_ERR_INVALID
    rc, msg = call(n=64, ld=32)
    assert rc == _lib.SN_ERR_INVALID

    def pack(src_dtype=_lib.SN_SRC_I64, src=fake, n=8, width=19, stride=19, dtype=_lib.SN_I32, dst=fake, rows=19, ld=32, drs=32):
        rc = lib.sn_pack_rows(src_dtype, src, n, width, stride, dtype, dst, rows, ld, drs, None)
        return rc, lib.sn_last_error().decode()
    assert pack(src=None)[0] == _lib.SN_ERR_INVALID and "NULL" in lib.sn_last_error().decode()
    for kw in (dict(n=0), dict(rows=18), dict(ld=4), dict(stride=18), dict(drs=16)):
        rc, msg = pack(**kw)
        assert rc == _lib.SN_ERR_INVALID and "sn_pack_rows: n_envs=" in msg, kw
    rc = lib.sn_counter_add(None, 5, None)
    assert rc == _lib.SN_ERR_INVALID


def test_stats_exchange_argument_validation(lib):
    """sn_mail_bytes / sn_stats_allreduce (the peer-memory all-reduce of the statistics vector) reject bad arguments
    before anything is launched (host-only: no GPU here)."""
    from multi_echelon_drl_b200 import _lib
    assert lib.sn_mail_bytes(1) == 2 * 1 * 16 * 16 and lib.sn_mail_bytes(8) == 2 * 8 * 16 * 16
    assert lib.sn_mail_bytes(0) == _lib.SN_ERR_INVALID and lib.sn_mail_bytes(17) == _lib.SN_ERR_INVALID
    fake = C.c_void_p(0x1000)

    def call(stats=fake, out=fake, mails=fake, rank=0, world=2, seq=1):
        return lib.sn_stats_allreduce(stats, out, mails, rank, world, seq, None), lib.sn_last_error().decode()
    assert call(stats=None)[0] == _lib.SN_ERR_INVALID
    assert call(mails=None)[0] == _lib.SN_ERR_INVALID
    for kw in (dict(world=0), dict(world=17), dict(rank=2), dict(rank=-1)):
        rc, msg = call(**kw)
        assert rc == _lib.SN_ERR_INVALID and "rank" in msg, kw
    rc, msg = call(seq=0)
    assert rc == _lib.SN_ERR_INVALID and "seq counts from 1" in msg
    assert lib.sn_mail_alloc(2, None, None) == _lib.SN_ERR_INVALID
    assert lib.sn_mail_open(None, None) == _lib.SN_ERR_INVALID
