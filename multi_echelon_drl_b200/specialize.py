"""Ahead-of-time specialisation of the distribution-tree kernels for ONE network.

The tree kernels of libsupplynet.so take the topology at run time (supplier table, customer ranks, latest-demand
sources, lead times as 0/1 multipliers from the constant bank): general, but 2.5x the instructions of the tuned
beer-game chain.  A deployment that simulates one network millions of times can bake those tables in:

    python -m multi_echelon_drl_b200.specialize network_configs/some_tree.yaml [--dtype int32]

compiles `csrc/supplynet_kernels.cu` once more with -DSN_STATIC_TOPO -DSN_ST_*=... (nvcc, sm_100a) into
`csrc/spec/libsupplynet_tree_<key>.so`, a library with the same C ABI that serves exactly that network and refuses
any other.  `BatchedSupplyNetwork` loads it, when present, for `sn_step` / `sn_rollout` / `sn_episode` (reset and
observe stay on the general library: same state layout).  Trees of five to eight nodes are compiled with eight
facilities per thread (-DSN_KF=8): a register-resident env where the general library spreads it over eight lanes.
Results are bit-identical to the general kernels (tests/test_gpu_trees.py); nothing is compiled at run time unless
`build()` is called or the environment says SN_SPECIALIZE=auto (then the first env built for a tree network compiles
its library if nvcc is on the path); SN_NO_SPECIALIZED=1 keeps every network on the general library.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from typing import Optional

from . import _lib
from .spec import ScenarioSpec

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SPEC_DIR = os.path.join(CSRC, "spec")
_DT = {"int32": 0, "float32": 1, "float64": 2, "i32": 0, "f32": 1, "f64": 2}
_loaded = {}


def tables(spec: ScenarioSpec) -> dict:
    """The tables `lower()` derives from an sn_spec (csrc/supplynet_kernels.cu) for a tree of up to eight nodes, eight
    entries each; phantom entries (k >= n_facilities) as there: external supplier, rank 0, lead times 2,2,2,1,1.. / 2.
    Trees of five to eight nodes are compiled with eight facilities per thread (-DSN_KF=8)."""
    if not spec.is_tree or spec.n_facilities > 8:
        raise ValueError("specialisation is offered for distribution trees of up to eight nodes")
    nf = spec.n_facilities
    sup = [spec.suppliers[k] if k < nf else -1 for k in range(8)]
    rank = [spec.customer_rank[k] if (k < nf and sup[k] >= 0) else 0 for k in range(8)]
    lat = []
    for j in range(8):
        best = -1
        for c in range(nf):
            if j < nf and sup[c] == j and (best < 0 or spec.order_pos[c] > spec.order_pos[best]):
                best = c
        lat.append(best)
    std_li = (2, 2, 2, 1, 1, 1, 1, 1)
    li = [spec.info_leadtime[k] if k < nf else std_li[k] for k in range(8)]
    ls = [spec.ship_leadtime[k] if k < nf else 2 for k in range(8)]
    return dict(nf=nf, kf=4 if nf <= 4 else 8, sup=sup, rank=rank, lat=lat, li=li, ls=ls, max_rank=max(rank))


def key(spec: ScenarioSpec, dtype: str) -> str:
    t = tables(spec)
    text = f"{t['nf']}|{t['kf']}|{t['sup']}|{t['rank']}|{t['lat']}|{t['li']}|{t['ls']}|{_DT[dtype]}|abi{_lib.ABI_VERSION}"
    return hashlib.sha1(text.encode()).hexdigest()[:12]


def lib_path(spec: ScenarioSpec, dtype: str) -> str:
    return os.path.join(SPEC_DIR, f"libsupplynet_tree_{key(spec, dtype)}.so")


def build(spec: ScenarioSpec, dtype: str = "int32", force: bool = False) -> str:
    """nvcc -DSN_STATIC_TOPO ... -> csrc/spec/libsupplynet_tree_<key>.so (about 20 s); returns the path."""
    out = lib_path(spec, dtype)
    src = os.path.join(CSRC, "supplynet_kernels.cu")
    newest = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    if not force and os.path.isfile(out) and os.path.getmtime(out) >= newest:
        return out
    os.makedirs(SPEC_DIR, exist_ok=True)
    t = tables(spec)
    defs = ["-DSN_STATIC_TOPO", "-DSN_TREE_ONLY", f"-DSN_ONLY_DTYPE={_DT[dtype]}", f"-DSN_ST_NF={t['nf']}",
            f"-DSN_ST_MAX_RANK={t['max_rank']}", f"-DSN_KF={t['kf']}"]
    for name, vals in (("SUP", t["sup"]), ("RANK", t["rank"]), ("LAT", t["lat"]), ("LI", t["li"]), ("LS", t["ls"])):
        defs += [f"-DSN_ST_{name}{k}={vals[k]}" for k in range(8)]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "-diag-suppress", "177", *defs, *os.environ.get("SN_SPEC_EXTRA_FLAGS", "").split(), "-shared", "-o", out, src]
    subprocess.run(cmd, check=True, cwd=CSRC)
    return out


def load(spec: ScenarioSpec, dtype: str) -> Optional[C.CDLL]:
    """The specialised library for this network and quantity type if it has been built, else None."""
    if not spec.is_tree or spec.n_facilities > 8 or os.environ.get("SN_NO_SPECIALIZED") == "1":
        return None
    path = lib_path(spec, dtype)
    if path in _loaded:
        return _loaded[path]
    lib = None
    if not os.path.isfile(path) and os.environ.get("SN_SPECIALIZE") == "auto":
        import shutil
        if shutil.which(os.environ.get("NVCC", "nvcc")):       # 7-10 s of nvcc once per network and quantity type
            build(spec, dtype)
    if os.path.isfile(path):
        lib = C.CDLL(path)
        for name in ("sn_abi_version", "sn_last_error", "sn_spec_validate", "sn_step_ex", "sn_rollout", "sn_episode"):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = _lib.SIGNATURES[name]
        c_spec = spec.to_c()
        if lib.sn_abi_version() != _lib.ABI_VERSION or lib.sn_spec_validate(C.byref(c_spec), _DT[dtype]) != 0:
            lib = None                         # stale build (other ABI / other tables): ignore it
    _loaded[path] = lib
    return lib


def main(argv=None):
    import argparse
    from .spec import specs_from_config
    from .utils.graph import read_yaml
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("network", help="path of a network_configs/*.yaml description")
    ap.add_argument("--dtype", default="int32", choices=["int32", "float32", "float64"])
    ap.add_argument("--agents", nargs="*", default=None, help="agent-managed facilities (default: all)")
    args = ap.parse_args(argv)
    specs = specs_from_config(read_yaml(args.network), args.agents, True, 100)
    print(build(specs[0], args.dtype))


if __name__ == "__main__":
    main()
