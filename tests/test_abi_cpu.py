"""CPU checks of the C-ABI library: it loads, exports every symbol include/avddpg_b200.h declares, the
ctypes struct layouts match, host-only entry points agree with the oracle, and compute entry points fail
loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from avddpg_b200 import _lib
from avddpg_b200.config import Config, env_params_from_config
from oracle import platoon_np as onp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "avddpg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(avd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "binding table and header disagree"


def test_struct_layouts_match():
    lib = _lib.load()
    for which, st in enumerate((_lib.EnvParams, _lib.EnvIO, _lib.Clock, _lib.NetDims, _lib.LearnIO)):
        assert lib.avd_sizeof(which) == C.sizeof(st)
    assert lib.avd_abi_version() == _lib.ABI_VERSION


@pytest.mark.parametrize("over", [dict(), dict(method="exact"), dict(method="exact", pl_leader_tau=0.25, timegap=1.3, dyn_coeff=0.15)])
def test_build_matrices_matches_oracle(over):
    conf = Config(**over)
    prm = env_params_from_config(conf, 5)
    mats = onp.follower_matrices(onp.EnvParams.from_config(conf), 5)
    for m, (A, B, Cc) in enumerate(mats):
        assert np.array_equal(np.array(list(prm.A[m]), dtype=np.float32).reshape(4, 4), A.astype(np.float32))
        assert np.array_equal(np.array(list(prm.B[m]), dtype=np.float32), B.astype(np.float32))
        assert np.array_equal(np.array(list(prm.C[m]), dtype=np.float32), Cc.astype(np.float32))


def test_error_convention():
    conf = Config()
    with pytest.raises(ValueError):
        env_params_from_config(conf, 17)
    with pytest.raises(ValueError):
        env_params_from_config(Config(model="ModelC"), 2)
    prm = env_params_from_config(conf, 2)
    prm.M = 99
    rc = _lib.load().avd_env_build_matrices(C.byref(prm), 0, 0.1, 1.0, 0.1, 0.1)
    assert rc == -1 and b"outside" in _lib.load().avd_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    from avddpg_b200.environment import BatchedPlatoons
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        BatchedPlatoons(4, 2, Config())
    with pytest.raises(RuntimeError):
        _lib.require_device()
