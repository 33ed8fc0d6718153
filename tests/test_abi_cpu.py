"""CPU-only checks of the drop-in boundary: the shared library loads, exports
every symbol include/spai_b200.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

import conftest

ROOT = conftest.ROOT


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from gflownet_spai_b200 import _lib
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "spai_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spai_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from gflownet_spai_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in spai_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared


def test_abi_version_and_error_string(lib):
    assert lib.spai_abi_version() == 2
    assert isinstance(lib.spai_last_error(), bytes)


def test_invalid_arguments_are_rejected_without_gpu(lib):
    out = ctypes.c_void_p()
    st = lib.spai_ctx_create(0, -5, 0, None, None, None, 0, None, None, None, ctypes.byref(out))
    assert st == 1 and not out.value          # SPAI_ERR_INVALID
    assert b"invalid" in lib.spai_last_error()


def test_product_path_fails_loudly_without_cuda(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from gflownet_spai_b200.env import PreconditionerEnv
    from gflownet_spai_b200._lib import SpaiError
    i = torch.tensor([[0, 1], [0, 1]])
    m = torch.sparse_coo_tensor(i, torch.ones(2), (2, 2))
    with pytest.raises(SpaiError):
        PreconditionerEnv(2, m, m)


def test_non_sparse_input_raises_value_error(lib):
    import torch
    from gflownet_spai_b200.env import PreconditionerEnv
    with pytest.raises(ValueError):
        PreconditionerEnv(2, torch.eye(2), torch.eye(2))
