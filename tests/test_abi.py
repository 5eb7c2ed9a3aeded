"""The C-ABI library builds, loads and exports every symbol include/zkcheck.h declares.
No compute calls: this runs without a GPU."""
import os
import re

from zkevm_specs_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "zkcheck.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zk_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = native.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libzkcheck.so does not export {name}"
    assert sorted(native.EXPORTS) == declared


def test_introspection_without_gpu():
    lib = native.lib()
    assert lib.zk_circuit_cols(native.CIRCUIT_BYTECODE) == 12
    assert lib.zk_circuit_cols(native.CIRCUIT_STATE) == 57
    assert lib.zk_table_cols(native.TABLE_RW) == 14
    cat = native.constraint_catalogue(native.CIRCUIT_BYTECODE)
    assert len(cat) == 22 and all(cls == native.ERR_ASSERT for _, cls in cat)
    assert cat[0][0].startswith("BC_FIRST_TAG")


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.NativeError):
        native.Context(0)
