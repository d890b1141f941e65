"""CPU: the C-ABI shared library loads and exports every symbol include/whisper_b200.h declares
(no compute calls - there is no GPU here), and the product never imports the oracle."""
import ast
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from whisper_b200 import build

    return build.build_library()


def test_library_exports_every_header_symbol(built_lib):
    from whisper_b200 import _lib

    names = _lib.header_symbols()
    assert len(names) >= 8
    lib = _lib.lib()          # binds every declared symbol; raises if one is missing
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.wb200_version()
    assert lib.wb200_launch_count() == 0


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "whisper_b200")):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    mods = [node.module or ""]
                if any(m == "oracle" or m.startswith("oracle.") for m in mods):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product modules import the oracle: {bad}"


def test_ops_fail_loudly_without_cuda():
    import torch

    from whisper_b200 import ops

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        ops.layernorm(torch.zeros(2, 8, dtype=torch.bfloat16), torch.ones(8), torch.zeros(8))
