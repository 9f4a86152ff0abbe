"""CPU: the C-ABI library loads and exports every symbol include/dcae_b200.h declares (no compute)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from dcae_b200 import _lib
    return _lib.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "dcae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcae_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from dcae_b200 import _lib
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dcae_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in dcae_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_range_coder_library_exports_its_header():
    """include/dcae_rans.h: every declared symbol is exported by libdcae_rans.so and bound in dcae_b200/ans.py."""
    import __graft_entry__ as ge
    ge.build()
    from dcae_b200 import ans
    src = open(os.path.join(ROOT, "include", "dcae_rans.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(dcae_[a-z0-9_]+)\s*\(", src)))
    lib = ans.load()
    assert len(names) >= 10 and set(names) == set(ans.SIGNATURES)
    for n in names:
        assert hasattr(lib, n), n
    import ctypes as C
    import subprocess
    import tempfile
    prog = '#include <stdio.h>\n#include "dcae_rans.h"\nint main(){printf("%zu\\n", sizeof(dcae_rans_tables));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "t"), os.path.join(td, "t.c")], check=True)
        assert int(subprocess.run([os.path.join(td, "t")], capture_output=True, text=True, check=True).stdout) == C.sizeof(ans.RansTables)


def test_version_and_error_channel(lib):
    assert lib.dcae_version() == 100
    assert lib.dcae_gc_num_partials(24576, 64) == 1184
    assert lib.dcae_gc_num_partials(4, 64) == 1
    # argument validation happens before any CUDA call, so it is testable without a GPU
    from dcae_b200 import _lib
    a = _lib.GcArgs()
    a.mode = 7
    assert lib.dcae_gc_fused(a, None) == -1
    assert b"bad mode" in lib.dcae_last_error()
    assert lib.dcae_slice_loop_workspace_bytes(16, 32, 48) > 24576 * 17000 * 4


def test_ctypes_struct_layout_matches_header(lib):
    """sizeof of the ctypes mirrors must match what the C compiler sees."""
    import ctypes as C
    import subprocess
    import tempfile
    from dcae_b200 import _lib
    prog = '#include <stdio.h>\n#include "dcae_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(dcae_gc_args), sizeof(dcae_operand), sizeof(dcae_epilogue), sizeof(dcae_weight), sizeof(dcae_slice_weights), sizeof(dcae_gc_bwd_args));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "t"), os.path.join(td, "t.c")], check=True)
        sizes = [int(v) for v in subprocess.run([os.path.join(td, "t")], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.GcArgs), C.sizeof(_lib.Operand), C.sizeof(_lib.Epilogue), C.sizeof(_lib.Weight), C.sizeof(_lib.SliceWeights),
                     C.sizeof(_lib.GcBwdArgs)]


def test_no_cpu_fallback():
    import torch
    from dcae_b200 import _lib
    from dcae_b200.gaussian_conditional import GaussianConditional
    gc = GaussianConditional(None)
    with pytest.raises(_lib.DcaeError):
        gc.quantize(torch.zeros(1, 4), "symbols", torch.zeros(1, 4))
    if not torch.cuda.is_available():
        from dcae_b200.entropy_model import EntropySliceLoop
        with pytest.raises((_lib.DcaeError, RuntimeError, AssertionError)):
            EntropySliceLoop({}, device="cuda:0")


def test_product_never_imports_the_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "dcae_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dp, f)).read().replace("cpu_baseline", ""), f
