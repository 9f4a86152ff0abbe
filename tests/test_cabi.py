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


def test_transform_operators_validate_their_arguments_before_any_launch(lib):
    """The argument checks of the transform-stack operators (include/dcae_b200.h) answer DCAE_E_INVALID with a message and
    launch nothing: callable without a GPU (pointers are never dereferenced on the host)."""
    fake = 0x1000                       # a non-null, 16-byte aligned "device pointer"
    err = lambda: lib.dcae_last_error().decode()
    # window 5 is not supported; head_dim must divide C; grid must be a multiple of the window; shift is 0 or window / 2
    assert lib.dcae_op_window_attention(fake, 288, 0, 96, 192, 96, 8, 5, 0, fake, 1, 16, 16, fake, 96, None, None) == -1 and "window" in err()
    assert lib.dcae_op_window_attention(fake, 288, 0, 96, 192, 100, 8, 8, 0, fake, 1, 16, 16, fake, 100, None, None) == -1 and "head_dim" in err()
    assert lib.dcae_op_window_attention(fake, 288, 0, 96, 192, 96, 8, 8, 0, fake, 1, 12, 16, fake, 96, None, None) == -1 and "multiple of the window" in err()
    assert lib.dcae_op_window_attention(fake, 288, 0, 96, 192, 96, 8, 8, 3, fake, 1, 16, 16, fake, 96, None, None) == -1 and "shift" in err()
    assert lib.dcae_op_window_attention(fake, 200, 0, 96, 192, 96, 8, 8, 0, fake, 1, 16, 16, fake, 96, None, None) == -1 and "leading dimension" in err()
    assert lib.dcae_op_window_attention(None, 288, 0, 96, 192, 96, 8, 8, 0, fake, 1, 16, 16, fake, 96, None, None) == -1
    # B = 0: nothing to do, no launch
    assert lib.dcae_op_window_attention(fake, 288, 0, 96, 192, 96, 8, 8, 0, fake, 0, 16, 16, fake, 96, None, None) == 0
    # space-to-depth: the slot must hold the channels and be a multiple of 4; depth-to-space: the input must hold 4 slots
    assert lib.dcae_op_space_to_depth(fake, 96, 96, 90, 1, 8, 8, fake, 384, None, None) == -1 and "Cs" in err()
    assert lib.dcae_op_space_to_depth(fake, 96, 96, 96, 1, 8, 8, fake, 380, None, None) == -1
    assert lib.dcae_op_depth_to_space(fake, 100, 96, 96, 96, 1, 8, 8, fake, 96, None, None) == -1 and "ld" in err()
    assert lib.dcae_op_depth_to_space(fake, 384, 96, 96, 64, 1, 8, 8, fake, 96, None, None) == -1
    # LayerNorm: any multiple of 4 up to 1024 now, nothing else
    assert lib.dcae_op_layernorm(fake, 96, fake, fake, 98, 10, fake, 96, None, None) == -1 and "multiple of 4" in err()
    assert lib.dcae_op_layernorm(fake, 96, fake, fake, 96, 0, fake, 96, None, None) == 0
