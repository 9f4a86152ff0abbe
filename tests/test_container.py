"""The bitstream container and image padding (dcae_b200/container.py) against the reference's own functions
(`/root/reference/compress_and_decompress.py:48-71, 110-148`, executed from the reference file when it is available) and
against hand-written known answers of the byte layout."""
import ast
import os
import struct

import pytest
import torch

from dcae_b200 import container

REF = os.path.join(os.environ.get("DCAE_REFERENCE_ROOT", "/root/reference"), "compress_and_decompress.py")


def _reference_functions():
    """Only the five pure functions of the reference script (the file itself imports packages that are absent here)."""
    import torch.nn.functional as F
    tree = ast.parse(open(REF).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("pad", "crop", "save_bin", "calculate_padding", "read_bin")]
    ns = {"F": F, "torch": torch, "os": os, "struct": struct}
    exec(compile(ast.Module(body=keep, type_ignores=[]), REF, "exec"), ns)
    return ns


def test_known_answer_layout():
    blob = container.pack_bin([[b"\x01\x02\x03"], [b"\xff"]], (513, 770))
    assert blob == b"\x02\x01" + b"\x03\x02" + b"\x00\x00\x00\x03" + b"\x01\x02\x03" + b"\x00\x00\x00\x01" + b"\xff"
    strings, z_shape, padding, size = container.unpack_bin(blob)
    assert strings == [[b"\x01\x02\x03"], [b"\xff"]] and size == (513, 770)
    assert z_shape == [640 // 64, 896 // 64] and padding == (63, 63, 63, 64)
    for cut in (5, 10, len(blob) - 1):
        with pytest.raises(ValueError):
            container.unpack_bin(blob[:cut])
    with pytest.raises(ValueError):
        container.pack_bin([[b""], [b""]], (70000, 1))


@pytest.mark.parametrize("hw", [(512, 768), (1365, 2048), (1, 1), (129, 255), (2160, 3840)])
def test_pad_crop_round_trip(hw):
    x = torch.rand(1, 3, *hw) if hw[0] * hw[1] < 1e6 else torch.zeros(1, 3, *hw)
    xp, padding = container.pad(x)
    assert xp.shape[2] % 128 == 0 and xp.shape[3] % 128 == 0 and xp.shape[2] - hw[0] < 128 and xp.shape[3] - hw[1] < 128
    assert torch.equal(container.crop(xp, padding), x)
    assert container.calculate_padding(*hw)[1] == padding


@pytest.mark.skipif(not os.path.isfile(REF), reason="reference script not available")
def test_against_the_reference_functions(tmp_path):
    ref = _reference_functions()
    x = torch.rand(1, 3, 300, 517, generator=torch.Generator().manual_seed(3))
    xp_r, pad_r = ref["pad"](x, 128)
    xp, pad_ = container.pad(x)
    assert torch.equal(xp, xp_r) and tuple(pad_) == tuple(pad_r)
    assert torch.equal(container.crop(xp, pad_), ref["crop"](xp_r, pad_r))
    assert container.calculate_padding(300, 517) == ref["calculate_padding"](300, 517)
    strings = [[bytes(range(200)) * 3], [b"zz-top"]]
    ref["save_bin"](strings, x.shape[-2:], "img.png", str(tmp_path))            # writes <tmp>/bin/img.bin
    theirs = open(os.path.join(str(tmp_path), "bin", "img.bin"), "rb").read()
    assert theirs == container.pack_bin(strings, x.shape[-2:])
    mine = os.path.join(str(tmp_path), "mine.bin")
    container.save_bin(strings, x.shape[-2:], mine)
    s_r, shape_r, pad2_r = ref["read_bin"](mine)                                # the reference reads our file
    s_m, shape_m, pad2_m = container.read_bin(os.path.join(str(tmp_path), "bin", "img.bin"))     # and we read theirs
    assert s_r == s_m == strings and list(shape_r) == list(shape_m) and tuple(pad2_r) == tuple(pad2_m)
