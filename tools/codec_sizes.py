"""Whole-codec image -> .bin container -> image on the library at the north star's image sizes (one image at a time like
the reference's compress_and_decompress.py): wall-clock of encode_image / decode_image incl. the range coder and all
host<->device copies, container size, and decode == the forward pass's x_hat."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dcae_b200 import DCAECodec, container
from dcae_b200.params import init_entropy_params
from dcae_b200.transforms import init_transform_params
P = dict(init_entropy_params(0, "lively")); P.update(init_transform_params(0))
codec = DCAECodec(P)
codec.update()
for H, W in ((512, 768), (1365, 2048), (2160, 3840)):
    x = torch.rand(1, 3, H, W, generator=torch.Generator().manual_seed(H)).cuda()
    for _ in range(2):
        blob = codec.encode_image(x)
    codec.decode_image(blob)
    torch.cuda.reset_peak_memory_stats()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    blob = codec.encode_image(x)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    x_hat = codec.decode_image(blob)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    xp, padding = container.pad(x)
    want = container.crop(codec.forward(xp)["x_hat"], padding).clamp(0, 1)
    print(f"{W}x{H}: encode {1e3 * (t1 - t0):.1f} ms, decode {1e3 * (t2 - t1):.1f} ms, {len(blob)} bytes, decode == forward: {bool(torch.equal(x_hat, want))}, "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
