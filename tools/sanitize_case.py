"""Small f16x3 workload for compute-sanitizer (tools/gpu_sanitize.sh): the 2 x 7 x 9 golden slice loop (every kernel of
the default path: tcgen05 GEMMs single + pair, kernel 1, kernel 3, LN / depthwise / gate / transposes, pack) checked
against the reference golden, so a sanitizer-clean run is also a correct run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from _util import load_golden, mismatch_rate  # noqa: E402
from dcae_b200.entropy_model import EntropySliceLoop  # noqa: E402
from dcae_b200.params import init_entropy_params  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "slice_loop_b2_7x9"
g = load_golden(case)
eng = EntropySliceLoop(init_entropy_params(7, "lively"), device="cuda:0", math=os.environ.get("SAN_MATH", "f16x3"), lanes=1)
y, ls, lm = (g[k].cuda() for k in ("y", "latent_scales", "latent_means"))
out = eng.compress(y, ls, lm, with_likelihoods=True)
host = eng.compress_to_host(y, ls, lm)
dec = eng.decompress(ls, lm, lambda i, idx: out["symbols"][i])
torch.cuda.synchronize()
assert torch.equal(dec["y_hat"], out["y_hat"])
print("sanitize_case", case, "symbol mismatch vs golden", mismatch_rate(out["symbols"].cpu(), g["symbols"]),
      "index mismatch", mismatch_rate(out["indexes"].cpu(), g["indexes"]), "launches", eng.last_launches)
