"""CPU: the oracle restatement must reproduce the outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import pytest
import torch

from _util import GOLDEN_CASES, load_golden, rel_err, mismatch_rate
from oracle.entropy_model import SliceLoopOracle


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_reproduces_reference_slice_loops(case, lively_params):
    g = load_golden(case)
    assert int(g["seed_params"]) == 7
    orc = SliceLoopOracle(lively_params)
    y, ls, lm = g["y"], g["latent_scales"], g["latent_means"]

    y_hat, means, scales, lik = orc.forward(y, ls, lm)
    # same math, same library, possibly different thread split of the CPU GEMMs -> tiny tolerance
    assert rel_err(means, g["means"]) < 2e-6
    assert rel_err(scales, g["scales"]) < 2e-6
    assert rel_err(y_hat, g["y_hat"]) < 2e-6 or mismatch_rate(torch.round(y_hat), torch.round(g["y_hat"])) < 1e-3
    assert rel_err(lik, g["lik"]) < 1e-4

    sym, idx, y_hat_c, _, _ = orc.compress(y, ls, lm)
    assert mismatch_rate(sym, g["symbols"]) < 1e-4
    assert mismatch_rate(idx, g["indexes"]) < 1e-4

    if "dec_indexes" in g:
        dec_y_hat, dec_idx = orc.decompress(ls, lm, lambda i, index: g["symbols"][i].float())
        assert mismatch_rate(dec_idx, g["dec_indexes"]) < 1e-4
        assert rel_err(dec_y_hat.clamp(0, 1), g["dec_y_hat"]) < 1e-5   # reference clamps x_hat (:908)
