"""Hot-path parameter inventory of the reference DCAE entropy model + a deterministic init recipe.

The state-dict key names and shapes below are the compatibility contract with reference
checkpoints (`/root/reference/models/dcae.py:451-477` for the dictionary
cross-attention module, `:531-539` for `dt`, `:584-611` for the cc_mean / cc_scale / lrp stacks).
`init_entropy_params` draws every tensor from a seeded CPU generator in the fixed order of
`entropy_param_spec()`, so the same weights can be rebuilt on any box without shipping 333 MB.

profile="default" follows PyTorch's default init of the reference modules (Linear / Conv2d:
U(-1/sqrt(fan_in), +1/sqrt(fan_in)) for weight and bias; LayerNorm 1/0; `dt ~ N(0,1)` dcae.py:534;
head scale 1 dcae.py:457; residual Scale 1 dcae.py:333).
profile="lively" additionally randomises the affine/scale vectors and widens the last cc_scale
layer so that symbols and CDF indexes cover many bins; used for parity fixtures (any state dict
is a valid input to the reference).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

NUM_SLICES = 5
M_LATENT = 320
SLICE_CH = M_LATENT // NUM_SLICES          # 64
DICT_NUM = 128
HEAD_NUM = 20
HEAD_DIM = 32
DICT_DIM = HEAD_NUM * HEAD_DIM             # 640
MLP_HIDDEN = 4 * DICT_DIM                  # fc1 out = 2560, split 1280/1280
CC_HID1, CC_HID2 = 224, 128


def cq(i: int) -> int:
    """query channels of slice i: latent_scales + latent_means + i previous y_hat slices (dcae.py:645)."""
    return 2 * M_LATENT + SLICE_CH * i


def cs(i: int) -> int:
    """support channels: query + dict_info (dcae.py:647)."""
    return cq(i) + M_LATENT


def cl(i: int) -> int:
    """lrp support channels: support + y_hat_slice (dcae.py:661)."""
    return cs(i) + SLICE_CH


def entropy_param_spec():
    """[(key, shape, kind, fan_in)] in a fixed order. kind in {w, b, ln_w, ln_b, ones, normal}."""
    spec = [("dt", (DICT_NUM, DICT_DIM), "normal", 0)]
    D = DICT_DIM
    for i in range(NUM_SLICES):
        p = f"dt_cross_attention.{i}."

        def lin(name, out_f, in_f, shape=None):
            spec.append((p + name + ".weight", shape or (out_f, in_f), "w", in_f))
            spec.append((p + name + ".bias", (out_f,), "b", in_f))

        def ln(name):
            spec.append((p + name + ".weight", (D,), "ln_w", 0))
            spec.append((p + name + ".bias", (D,), "ln_b", 0))

        spec.append((p + "scale", (HEAD_NUM, 1, 1), "head_scale", 0))
        lin("x_trans", D, cq(i))
        ln("ln_scale")
        lin("msa.s", D, D, (D, D, 1, 1))
        spec.append((p + "msa.spatial_atte.conv1.weight", (1, 2, 7, 7), "w", 2 * 49))
        for j in range(3):
            q = f"msa.dense.conv_layers.{j}.1."
            lin(q + "in_trans", D, D, (D, D, 1, 1))
            spec.append((p + q + "dw_conv.weight", (D, 1, 3, 3), "w", 9))
            spec.append((p + q + "dw_conv.bias", (D,), "b", 9))
            lin(q + "out_trans", D, D, (D, D, 1, 1))
        lin("msa.dense.proj", D, 4 * D, (D, 4 * D, 1, 1))
        ln("lnx")
        lin("q_trans", D, D)
        ln("dict_ln")
        lin("k", D, D)
        lin("linear", D, D)
        ln("ln_mlp")
        lin("mlp.fc1", MLP_HIDDEN, D)
        spec.append((p + "mlp.dwconv.dwconv.weight", (MLP_HIDDEN // 2, 1, 3, 3), "w", 9))
        spec.append((p + "mlp.dwconv.dwconv.bias", (MLP_HIDDEN // 2,), "b", 9))
        lin("mlp.fc2", D, MLP_HIDDEN // 2)
        lin("output_trans.0", M_LATENT, D)
        for r in (1, 2, 3):
            spec.append((p + f"res_scale_{r}.scale", (D,), "res_scale", 0))
    for fam, cin in (("cc_mean_transforms", cs), ("cc_scale_transforms", cs), ("lrp_transforms", cl)):
        for i in range(NUM_SLICES):
            chans = [cin(i), CC_HID1, CC_HID2, SLICE_CH]
            for li, layer in enumerate((0, 2, 4)):
                ci, co = chans[li], chans[li + 1]
                spec.append((f"{fam}.{i}.{layer}.weight", (co, ci, 3, 3), "w", ci * 9))
                spec.append((f"{fam}.{i}.{layer}.bias", (co,), "b", ci * 9))
    return spec


def entropy_param_shapes() -> "OrderedDict[str, tuple]":
    return OrderedDict((k, s) for k, s, _, _ in entropy_param_spec())


def init_entropy_params(seed: int = 0, profile: str = "default") -> "OrderedDict[str, torch.Tensor]":
    """Deterministic CPU fp32 state dict of every hot-path tensor (83.2 M parameters)."""
    if profile not in ("default", "lively"):
        raise ValueError(profile)
    lively = profile == "lively"
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    out = OrderedDict()

    def uni(shape, lo, hi):
        return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    for key, shape, kind, fan_in in entropy_param_spec():
        if kind in ("w", "b"):
            bound = 1.0 / math.sqrt(fan_in)
            t = uni(shape, -bound, bound)
        elif kind == "normal":
            t = torch.randn(shape, generator=g, dtype=torch.float32)
        elif kind == "ln_w":
            t = uni(shape, 0.5, 1.5) if lively else torch.ones(shape)
        elif kind == "ln_b":
            t = uni(shape, -0.2, 0.2) if lively else torch.zeros(shape)
        elif kind == "res_scale":
            t = uni(shape, 0.5, 1.5) if lively else torch.ones(shape)
        elif kind == "head_scale":
            t = uni(shape, 0.5, 2.0) if lively else torch.ones(shape)
        else:
            raise AssertionError(kind)
        out[key] = t
    if lively:
        # Spread the predicted scales over the whole table (with plain random init
        # nearly every scale is below the 0.11 bound and every index is 0) and make the means
        # large enough that round(y - mu) is not simply round(y).
        for i in range(NUM_SLICES):
            out[f"cc_scale_transforms.{i}.4.weight"] *= 40.0
            out[f"cc_scale_transforms.{i}.4.bias"] = uni((SLICE_CH,), 0.0, 4.0)
            out[f"cc_mean_transforms.{i}.4.weight"] *= 20.0
            out[f"cc_mean_transforms.{i}.4.bias"] = uni((SLICE_CH,), -2.0, 2.0)
            out[f"lrp_transforms.{i}.4.weight"] *= 10.0
    return out
