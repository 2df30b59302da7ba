"""CPU oracle for the HuBERT-Soft content encoder (TEST INFRASTRUCTURE): torch-CPU fp32 restatement of
`HubertSoft.units` (hubert/hubert_model.py:64-72 -> encode :39-48: FeatureExtractor :75-95, FeatureProjection :98-109,
PositionalConvEmbedding :112-128, LayerNorm, 12 x nn.TransformerEncoderLayer(768, 12, 3072, gelu, batch_first; post-LN)
:131-153, proj) over the reference's state-dict format.

Parity status: the reference ships no golden vectors for this path, so the pin is the reference code itself:
`oracle/make_golden.py:hubert_case` instantiates the unmodified `hubert_model.HubertSoft`, loads the synthetic state dict,
checks this restatement against `model.units(wav)` and writes `tests/golden/hubert_*.npz`;
`tests/test_oracle_cpu.py::test_hubert_oracle_matches_golden` re-checks the fixtures everywhere and
`::test_hubert_oracle_matches_reference_live` repeats the live comparison where /root/reference exists."""
from __future__ import annotations

import torch
import torch.nn.functional as F

CONV_KERNELS = (3, 3, 3, 3, 2, 2)   # conv1 .. conv6, all stride 2 (hubert_model.py:80-85)


def frames(n_samples: int) -> int:
    """Output frames of `units` for n_samples of 16 kHz audio (pad 40 + 40, k10 s5, then the six stride-2 convs)."""
    t = (n_samples + 80 - 10) // 5 + 1
    for k in CONV_KERNELS:
        t = (t - k) // 2 + 1
    return t


def feature_extractor(sd, wav):
    """hubert_model.py:87-95.  wav [B, 1, N] (already padded)."""
    p = "feature_extractor."
    x = F.conv1d(wav, sd[p + "conv0.weight"], stride=5)
    x = F.gelu(F.group_norm(x, 512, sd[p + "norm0.weight"], sd[p + "norm0.bias"]))
    for i in range(1, 7):
        x = F.gelu(F.conv1d(x, sd[p + f"conv{i}.weight"], stride=2))
    return x


def positional(sd, x):
    """hubert_model.py:124-128 with weight_norm(dim=2) folded (:122).  x [B, T, 768]."""
    p = "positional_embedding.conv."
    w = torch._weight_norm(sd[p + "weight_v"], sd[p + "weight_g"], 2)
    y = F.conv1d(x.transpose(1, 2), w, sd[p + "bias"], padding=64, groups=16)
    return F.gelu(y[:, :, :-1]).transpose(1, 2)


def encoder_layer(sd, p, x, n_head=12):
    """nn.TransformerEncoderLayer, norm_first=False: x = LN1(x + SA(x)); x = LN2(x + W2 gelu(W1 x))."""
    B, T, D = x.shape
    qkv = F.linear(x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
    q, k, v = [t.view(B, T, n_head, D // n_head).transpose(1, 2) for t in qkv.split(D, dim=-1)]
    w = torch.softmax(q @ k.transpose(-1, -2) / (D // n_head) ** 0.5, dim=-1)
    a = (w @ v).transpose(1, 2).reshape(B, T, D)
    a = F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
    x = F.layer_norm(x + a, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    f = F.linear(F.gelu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])), sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return F.layer_norm(x + f, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"])


@torch.no_grad()
def units(sd, wav, taps: dict | None = None):
    """HubertSoft.units (hubert_model.py:68-72): wav [B, 1, N] fp32 -> [B, T, 256].  taps collects intermediates."""
    n_layer = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.layers."))
    x = feature_extractor(sd, F.pad(wav, (40, 40)))
    if taps is not None:
        taps["features"] = x.clone()
    x = x.transpose(1, 2)
    x = F.layer_norm(x, (512,), sd["feature_projection.norm.weight"], sd["feature_projection.norm.bias"])
    x = F.linear(x, sd["feature_projection.projection.weight"], sd["feature_projection.projection.bias"])
    if taps is not None:
        taps["projected"] = x.clone()
    x = x + positional(sd, x)
    x = F.layer_norm(x, (768,), sd["norm.weight"], sd["norm.bias"])
    if taps is not None:
        taps["embedded"] = x.clone()
    for i in range(n_layer):
        x = encoder_layer(sd, f"encoder.layers.{i}.", x)
        if taps is not None and i == 0:
            taps["layer0"] = x.clone()
    if taps is not None:
        taps["encoded"] = x.clone()
    return F.linear(x, sd["proj.weight"], sd["proj.bias"])
