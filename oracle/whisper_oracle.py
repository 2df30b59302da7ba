"""CPU oracle for the PPG extractor (TEST INFRASTRUCTURE): torch-CPU fp32 restatement of the
truncated Whisper AudioEncoder the reference runs (whisper/model.py:132-163 with the loader's
surgery, whisper/inference.py:11-29: decoder deleted, last quarter of the encoder blocks deleted,
ln_post kept).  Consumes the reference checkpoint format {"dims", "model_state_dict"}.

Parity status: the reference ships no golden vectors for this path, so the pin is the reference code
itself: `oracle/make_golden.py:whisper_case` builds the unmodified `whisper.model.Whisper` through the
loader surgery, checks this restatement against `model.encoder(mel)` (max-abs 0.0) and writes
`tests/golden/whisper_*.npz`; `tests/test_oracle_cpu.py::test_whisper_oracle_matches_golden` re-checks the
fixtures everywhere and `::test_whisper_oracle_matches_reference_live` repeats the live comparison where
/root/reference exists."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """whisper/model.py:48-54"""
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1)


def kept_layers(dims: dict) -> int:
    """whisper/inference.py:17-19: del encoder.blocks[-(n // 4):]"""
    n = dims["n_audio_layer"]
    return n - n // 4


def attention(sd, p, x, n_head):
    """MultiHeadAttention.forward/qkv_attention, whisper/model.py:66-101 (self-attention, no mask)."""
    q = F.linear(x, sd[p + ".query.weight"], sd[p + ".query.bias"])
    k = F.linear(x, sd[p + ".key.weight"])
    v = F.linear(x, sd[p + ".value.weight"], sd[p + ".value.bias"])
    b, t, d = q.shape
    scale = (d // n_head) ** -0.25
    q = q.view(b, t, n_head, -1).permute(0, 2, 1, 3) * scale
    k = k.view(b, t, n_head, -1).permute(0, 2, 3, 1) * scale
    v = v.view(b, t, n_head, -1).permute(0, 2, 1, 3)
    w = F.softmax((q @ k).float(), dim=-1).to(q.dtype)
    o = (w @ v).permute(0, 2, 1, 3).flatten(start_dim=2)
    return F.linear(o, sd[p + ".out.weight"], sd[p + ".out.bias"])


def audio_encoder(ckpt: dict, mel: torch.Tensor, stages: dict | None = None) -> torch.Tensor:
    """AudioEncoder.forward, whisper/model.py:144-163, on mel [B, n_mels, n] -> [B, ceil(n/2), D]."""
    dims, sd = ckpt["dims"], ckpt["model_state_dict"]
    D, H = dims["n_audio_state"], dims["n_audio_head"]
    with torch.no_grad():
        x = F.gelu(F.conv1d(mel, sd["encoder.conv1.weight"], sd["encoder.conv1.bias"], padding=1))
        x = F.gelu(F.conv1d(x, sd["encoder.conv2.weight"], sd["encoder.conv2.bias"], stride=2, padding=1))
        x = x.permute(0, 2, 1)
        assert x.shape[1] <= dims["n_audio_ctx"], "incorrect audio shape"
        x = x + sinusoids(dims["n_audio_ctx"], D)[: x.shape[1]]
        if stages is not None:
            stages["stem"] = x
        for i in range(kept_layers(dims)):
            b = f"encoder.blocks.{i}"
            h = F.layer_norm(x, (D,), sd[b + ".attn_ln.weight"], sd[b + ".attn_ln.bias"])
            x = x + attention(sd, b + ".attn", h, H)
            h = F.layer_norm(x, (D,), sd[b + ".mlp_ln.weight"], sd[b + ".mlp_ln.bias"])
            h = F.gelu(F.linear(h, sd[b + ".mlp.0.weight"], sd[b + ".mlp.0.bias"]))
            x = x + F.linear(h, sd[b + ".mlp.2.weight"], sd[b + ".mlp.2.bias"])
            if stages is not None:
                stages[f"block{i}"] = x
        return F.layer_norm(x, (D,), sd["encoder.ln_post.weight"], sd["encoder.ln_post.bias"])


# ------------------------------------------------------------------ log-mel front end (whisper/audio.py)
SAMPLE_RATE, N_FFT, HOP_LENGTH, N_MELS = 16000, 400, 160, 80


def slaney_mel_filterbank(n_mels: int = N_MELS, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """The matrix whisper/audio.py:54-65 obtains from `librosa.filters.mel(sr=16000, n_fft=400, n_mels=80)`
    (librosa defaults: Slaney mel scale — linear below 1 kHz at 200/3 Hz per mel, logarithmic above with
    27 steps per factor 6.4 — triangular filters between consecutive mel-spaced corner frequencies,
    each scaled by 2 / (f_hi - f_lo)).  librosa is a requirements.txt dependency that is absent from
    this image and the reference ships no copy of the matrix: the filterbank is restated from the
    published definition and pinned against an independent third-party implementation that IS in the
    image — transformers 5.5 `audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney")` and
    `WhisperFeatureExtractor._np_extract_fbank_features` (tests/test_oracle_cpu.py: 1e-9 / 1e-6)."""
    def hz_to_mel(f):
        return f / (200.0 / 3) if f < 1000.0 else 15.0 + math.log(f / 1000.0) / (math.log(6.4) / 27.0)

    def mel_to_hz(m):
        return (200.0 / 3) * m if m < 15.0 else 1000.0 * math.exp((math.log(6.4) / 27.0) * (m - 15.0))

    top = hz_to_mel(sr / 2.0)
    corners = [mel_to_hz(top * i / (n_mels + 1)) for i in range(n_mels + 2)]
    nb = n_fft // 2 + 1
    fb = np.zeros((n_mels, nb), dtype=np.float64)
    for m in range(n_mels):
        lo, mid, hi = corners[m], corners[m + 1], corners[m + 2]
        for k in range(nb):
            f = k * sr / n_fft
            tri = min((f - lo) / (mid - lo), (hi - f) / (hi - mid))
            if tri > 0.0:
                fb[m, k] = tri * 2.0 / (hi - lo)
    return fb.astype(np.float32)


def log_mel_spectrogram(audio: torch.Tensor, n_mels: int = N_MELS) -> torch.Tensor:
    """whisper/audio.py:68-100: Hann STFT (400 / 160, torch.stft defaults = centred, reflect padding),
    squared magnitude without the last frame, mel projection, log10 of the 1e-10 clamp, `max - 8`
    floor over the whole chunk, (x + 4) / 4.  audio [n_samples] -> [n_mels, n_samples // 160]."""
    window = torch.hann_window(N_FFT)
    stft = torch.stft(audio.float(), N_FFT, HOP_LENGTH, window=window, return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    mel_spec = torch.from_numpy(slaney_mel_filterbank(n_mels)) @ magnitudes
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0
