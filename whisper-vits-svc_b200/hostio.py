"""Host loop and wire formats of the inference CLIs (svc_inference.py of the reference).

Arithmetic-free glue, restated from the reference so the drop-in produces byte-compatible
outputs: checkpoint loading (svc_inference.py:61-74), feature preparation (:172-199), the
2500-frame chunk loop with +-10-frame overlap discard (:77-134), the pitch CSV format
(pitch/inference.py:102-119).  The only structural change: chunks of equal length are run as
batch items of ONE device call instead of a Python loop of batch-1 calls (chunks are
independent: the overlap is discarded, not blended).
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np
import torch


def load_svc_model(checkpoint_path: str, model):
    """svc_inference.py:61-74: key-by-key copy; missing keys are printed and keep their default."""
    assert os.path.isfile(checkpoint_path), checkpoint_path
    ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    saved = ckpt["model_g"]
    state = model.state_dict()
    new = {}
    for k, v in state.items():
        if k in saved:
            new[k] = saved[k]
        else:
            print("%s is not in the checkpoint" % k)
            new[k] = v
    model.load_state_dict(new)
    return model


def save_svc_model(model, checkpoint_path: str):
    """svc_export.py:40-45: {'model_g': state_dict}."""
    torch.save({"model_g": {k: v.detach().cpu() for k, v in model.state_dict().items()}}, checkpoint_path)


def load_csv_pitch(path: str) -> List[int]:
    """pitch/inference.py:113-119: last comma-separated field of every line, as int."""
    pitch = []
    with open(path, "r", encoding="utf-8") as f:
        for line in f.readlines():
            pitch.append(int(line.strip().split(",")[-1]))
    return pitch


def save_csv_pitch(pitch: Sequence[float], path: str):
    """pitch/inference.py:102-110: '{m}m {s}s {ms:3d},{int Hz}' every 10 ms."""
    with open(path, "w", encoding="utf-8") as f:
        for i in range(len(pitch)):
            t = i * 10
            minute = t // 60000
            seconds = (t - minute * 60000) // 1000
            millisecond = t % 1000
            print(f"{minute}m {seconds}s {millisecond:3d},{int(pitch[i])}", file=f)


def prepare_features(ppg_path: str, vec_path: str, pit_path: str, shift: int = 0):
    """svc_inference.py:172-199: 20 ms rows repeated x2, CSV pitch, optional 2^(k/12) shift."""
    ppg = torch.FloatTensor(np.repeat(np.load(ppg_path), 2, 0))
    vec = torch.FloatTensor(np.repeat(np.load(vec_path), 2, 0))
    pit = load_csv_pitch(pit_path)
    if shift != 0:
        pit = np.array(pit)
        pit = pit * (2 ** (shift / 12))
    pit = torch.FloatTensor(pit)
    return ppg, vec, pit


def chunk_plan(all_frame: int, hop_size: int, out_chunk: int = 2500, hop_frame: int = 10
               ) -> List[Tuple[int, int, int, int]]:
    """The reference's chunk arithmetic (svc_inference.py:94-115,129-131) as data:
    [(cut_s, cut_e, cut_s_out, cut_e_out)] in frames / samples; cut_e_out is the (negative) end
    index of the slice applied to the chunk's output (-1 on the last chunk, which drops one sample)."""
    plan = []
    out_index = 0
    while out_index < all_frame:
        if out_index == 0:
            cut_s, cut_s_out = 0, 0
        else:
            cut_s, cut_s_out = out_index - hop_frame, hop_frame * hop_size
        if out_index + out_chunk + hop_frame > all_frame:
            cut_e, cut_e_out = all_frame, -1
        else:
            cut_e, cut_e_out = out_index + out_chunk + hop_frame, -1 * hop_frame * hop_size
        plan.append((cut_s, cut_e, cut_s_out, cut_e_out))
        out_index += out_chunk
    return plan


def svc_infer(model, spk, pit, ppg, vec, hp, device, write_pit_wav: str | None = "svc_out_pit.wav",
              rand_ini=None, noise=None, eps_fn=None, max_batch: int = 16):
    """svc_inference.py:77-134.  Returns the float32 waveform as a numpy array of length
    n_frames*hop - 1 (the reference's last-chunk slice).  `rand_ini`/`noise` and
    `eps_fn(chunk_idx, 1, n_frames) -> [1, inter_channels, n_frames]` (one call per chunk) inject the
    reference's random draws for parity tests."""
    len_min = min(pit.size(0), vec.size(0), ppg.size(0))
    pit, vec, ppg = pit[:len_min], vec[:len_min, :], ppg[:len_min, :]
    hop = int(hp.data.hop_length)
    with torch.no_grad():
        spk = spk.unsqueeze(0).to(device)
        source = model.pitch2source(pit.unsqueeze(0).to(device), rand_ini=rand_ini, noise=noise)
        if write_pit_wav:
            from scipy.io.wavfile import write
            write(write_pit_wav, hp.data.sampling_rate, model.source2wav(source))
        plan = chunk_plan(len_min, hop)
        outs: List[np.ndarray | None] = [None] * len(plan)
        # group chunk indices by length; equal-length chunks become one batch
        by_len = {}
        for i, (cs, ce, _, _) in enumerate(plan):
            by_len.setdefault(ce - cs, []).append(i)
        for n, idxs in by_len.items():
            for g0 in range(0, len(idxs), max_batch):
                grp = idxs[g0:g0 + max_batch]
                B = len(grp)
                sub_ppg = torch.stack([ppg[plan[i][0]:plan[i][1]] for i in grp]).to(device)
                sub_vec = torch.stack([vec[plan[i][0]:plan[i][1]] for i in grp]).to(device)
                sub_pit = torch.stack([pit[plan[i][0]:plan[i][1]] for i in grp]).to(device)
                sub_len = torch.LongTensor([n] * B).to(device)
                sub_har = torch.cat([source[:, :, plan[i][0] * hop:plan[i][1] * hop] for i in grp], 0)
                eps = None
                if eps_fn is not None:
                    eps = torch.cat([eps_fn(i, 1, n) for i in grp], 0)
                out = model.inference(sub_ppg, sub_vec, sub_pit, spk.expand(B, -1), sub_len, sub_har, eps=eps)
                out = out[:, 0].cpu().numpy()
                for b, i in enumerate(grp):
                    outs[i] = out[b][plan[i][2]:plan[i][3]]
        return np.concatenate(outs).astype(np.float32)
