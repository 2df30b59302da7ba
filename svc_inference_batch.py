#!/usr/bin/env python
"""Drop-in for the reference's svc_inference_batch.py (same flags; outputs ./_svc_out/<name>.wav).

The reference loads Whisper once, writes a .ppg.npy per file and then starts one
`python svc_inference.py` subprocess per file (svc_inference_batch.py:34-43), re-loading the SVC
checkpoint and re-spawning the HuBERT / pitch extractors every time.  Here the model is loaded
once per GPU, files are sharded across the ranks of a `torchrun` launch (one process per GPU, one
NCCL broadcast of the packed weights, no collective on the data path) and every file runs
in-process.  HuBERT / CREPE are outside the B200 hot path: `<name>.vec.npy` and
`<name>.pit.csv` must sit next to `<name>.wav` (or in --feat); a file without them is reported and
skipped — the reference would silently produce nothing for it either (subprocess exit codes are
ignored there).

    python svc_inference_batch.py --config configs/base.yaml --model sovits5.0.pth --wave waves/ --spk singer.npy
    python -m torch.distributed.run --nproc-per-node 8 svc_inference_batch.py ...   # 8 GPUs
"""
import argparse
import os
import sys

sys.path.append(os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
from scipy.io.wavfile import write

from whisper_vits_svc_b200 import hostio, hparams, models, shard

out_path = "./_svc_out"


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--config', type=str, required=True, help="yaml file for config.")
    parser.add_argument('--model', type=str, required=True, help="path of model for evaluation")
    parser.add_argument('--wave', type=str, required=True, help="Path of raw audio.")
    parser.add_argument('--spk', type=str, required=True, help="Path of speaker.")
    parser.add_argument('--shift', type=int, default=0, help="Pitch shift key.")
    parser.add_argument('--feat', type=str, default=None, help="dir holding <name>.{ppg,vec}.npy / .pit.csv")
    parser.add_argument('--whisper', type=str, default=os.path.join("whisper_pretrain", "large-v2.pt"))
    args = parser.parse_args()
    wave_path = args.wave
    assert os.path.isdir(wave_path), f"{wave_path} is not folder"
    os.makedirs(out_path, exist_ok=True)
    waves = sorted(f for f in os.listdir(wave_path) if f.endswith(".wav"))
    if not torch.cuda.is_available():
        raise SystemExit("this build has no CPU path: a CUDA (sm_100a) device is required")
    rank, local, world = shard.init()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    hp = hparams.load_hparams(args.config)
    model = models.SynthesizerInfer(hp.data.filter_length // 2 + 1, hp.data.segment_size // hp.data.hop_length, hp,
                                    precision=3)
    blob = table = None
    if rank == 0:
        hostio.load_svc_model(args.model, model)
        model.to(device)
        blob, table = model.packed_blob()
    else:
        model.to(device)
    if world > 1:
        blob, table = shard.broadcast_blob(blob, table, device)
        if rank != 0:
            model.install_blob(blob, table)
    sizes = [os.path.getsize(os.path.join(wave_path, f)) for f in waves]
    mine = shard.assign(sizes, world)[rank]
    feat = args.feat or wave_path
    spk = torch.FloatTensor(np.load(args.spk))
    whisper = None
    done, failed = 0, []
    for i in mine:
        name = waves[i]
        stem = os.path.join(feat, name[:-4])
        try:
            ppg_p, vec_p, pit_p = stem + ".ppg.npy", stem + ".vec.npy", stem + ".pit.csv"
            if not (os.path.isfile(vec_p) and os.path.isfile(pit_p)):
                raise FileNotFoundError(f"{vec_p} / {pit_p} missing (HuBERT and CREPE extractors are out of scope)")
            if not os.path.isfile(ppg_p):
                from whisper_vits_svc_b200 import whisper_infer
                if whisper is None:
                    whisper = whisper_infer.load_model(args.whisper, device)
                ppg_p = os.path.join(out_path, name + ".ppg.npy")
                whisper_infer.pred_ppg(whisper, os.path.join(wave_path, name), ppg_p, device)
            ppg, vec, pit = hostio.prepare_features(ppg_p, vec_p, pit_p, args.shift)
            audio = hostio.svc_infer(model, spk, pit, ppg, vec, hp, device, write_pit_wav=None)
            write(os.path.join(out_path, name), hp.data.sampling_rate, audio)
            done += 1
        except Exception as e:  # a failed utterance must not stop the shard
            failed.append((name, repr(e)))
            print(f"[rank {rank}] {name}: FAILED {e}")
    total = shard.sum_over_ranks(float(done), device)
    if rank == 0:
        print(f"svc_inference_batch: {int(total)}/{len(waves)} files converted on {world} GPU(s)")
    if world > 1:
        shard.barrier()
        torch.distributed.destroy_process_group()
    return 0 if not failed else 1


if __name__ == '__main__':
    sys.exit(main())
