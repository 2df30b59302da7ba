#!/usr/bin/env python
"""bench.py — SVC inference throughput (audio samples/s) on N B200s of one node.

Workload (BASELINE.json configs[3], the configuration the 1->8 GPU metric is quoted on): full
SynthesizerInfer — NSF source + prior encoder + reverse flow + NSF-BigVGAN generator — on a batch
of 32 synthetic 10 s utterances (T=1000 frames -> 320,000 samples each) PER GPU (weak scaling:
utterances shard across ranks, no data-path collective).  A step = one pass of that hot path
over one batch.  Seeded synthetic weights in the reference checkpoint format (no pretrained
weights exist offline), synthetic inputs per SURVEY.md §8d.

  value : whole-job samples/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public SynthesizerInfer API with HOST pinned inputs
          (H2D of ppg/vec/pit/spk/lengths and D2H of the waveform inside the timed region)
  roofline / cpu_baseline : see DESIGN.md §Measurement

`--impl reference` times the reference's own CPU algorithm (the oracle restatement, which was
checked bit-exact against the imported reference; the Python reference itself cannot travel to
the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "audio samples/sec (SVC infer: F0+PPG+vec+spk -> waveform)"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU per step")
    ap.add_argument("--frames", type=int, default=1000, help="frames per utterance (100 fps)")
    ap.add_argument("--precision", type=int, default=3, help="3 bf16x3 tensor-core (parity grade, default), 1 bf16, 0 fp32 CUDA cores")
    ap.add_argument("--workload", default="svc", choices=["svc", "whisper"],
                    help="svc = BASELINE configs[3] (headline); whisper = configs[2] PPG extraction, 16 x 30 s log-mel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sust=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.p = gpu_index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_threads() -> int:
    """Threads for the CPU arm.  torch's intra-op pool over-subscribes badly on the many tiny
    depthwise convolutions of this path: on the 128-core GPU-box host 128 threads ran 60x slower
    than 8 (gpurun_out of round 1), so the arm uses the count that is fastest in practice, capped
    by what the box has; SVCB_CPU_THREADS overrides."""
    env = os.environ.get("SVCB_CPU_THREADS")
    if env:
        return max(1, int(env))
    return max(1, min(os.cpu_count() or 1, 16))


def synth_inputs(hp, B, T, seed):
    from tests.util import make_inputs
    return make_inputs(seed, B, T, hp)


# ------------------------------------------------------------------------------ reference arm
def cpu_reference_run(hp, sd, T, repeats):
    """The reference's CPU algorithm (oracle port) on one utterance of T frames; returns best s."""
    from oracle import svc_oracle as O
    d = synth_inputs(hp, 1, T, 4242)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            src = O.pitch2source(sd, hp, d["pit"], d["rand_ini"], d["noise"])
            O.synthesizer_infer(sd, hp, d["ppg"], d["vec"], d["pit"], d["spk"], d["ppg_l"], src, d["eps"])
        best = min(best, time.perf_counter() - t0)
    return best


def run_reference(args, hp, sd):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = cpu_threads()
    torch.set_num_threads(cores)
    T = 250  # bounded sample: one 2.5 s utterance per step (CPU throughput is length-independent)
    hop = int(np.prod(list(hp.gen.upsample_rates)))
    for _ in range(max(args.warmup, 1)):
        cpu_reference_run(hp, sd, T, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_run(hp, sd, T, 1)
    dt = time.perf_counter() - t0
    val = args.steps * T * hop / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded weights + inputs)",
        "config": {"workload": "SynthesizerInfer (configs[3]) on the host CPU, 1 x 2.5 s utterance per step",
                   "frames": T, "batch": 1},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x one {T / 100:.1f} s utterance, torch CPU fp32, {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------ our arm
def run_ours(args, hp, sd):
    from whisper_vits_svc_b200 import _lib, models, shard

    rank, local, world = shard.init()
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs CUDA devices; there is no CPU path"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    B, T = args.batch, args.frames
    hop = int(np.prod(list(hp.gen.upsample_rates)))
    L = T * hop

    m = models.SynthesizerInfer(hp.data.filter_length // 2 + 1, hp.data.segment_size // hp.data.hop_length, hp,
                                precision=args.precision)
    if world == 1:
        m.load_state_dict(sd)
        m.to(dev)
        m._ensure()
    else:  # rank 0 packs, one NCCL broadcast of the packed blob
        blob = table = None
        if rank == 0:
            m.load_state_dict(sd)
            m.to(dev)
            blob, table = m.packed_blob()
        else:
            m.to(dev)
        blob, table = shard.broadcast_blob(blob, table, dev)
        if rank != 0:
            m.install_blob(blob, table)

    d = synth_inputs(hp, B, T, 1000 + rank)
    dv = {k: v.to(dev) for k, v in d.items()}
    host = {k: d[k].pin_memory() for k in ("ppg", "vec", "pit", "spk", "ppg_l")}
    wave_host = torch.empty(B, 1, L, dtype=torch.float32).pin_memory()
    h2d = sum(host[k].numel() * host[k].element_size() for k in host)
    d2h = wave_host.numel() * 4
    launches = [0]

    def step_device():
        src = m.pitch2source(dv["pit"], rand_ini=dv["rand_ini"], noise=dv["noise"])
        launches[0] += _lib.last_launch_count()
        w = m.inference(dv["ppg"], dv["vec"], dv["pit"], dv["spk"], dv["ppg_l"], src, eps=dv["eps"])
        launches[0] += _lib.last_launch_count()
        return w

    def step_e2e():
        x = {k: host[k].to(dev, non_blocking=True) for k in host}
        src = m.pitch2source(x["pit"])  # device-side RNG draws, as the reference does on its device
        w = m.inference(x["ppg"], x["vec"], x["pit"], x["spk"], x["ppg_l"], src)
        wave_host.copy_(w, non_blocking=True)
        return w

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        shard.barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches[0] = 0
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        shard.barrier()
        clocks = sampler.stop() if sampler else None
        ms = shard.max_over_ranks(float(e0.elapsed_time(e1)), dev)
        return ms, clocks

    sampler = ClockSampler(local) if rank == 0 else None
    ms, clocks = timed(step_device, args.steps, args.warmup, sampler)
    n_launch = launches[0]
    total_samples = float(world * B * L * args.steps)
    value = total_samples / (ms * 1e-3)
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, min(args.warmup, 2)))
    e2e_value = total_samples / (ms_e2e * 1e-3)

    roof, kernels = None, None
    if rank == 0 and not args.no_roofline:
        peaks = load_peaks()
        lib.svcb_timing_enable(1)
        for _ in range(args.steps):
            step_device()
        torch.cuda.synchronize(dev)
        rep = lib.svcb_timing_report().decode()
        lib.svcb_timing_enable(0)
        import re
        if os.environ.get("SVCB_DUMP_KERNELS"):
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "kernels_raw.txt"), "w") as f:
                f.write(rep)
        fam = {}
        for line in rep.strip().splitlines():
            nm, n, tms, fl, by = line.split()
            # launches of one kernel with different shapes are tagged _c<ch>k<taps> / _<cin>to<cout>_...: one family
            f = re.sub(r"(_c\d+(k\d+)?|_\d+to\d+_k\d+_o\d+)$", "", nm)
            a = fam.setdefault(f, dict(name=f, launches=0, ms=0.0, flops=0.0, bytes=0.0))
            a["launches"] += int(n); a["ms"] += float(tms); a["flops"] += float(fl); a["bytes"] += float(by)
        rows = sorted(fam.values(), key=lambda r: -r["ms"])
        tot = sum(r["ms"] for r in rows) or 1.0
        kernels = [dict(name=r["name"], launches=r["launches"], share=round(r["ms"] / tot, 4),
                        ms_per_step=round(r["ms"] / args.steps, 3),
                        tflops=round(r["flops"] / (r["ms"] * 1e-3) / 1e12, 2) if r["ms"] else 0.0,
                        gbs=round(r["bytes"] / (r["ms"] * 1e-3) / 1e9, 1) if r["ms"] else 0.0) for r in rows]
        top = rows[0]
        # which roof bounds the dominant kernel: its arithmetic intensity against the ridge of the
        # measured peaks (tensor peak / HBM peak); CUDA-core kernels above the ridge are reported
        # against the tensor roof too (the honest, unflattering denominator) with the fp32-FMA
        # fraction given beside it
        ridge = peaks["tf_sust"] * 1e12 / (peaks["hbm"] * 1e9)
        intensity = top["flops"] / max(top["bytes"], 1.0)
        fp32_peak_tf = 148 * 128 * 2 * 1.965e9 / 1e12  # 148 SMs x 128 FMA lanes x 2 x max clock
        if intensity > 0.1 * ridge:
            ach = top["flops"] / (top["ms"] * 1e-3) / 1e12
            peak = peaks["tf_sust"]
            roof = {"kernel": top["name"], "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak, "traffic": None}
            if "tc" not in top["name"] and "gemm" not in top["name"]:
                roof["note"] = ("this kernel runs on the fp32 FMA pipe (no tensor-core instructions); "
                                f"fraction of the {fp32_peak_tf:.1f} TFLOP/s fp32 peak = {ach / fp32_peak_tf:.3f}")
        else:
            ach = top["bytes"] / (top["ms"] * 1e-3) / 1e9
            peak = peaks["hbm"]
            roof = {"kernel": top["name"], "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": None}
        try:  # measured DRAM traffic of this family (ncu --set full capture committed under profiles/)
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as tf:
                tr = json.load(tf).get(top["name"])
            if tr:
                roof["traffic"] = tr["bytes_per_launch"]
                roof["traffic_unit"] = "bytes per launch (dram read+write, " + tr["source"] + ")"
                roof["algorithmic_bytes_per_launch"] = round(top["bytes"] / max(top["launches"], 1))
        except (OSError, ValueError, KeyError):
            pass
        roof["intensity_flop_per_byte"] = round(intensity, 1)
        roof["peak_source"] = f"of {peaks['src']} (MEASURED_PEAKS.json sustained bf16 / copy bandwidth)"
        roof["how"] = (f"CUDA events around every launch of this kernel over {args.steps} steps identical to the "
                       "timed region (svcb_timing_enable); achieved = summed algorithmic FLOPs (or bytes) / summed duration")
        roof["launches_per_step"] = top["launches"] // args.steps
        roof["share_of_step"] = round(top["ms"] / tot, 4)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = cpu_threads()
        torch.set_num_threads(cores)
        Tc = 250
        cpu_reference_run(hp, sd, 50, 1)  # warm-up of the oneDNN primitives
        sec = cpu_reference_run(hp, sd, Tc, 2)
        cpu = {"value": Tc * hop / sec, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"best of 2 x one {Tc / 100:.1f} s utterance (same per-item workload), oracle port of the "
                         f"reference CPU path, torch fp32, {cores} threads, {sec:.2f} s"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f32", 1: "bf16", 3: "bf16x3 (bf16 operands split 3-way, fp32 accumulate) + f32"}[args.precision],
            "data": "synthetic (seeded weights in the reference checkpoint format + seeded inputs)",
            "config": {"workload": "BASELINE configs[3]: full SynthesizerInfer (F0->NSF source, prior, flow, generator), "
                                   f"{B} x {T / 100:.0f} s utterances per GPU per step, 32 kHz/hop 320",
                       "batch_per_gpu": B, "frames": T, "samples_per_item": L, "parallelism": f"utterance-shard x{world}",
                       "cache": "inputs (ppg+vec+noise+eps ~ %.0f MB) exceed the 126 MB L2" % (
                           (dv["ppg"].numel() + dv["vec"].numel() + dv["noise"].numel() + dv["eps"].numel()) * 4 / 1e6),
                       "rtf_32k": (ms * 1e-3 / args.steps) / (world * B * L / 32000.0),
                       "rtf_24k_label": (ms * 1e-3 / args.steps) / (world * B * L / 24000.0)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": n_launch,
            "clocks": clocks,
        }
        if roof:
            out["roofline"] = roof
            out["kernels"] = kernels
        if cpu:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def run_whisper(args):
    """BASELINE configs[2]: truncated Whisper large-v2 encoder (24 blocks), 16 x 30 s log-mel per step."""
    from whisper_vits_svc_b200 import _lib, synth, whisper_infer
    assert torch.cuda.is_available()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    lib = _lib.load()
    ck = synth.whisper_checkpoint(seed=1234)
    enc = whisper_infer.WhisperB200(ck, dev).encoder
    B, n = 16, 3000
    g = torch.Generator().manual_seed(0)
    mel = torch.randn(B, 80, n, generator=g).clamp(-1, 1.5)
    mel_d = mel.to(dev)
    mel_h = mel.pin_memory()
    out_h = torch.empty(B, 1500, 1280).pin_memory()
    for _ in range(args.warmup):
        enc(mel_d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        enc(mel_d)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    e0.record()
    for _ in range(args.steps):
        out_h.copy_(enc(mel_h.to(dev, non_blocking=True)), non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    # front end on the device: 16 x 30 s of 16 kHz audio -> log-mel (+ the extractor's
    # noise term) -> encoder; `from_audio` = pinned host audio in, PPG back on the host
    audio = (torch.randn(B, n * 160, generator=g) * 0.1)
    noise_d = torch.randn(B, 80, n, generator=g).to(dev)
    audio_d, audio_h = audio.to(dev), audio.pin_memory()
    for _ in range(2):
        enc.log_mel(audio_d, noise_d, 0.1)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        enc.log_mel(audio_d, noise_d, 0.1)
    e1.record(); torch.cuda.synchronize()
    ms_fe = e0.elapsed_time(e1) / args.steps
    e0.record()
    for _ in range(args.steps):
        out_h.copy_(enc(enc.log_mel(audio_h.to(dev, non_blocking=True), noise_d, 0.1)), non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms_audio = e0.elapsed_time(e1) / args.steps
    lib.svcb_timing_enable(1)
    for _ in range(args.steps):
        enc(enc.log_mel(audio_d, noise_d, 0.1))
    torch.cuda.synchronize()
    rep = lib.svcb_timing_report().decode(); lib.svcb_timing_enable(0)
    peaks = load_peaks()
    ks = []
    for line in rep.strip().splitlines():
        nm, nl, tms, fl, by = line.split()
        ks.append(dict(name=nm, launches=int(nl), ms_per_step=round(float(tms) / args.steps, 3),
                       tflops=round(float(fl) / (float(tms) * 1e-3) / 1e12, 1), gbs=round(float(by) / (float(tms) * 1e-3) / 1e9, 1)))
    ks.sort(key=lambda k: -k["ms_per_step"])
    flops = 1708.6e9 * B
    audio_s = 30.0 * B
    top = ks[0]
    out = {"metric": "audio seconds/sec (Whisper-large-v2 truncated encoder, PPG extraction)", "value": audio_s / (ms * 1e-3),
           "unit": "audio s/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": "BASELINE configs[2]: 16 x 30 s log-mel [16,80,3000] -> PPG [16,1500,1280], 24 blocks",
                      "tflops_model": flops / (ms * 1e-3) / 1e12},
           "e2e": {"value": audio_s / (ms_e2e * 1e-3), "unit": "audio s/s", "h2d_bytes_per_step": mel.numel() * 4,
                   "d2h_bytes_per_step": out_h.numel() * 4},
           "frontend": {"what": "svcb_whisper_log_mel: [16, 480000] audio -> [16, 80, 3000] log-mel + noise, on the device",
                        "ms_per_step": ms_fe, "audio_s_per_s": audio_s / (ms_fe * 1e-3)},
           "from_audio": {"what": "pinned host audio -> log-mel -> encoder -> PPG on the host", "ms_per_step": ms_audio,
                          "value": audio_s / (ms_audio * 1e-3), "unit": "audio s/s",
                          "h2d_bytes_per_step": audio.numel() * 4, "d2h_bytes_per_step": out_h.numel() * 4},
           "roofline": {"kernel": top["name"], "bound": "tensor", "achieved": top["tflops"], "peak": peaks["tf_sust"],
                        "unit": "TFLOP/s", "frac": top["tflops"] / peaks["tf_sust"], "traffic": None},
           "kernels": ks, "gpu_launches": int(sum(k["launches"] for k in ks))}
    print(json.dumps(out), flush=True)


def main():
    args = parse()
    if args.workload == "whisper":
        return run_whisper(args)
    from whisper_vits_svc_b200 import hparams, synth
    hp = hparams.load_hparams(os.path.join(ROOT, "configs", "base.yaml"))
    sd = synth.svc_state_dict(hp, 1234)
    if args.impl == "reference":
        run_reference(args, hp, sd)
    else:
        run_ours(args, hp, sd)


if __name__ == "__main__":
    main()
