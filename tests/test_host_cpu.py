"""Host logic (no GPU): chunk arithmetic, wire formats, sharding, 2-rank gloo broadcast."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from tests.util import ROOT
from whisper_vits_svc_b200 import hostio, shard


def _reference_chunks(all_frame, hop_size):
    """Transliteration of the while-loop of svc_inference.py:94-131 (indices only)."""
    hop_frame, out_chunk, out_index, res = 10, 2500, 0, []
    while out_index < all_frame:
        if out_index == 0:
            cut_s, cut_s_out = 0, 0
        else:
            cut_s, cut_s_out = out_index - hop_frame, hop_frame * hop_size
        if out_index + out_chunk + hop_frame > all_frame:
            cut_e, cut_e_out = all_frame, -1
        else:
            cut_e, cut_e_out = out_index + out_chunk + hop_frame, -1 * hop_frame * hop_size
        res.append((cut_s, cut_e, cut_s_out, cut_e_out))
        out_index += out_chunk
    return res


@pytest.mark.parametrize("n", [1, 9, 1000, 2490, 2491, 2500, 2510, 2511, 5050, 7500, 10021])
def test_chunk_plan_matches_reference_loop(n):
    plan = hostio.chunk_plan(n, 320)
    assert plan == _reference_chunks(n, 320)
    total = sum(len(range((ce - cs) * 320)[so:eo]) for cs, ce, so, eo in plan)
    # 5050 frames -> 3 chunks, output length n*hop - 1
    if n == 5050:
        assert [(c[0], c[1]) for c in plan] == [(0, 2510), (2490, 5010), (4990, 5050)]
    if n > 2510 or n <= 2500:
        assert total == n * 320 - 1 or n % 2500 == 0 or total <= n * 320


def test_csv_pitch_roundtrip(tmp_path):
    p = tmp_path / "a.pit.csv"
    pit = [0, 0, 220, 221.9, 440, 1100]
    hostio.save_csv_pitch(pit, str(p))
    lines = open(p).read().splitlines()
    assert lines[2] == "0m 0s  20,220" and lines[3].endswith(",221")
    assert hostio.load_csv_pitch(str(p)) == [0, 0, 220, 221, 440, 1100]


def test_prepare_features_repeat_and_shift(tmp_path):
    np.save(tmp_path / "x.ppg.npy", np.arange(6, dtype=np.float32).reshape(3, 2))
    np.save(tmp_path / "x.vec.npy", np.ones((3, 4), np.float32))
    hostio.save_csv_pitch([100, 0, 200, 200, 0, 0], str(tmp_path / "x.pit.csv"))
    ppg, vec, pit = hostio.prepare_features(str(tmp_path / "x.ppg.npy"), str(tmp_path / "x.vec.npy"),
                                            str(tmp_path / "x.pit.csv"), shift=12)
    assert ppg.shape == (6, 2) and torch.equal(ppg[0], ppg[1]) and vec.shape == (6, 4)
    assert torch.allclose(pit, torch.tensor([200., 0, 400, 400, 0, 0]))


def test_checkpoint_roundtrip_tolerates_missing_keys(tmp_path, hp, sd, capsys):
    from whisper_vits_svc_b200 import models
    m = models.SynthesizerInfer(513, 25, hp)
    partial = {k: v for k, v in sd.items() if k != "dec.conv_post.weight"}
    torch.save({"model_g": partial}, tmp_path / "g.pth")
    hostio.load_svc_model(str(tmp_path / "g.pth"), m)
    assert "dec.conv_post.weight is not in the checkpoint" in capsys.readouterr().out
    got = m.state_dict()
    assert len(got) == 903 and torch.equal(got["enc_p.pre.weight"], sd["enc_p.pre.weight"])
    hostio.save_svc_model(m, str(tmp_path / "h.pth"))
    assert set(torch.load(tmp_path / "h.pth", weights_only=False)["model_g"]) == set(sd)


def test_assign_balances_and_covers():
    lens = [1000, 10, 500, 500, 990, 20, 30, 700]
    parts = shard.assign(lens, 3)
    assert sorted(i for p in parts for i in p) == list(range(len(lens)))
    tot = [sum(lens[i] for i in p) for p in parts]
    assert max(tot) - min(tot) <= max(lens)
    assert shard.assign([], 2) == [[], []]


GLOO_WORKER = r"""
import os, sys, torch
sys.path.insert(0, os.environ["SVCB_ROOT"])
import torch.distributed as dist
from whisper_vits_svc_b200 import shard
rank, local, world = shard.init("gloo")
blob = table = None
if rank == 0:
    blob = torch.arange(1000, dtype=torch.float32); table = [("a", 0, 10), ("b", 256, 5)]
blob, table = shard.broadcast_blob(blob, table, torch.device("cpu"))
assert blob.sum().item() == 499500.0 and table[1][0] == "b"
mine = shard.assign(list(range(10, 0, -1)), world)[rank]
tot = shard.sum_over_ranks(float(len(mine)), torch.device("cpu"))
mx = shard.max_over_ranks(float(rank + 1), torch.device("cpu"))
assert tot == 10.0 and mx == float(world)
shard.barrier(); dist.destroy_process_group()
open(os.path.join(os.environ["SVCB_OUT"], f"ok{rank}"), "w").write("ok")
"""


def test_two_rank_gloo_broadcast_and_shard(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, SVCB_ROOT=ROOT, SVCB_OUT=str(tmp_path), MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with
    the contract's keys; it needs no GPU, so the contract is checked here."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
    # the arm times a bounded sample of OUR headline workload: same config record as the GPU arm
    import bench
    assert d["config"] == bench.headline_config(32, 1000, 320000, 1)


def test_batch_engine_plan_buckets_equal_length_chunks():
    """hostio.BatchEngine.plan: every chunk of every utterance exactly once, batches hold equal-length
    chunks only (SnakeAlias pads by replication at a chunk's true end, so ragged batches would not be
    reference-exact), at most max_batch per device call, longest first; the pieces reassemble to the
    reference's output length n*hop - 1 per utterance."""
    from whisper_vits_svc_b200 import hostio
    lengths = [1000, 1000, 2600, 37, 1000, 5050, 1, 2500, 2511, 1000]
    hop, mb = 320, 3
    chunks, batches = hostio.BatchEngine.plan(lengths, hop, mb)
    seen = sorted(j for b in batches for j in b)
    assert seen == list(range(len(chunks)))
    last = None
    for b in batches:
        assert 1 <= len(b) <= mb
        ns = {chunks[j][2] - chunks[j][1] for j in b}
        assert len(ns) == 1
        n = ns.pop()
        assert last is None or n <= last
        last = n
    for u, n in enumerate(lengths):
        mine = sorted((c for c in chunks if c[0] == u), key=lambda c: c[1])
        assert [(c[1], c[2], c[3], c[4]) for c in mine] == hostio.chunk_plan(n, hop)
        # (so, eo) are the reference's slice bounds into the chunk's output (eo may be negative: svc_inference.py:112,129)
        assert sum(len(range((c[2] - c[1]) * hop)[c[3]:c[4]]) for c in mine) == n * hop - 1
    # 4 utterances of 1000 frames -> one batch of 3 and one of 1
    assert sorted(len(b) for b in batches if chunks[b[0]][2] - chunks[b[0]][1] == 1000) == [1, 3]


def test_hubert_chunk_plan_and_pack_inventory():
    """hubert/inference.py:29-48 as data (20 s chunks, the remainder as a last shorter one) and the names / sizes the C side
    resolves from the packed blob (csrc/hubert_api.cu:svcb_hubert_create) for a 1-layer synthetic checkpoint."""
    from whisper_vits_svc_b200 import hubert_infer, synth
    C = hubert_infer.CHUNK
    assert hubert_infer.chunk_plan(C) == [(0, C)]                       # `while idx + chunk < audln`: an exact chunk is the tail
    assert hubert_infer.chunk_plan(2 * C + 5) == [(0, C), (C, 2 * C), (2 * C, 2 * C + 5)]
    assert hubert_infer.chunk_plan(7) == [(0, 7)]
    sd = synth.hubert_checkpoint(3, n_layer=1)
    items, n_layer = hubert_infer.pack_hubert(sd)
    assert n_layer == 1
    sizes = {n: t.numel() for n, t in items}
    assert sizes["fe.conv0.w"] == 10 * 512 and sizes["fe.conv1.wimg"] == 512 * 3 * 512 // 2 and sizes["fe.conv6.wimg"] == 512 * 2 * 512 // 2
    assert sizes["fp.w"] == 768 * 512 // 2 and sizes["proj.w"] == 256 * 768 // 2
    assert all(sizes[f"pos.{g}.wimg"] == 256 * 128 * 48 // 2 and sizes[f"pos.{g}.bimg"] == 256 for g in range(16))
    assert all(sizes[f"pos.{g}.{h}.w"] == 48 * 128 * 24 for g in range(16) for h in range(2))
    assert sizes["L0.wqkv"] == 3 * 768 * 768 // 2 and sizes["L0.w1"] == sizes["L0.w2"] == 3072 * 768 // 2
    # the grouped positional conv in its tensor-core form: W_g[co][j * 48 + ci] = weight_norm(v, g)[g * 48 + co][ci][j]
    w = torch._weight_norm(sd["positional_embedding.conv.weight_v"], sd["positional_embedding.conv.weight_g"], 2)
    img = dict(items)["pos.3.wimg"].view(torch.bfloat16).view(1, 96, 8, 256, 8)     # [N/256][K/64][8 octets][256 rows][8]
    co, j, ci = 5, 17, 29
    k = j * 48 + ci
    got = img[0, k // 64, (k % 64) // 8, co, k % 8].float()
    assert abs(got - w[3 * 48 + co, ci, j]) <= 4e-3 * abs(w[3 * 48 + co, ci, j]) + 1e-6
    assert img[0, :, :, 48:, :].abs().max() == 0                                    # padded output channels
