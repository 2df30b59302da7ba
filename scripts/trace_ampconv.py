"""Wait-cycle counters of CTA 0 of one amp_conv_tc launch (debug; run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from whisper_vits_svc_b200 import ops, synth, hparams, _lib

hp = hparams.load_hparams(os.path.join(os.path.dirname(__file__), "..", "configs", "base.yaml"))
sd = synth.svc_state_dict(hp, 1234)
fu, fd = sd["dec.activation_post.upsample.filter"], sd["dec.activation_post.downsample.lowpass.filter"]
g = torch.Generator().manual_seed(0)
lib = _lib.load()
for C, L, K, dil in ((80, 20000, 11, 5), (80, 20000, 3, 1), (160, 5000, 11, 5)):
    x = torch.randn(32, C, L, generator=g).cuda()
    w = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
    b = torch.randn(C, generator=g).cuda() * 0.1
    a1, b1 = [torch.randn(C, generator=g) * 0.4 for _ in range(2)]
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    for rep in range(3):
        lib.svcb_debug_s2d_trace(buf.data_ptr() if rep == 2 else None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.amp_conv_tc(x, a1, b1, fu, fd, w, b, dilation=dil, res=x)
        e1.record(); torch.cuda.synchronize()
    lib.svcb_debug_s2d_trace(None)
    t = buf.cpu().tolist()
    tiles = max(t[44], 1)
    print(f"--- C={C} L={L} K={K} dil={dil}: op (snake_pack + conv) {e0.elapsed_time(e1):.3f} ms; CTA 0: {tiles} tiles")
    print(f"  producer: wait a_empty {t[32]} w_empty {t[33]} total {t[34]}")
    print(f"  mma: wait a_full {t[40]} t_empty {t[41]} w_full {t[42]} total {t[43]}  -> per tile: total {t[43] // tiles}, w_full {t[42] // tiles}, t_empty {t[41] // tiles}, a_full {t[40] // tiles}")
    print(f"  epilogue grp0: wait t_full {t[48]} total {t[49]} | grp1: wait t_full {t[52]} total {t[53]}")
