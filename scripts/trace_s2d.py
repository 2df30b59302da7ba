"""Timeline of CTA 0 of one amp_s2d link launch (clock64 stamps, see csrc/amp_s2d.cu S2D_TRACE)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from whisper_vits_svc_b200 import ops, synth, hparams, _lib

hp = hparams.load_hparams(os.path.join(os.path.dirname(__file__), "..", "configs", "base.yaml"))
sd = synth.svc_state_dict(hp, 1234)
fu, fd = sd["dec.activation_post.upsample.filter"], sd["dec.activation_post.downsample.lowpass.filter"]
g = torch.Generator().manual_seed(0)
lib = _lib.load()
names = ["P:wait a_empty", "P:a_empty ok", "P:A issued", "P:w slot0 ok", "M:wait a_full", "M:a_full ok", "M:t_empty ok", "M:w_hi tap0",
         "M:w_hi tap1", "M:tile issued", "E0:wait t_full", "E0:t_full ok", "E0:tile done", "E15:tile done"]
for C, L, K, dil, use_res in ((20, 160000, 3, 1, True), (20, 160000, 3, 1, False), (20, 160000, 11, 5, True), (10, 320000, 7, 1, True)):
    x = torch.randn(32, C, L, generator=g).cuda()
    w = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
    b = torch.randn(C, generator=g).cuda() * 0.1
    a1, b1, a2, b2 = [torch.randn(C, generator=g) * 0.4 for _ in range(4)]
    buf = torch.zeros(32 * 16, dtype=torch.int64, device="cuda")
    for rep in range(2):
        lib.svcb_debug_s2d_trace(buf.data_ptr() if rep == 1 else None)
        ops.amp_s2d_link(x, a1, b1, fu, fd, w, b, dilation=dil, res=x if use_res else None, alpha_out=a2, beta_out=b2)
        torch.cuda.synchronize()
    lib.svcb_debug_s2d_trace(None)
    t = buf.cpu().view(32, 16)
    t0 = int(t[0, 4])
    print(f"--- C={C} K={K} dil={dil} res={use_res}: cycles relative to the MMA warp's first stamp")
    print("tile " + " ".join(f"{n:>15s}" for n in names))
    for i in list(range(0, 8)) + [20, 21]:
        print(f"{i:4d} " + " ".join(f"{int(t[i, k]) - t0:15d}" for k in range(14)))
    d = (t[8:30, 9] - t[7:29, 9]).float()
    print(f"steady-state cycles per tile (MMA warp): mean {float(d.mean()):.0f}")
