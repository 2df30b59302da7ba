"""Stage-sized launches of the space-to-depth AMP link for ncu (c2-type: residual + fp32 out + image out)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from whisper_vits_svc_b200 import ops, synth, hparams

hp = hparams.load_hparams(os.path.join(os.path.dirname(__file__), "..", "configs", "base.yaml"))
sd = synth.svc_state_dict(hp, 1234)
fu, fd = sd["dec.activation_post.upsample.filter"], sd["dec.activation_post.downsample.lowpass.filter"]
g = torch.Generator().manual_seed(0)
for C, L, K, dil in ((20, 160000, 3, 1), (40, 80000, 11, 5), (10, 320000, 7, 1)):
    x = torch.randn(32, C, L, generator=g).cuda()
    w = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
    b = torch.randn(C, generator=g).cuda() * 0.1
    a1, b1, a2, b2 = [torch.randn(C, generator=g) * 0.4 for _ in range(4)]
    for _ in range(2):
        y, ya = ops.amp_s2d_link(x, a1, b1, fu, fd, w, b, dilation=dil, res=x, alpha_out=a2, beta_out=b2)
    torch.cuda.synchronize()
print("done")
