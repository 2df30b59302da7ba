"""Wait-cycle counters of one CTA of the Whisper attention kernel (debug; run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import importlib
_lib = importlib.import_module("whisper_vits_svc_b200._lib")
lib = _lib.load()
B, T, heads = 16, 1500, 20
D = heads * 64
qkv = torch.randn(B, T, 3 * D, device="cuda").bfloat16()
out = torch.zeros(B, T, D, device="cuda", dtype=torch.bfloat16)
scratch = torch.empty(int(lib.svcb_op_attention_tc_bf16_scratch_bytes(B, T, D)), dtype=torch.uint8, device="cuda")
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    lib.svcb_debug_s2d_trace(buf.data_ptr() if rep == 2 else None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.svcb_op_attention_tc_bf16(qkv.data_ptr(), out.data_ptr(), B, T, D, heads, 0, scratch.data_ptr(), scratch.numel(), s), "attn")
    e1.record(); torch.cuda.synchronize()
    print("op ms (with image conversions)", e0.elapsed_time(e1))
lib.svcb_debug_s2d_trace(None)
t = buf.cpu().tolist()
print("softmax: wait s_full %d | busy %d | rescale rounds %d | total %d" % (t[0], t[1], t[3], t[6]))
print("mma: wait K %d p_full %d V %d | total %d" % (t[8], t[9], t[10], t[13]))
