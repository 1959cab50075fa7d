"""Representative hot-path launches for `ncu --set full` (two launches each; profile with
-k regex:'tc_gemm|conv3x3|stem_pool|attention_tc|preprocess_identity|fuse_compound|w2v_conv0|layernorm|jpeg_')."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import nets, ops, _lib, synthetic as syn, get_weights_matrices as gwm

_lib.require_device()
dev = "cuda"
torch.manual_seed(0)
bf = torch.bfloat16
B = 256


def conv(n, h, w, cin, cout, k, stride=1, res=False, reps=2):
    x = torch.randn(n, h, w, cin, device=dev).to(bf)
    wt = (torch.randn(cout, k * k * cin, device=dev) / (cin * k * k) ** 0.5).to(bf)
    b = torch.randn(cout, device=dev)
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    r = torch.randn(n, ho, wo, cout, device=dev).to(bf) if res else None
    for _ in range(reps):
        ops.conv2d_nhwc(x, wt, b, kh=k, kw=k, stride=stride, pad_h=(k - 1) // 2, pad_w=(k - 1) // 2, residual=r, act=ops.ACT_RELU)
    torch.cuda.synchronize()


# VS layers at batch 256 (two launches each)
net = nets.VSNet(syn.make_vs_state_dict(0, "default"), "bf16", dev)
crops = torch.randint(0, 256, (1024, 224, 224, 3), dtype=torch.uint8, device=dev)
xin = net.alloc_input(1024)
for _ in range(2):
    ops.preprocess(crops, 1024, xin, 1)                                  # K1 at 1024 crops
for _ in range(2):
    ops.stem_pool(xin[:B], net.w["stem_packed"], net.w["stem"].bias)     # fused stem + pool
for _ in range(2):
    ops.stem_pool_u8(crops[:B], net.w["stem_packed"], net.w["stem"].bias)   # K1 fused into the stem (measured slower: why?)
conv(B, 55, 55, 64, 64, 3)                   # l1.c2  halo 3x3, resident weights
conv(B, 28, 28, 128, 128, 3)                 # l2.c2  halo 3x3, streamed weights
conv(B, 14, 14, 256, 256, 3)                 # l3.c2  two-SM implicit GEMM (MMA bound)
conv(B, 55, 55, 64, 256, 1, res=True)        # l1.c3  flat epilogue + residual (HBM bound)
conv(B, 14, 14, 256, 1024, 1, res=True)      # l3.c3  flat epilogue + residual
conv(B, 55, 55, 256, 64, 1)                  # l1.c1  64-wide tiles (HBM bound)
xc = torch.randn(B * 55 * 55, 128, device=dev).to(bf); wc = (torch.randn(256, 128, device=dev) / 11).to(bf); bc = torch.zeros(256, device=dev)
for _ in range(2):
    ops.linear(xc, wc, bc, act=ops.ACT_RELU)  # l1.0 conv3 + projection shortcut as one K = 128 GEMM (no residual read)
# audio: FFN GEMM with GELU and attention, 64 windows
x = torch.randn(64 * 199, 1024, device=dev).to(bf); w = (torch.randn(4096, 1024, device=dev) / 32).to(bf); b = torch.zeros(4096, device=dev)
for _ in range(2):
    ops.linear(x, w, b, act=ops.ACT_GELU)
f4 = torch.randn(64 * 199, 4096, device=dev).to(bf); w2 = (torch.randn(1024, 4096, device=dev) / 64).to(bf); b2 = torch.zeros(1024, device=dev)
for _ in range(2):
    ops.linear(f4, w2, b2, residual=x)        # FFN-out GEMM: barrier epilogue with the 4-deep residual ring
qkv = torch.randn(64 * 199, 3072, device=dev).to(bf)
for _ in range(2):
    ops.attention(qkv, 64, 199, 16, 64, 0.125)
# K4 at 1.5 M frames
n = 1_500_000
ps = [torch.softmax(torch.randn(n, 7, device=dev), 1).contiguous() for _ in range(3)]
for _ in range(2):
    ops.fuse_compound(ps[0], ps[1], ps[2], gwm.class_weights(gwm.weights_3), [1, 1, 1], False, True)
# baseline-JPEG decode of 1500 crops 224x224 (un-stuff, Huffman, IDCT, colour)
import cv2
from avcer_b200 import jpeg
files = [cv2.imencode(".jpg", c)[1].tobytes() for c in syn.make_crops(3, 30)] * 50
for _ in range(2):
    jpeg.decode_batch(files, dev)
torch.cuda.synchronize()
print("prof_kernels done")
