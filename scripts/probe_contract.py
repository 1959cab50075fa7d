"""GPU probe: tcgen05 / SIMT contraction kernel against torch (development aid, not a test)."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from avcer_b200 import ops, _lib

_lib.require_device()
dev = "cuda"
torch.manual_seed(0)


def ref_conv(x, w4, bias, stride, pad, residual, act):
    # x [N,H,W,C] -> torch NCHW fp32
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w4.float(), bias, stride=stride, padding=pad)
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    if act == ops.ACT_RELU:
        y = F.relu(y)
    elif act == ops.ACT_GELU:
        y = F.gelu(y)
    return y


def run_conv(name, n, h, w, cin, cout, k, stride, dtype, residual=False, act=ops.ACT_RELU):
    try:
        x = torch.randn(n, h, w, cin, device=dev).to(dtype)
        w4 = (torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5).to(dtype)
        bias = torch.randn(cout, device=dev)
        wt = w4.permute(0, 2, 3, 1).reshape(cout, k * k * cin).contiguous()
        pad = (k - 1) // 2
        ho = (h + 2 * pad - k) // stride + 1
        wo = (w + 2 * pad - k) // stride + 1
        res = torch.randn(n, ho, wo, cout, device=dev).to(dtype) if residual else None
        y = ops.conv2d_nhwc(x, wt, bias, kh=k, kw=k, stride=stride, pad_h=pad, pad_w=pad, residual=res, act=act)
        torch.cuda.synchronize()
        r = ref_conv(x, w4, bias, stride, pad, res, act)
        err = (y.float() - r).abs().max().item()
        tol = 0.06 if dtype == torch.bfloat16 else 1e-3
        print(f"[{'OK ' if err < tol else 'BAD'}] {name}: max|err|={err:.4g} ref_max={r.abs().max().item():.3g}", flush=True)
        if err >= tol:
            bad = ((y.float() - r).abs() > tol).nonzero()
            print("   first bad idx:", bad[:5].tolist(), "count", bad.shape[0], "of", r.numel(), flush=True)
    except Exception:
        print(f"[EXC] {name}")
        traceback.print_exc()
        sys.stdout.flush()


def run_linear(name, m, k, n, dtype, act=ops.ACT_NONE, residual=False, out_f32=False):
    try:
        x = torch.randn(m, k, device=dev).to(dtype)
        w = (torch.randn(n, k, device=dev) / k ** 0.5).to(dtype)
        b = torch.randn(n, device=dev)
        res = torch.randn(m, n, device=dev).to(dtype) if residual else None
        y = ops.linear(x, w, b, residual=res, act=act, out_dtype=torch.float32 if out_f32 else None)
        torch.cuda.synchronize()
        r = x.float() @ w.float().t() + b
        if res is not None:
            r = r + res.float()
        if act == ops.ACT_GELU:
            r = F.gelu(r)
        if act == ops.ACT_RELU:
            r = F.relu(r)
        err = (y.float() - r).abs().max().item()
        tol = 0.06 if (dtype == torch.bfloat16 and not out_f32) else 2e-3
        print(f"[{'OK ' if err < tol else 'BAD'}] {name}: max|err|={err:.4g}", flush=True)
    except Exception:
        print(f"[EXC] {name}")
        traceback.print_exc()
        sys.stdout.flush()


bf, f32 = torch.bfloat16, torch.float32
print("device:", torch.cuda.get_device_name(0), "sms", _lib.load().avcer_num_sms(), flush=True)
run_linear("linear f32 300x256x128", 300, 256, 128, f32)
run_linear("linear bf16 128x64x128", 128, 64, 128, bf)
run_linear("linear bf16 300x256x128", 300, 256, 128, bf)
run_linear("linear bf16 1000x512x256 gelu+res", 1000, 512, 256, bf, act=ops.ACT_GELU, residual=True)
run_linear("linear bf16 4096x1024x4096", 4096, 1024, 4096, bf)
run_linear("linear bf16 199x1024x64 (BN=64)", 199, 1024, 64, bf)
run_linear("linear bf16 out f32 257x512x2048", 257, 512, 2048, bf, out_f32=True)
run_linear("linear bf16 K=32 (BK=32) 500x32x64", 500, 32, 64, bf)
run_linear("linear bf16 K=96 (BK=32) 500x96x64", 500, 96, 64, bf)
run_conv("conv1x1 f32 4x55x55 64->256", 4, 55, 55, 64, 256, 1, 1, f32)
run_conv("conv3x3 f32 2x14x14 64->64", 2, 14, 14, 64, 64, 3, 1, f32)
run_conv("conv1x1 s2 f32 2x55x55 64->128", 2, 55, 55, 64, 128, 1, 2, f32)
run_conv("conv1x1 bf16 4x55x55 64->256", 4, 55, 55, 64, 256, 1, 1, bf)
run_conv("conv1x1 bf16 4x55x55 256->64 res", 4, 55, 55, 256, 64, 1, 1, bf, residual=True)
run_conv("conv3x3 bf16 4x55x55 64->64", 4, 55, 55, 64, 64, 3, 1, bf)
run_conv("conv3x3 bf16 8x28x28 128->128", 8, 28, 28, 128, 128, 3, 1, bf)
run_conv("conv3x3 bf16 32x14x14 256->256", 32, 14, 14, 256, 256, 3, 1, bf)
run_conv("conv3x3 bf16 5x7x7 512->512", 5, 7, 7, 512, 512, 3, 1, bf)
run_conv("conv1x1 s2 bf16 4x55x55 256->128", 4, 55, 55, 256, 128, 1, 2, bf)
run_conv("conv1x1 s2 bf16 3x14x14 1024->512 res", 3, 14, 14, 1024, 512, 1, 2, bf, residual=True)

# two-SM kernel, FLAT per-warp epilogue: ragged M (odd tile count, rows not a multiple of 32), residual, activations
run_linear("flat bf16 50003x256x1024 relu+res", 50003, 256, 1024, bf, act=ops.ACT_RELU, residual=True)
run_linear("flat bf16 50176x1024x256 relu", 50176, 1024, 256, bf, act=ops.ACT_RELU)
run_linear("flat bf16 12736x1024x4096 gelu", 12736, 1024, 4096, bf, act=ops.ACT_GELU)
run_linear("flat bf16 12736x4096x1024 res", 12736, 4096, 1024, bf, residual=True)
run_linear("flat bf16 20001x64x256 res", 20001, 64, 256, bf, residual=True)
run_conv("conv1x1 bf16 256x14x14 256->1024 res (flat)", 256, 14, 14, 256, 1024, 1, 1, bf, residual=True)

# quick timing of a big GEMM
try:
    m, k, n = 8192, 4096, 4096
    x = torch.randn(m, k, device=dev).to(bf); w = torch.randn(n, k, device=dev).to(bf); b = torch.zeros(n, device=dev)
    out = torch.empty(m, n, device=dev, dtype=bf)
    for _ in range(3): ops.linear(x, w, b, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.linear(x, w, b, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"GEMM 8192x4096x4096 bf16: {ms:.3f} ms  {2*m*k*n/ms/1e9:.1f} TFLOP/s", flush=True)
except Exception:
    traceback.print_exc()
print("probe done", flush=True)
