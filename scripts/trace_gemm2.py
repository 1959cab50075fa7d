"""GPU probe: per-tile clock64 timeline of the two-SM tcgen05 kernel (avcer_debug_set_trace) for a few VS shapes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import ops, _lib

_lib.require_device()
lib = ctypes.CDLL(_lib.LIB_PATH)
dev = "cuda"
bf = torch.bfloat16
CTAS, TILES, SLOTS = 4, 64, 16
NAMES_FLAT = ["prod.begin", "prod.end", "mma.begin", "mma.acc_free", "mma.first_full", "mma.commit", "epi.begin", "epi.tfull",
              "h0.slabfree", "h0.res", "h0.ldtm", "h0.sts", "h0.fence", "h0.store", "h1.store", "epi.release"]
NAMES = ["prod.begin", "prod.end", "mma.begin", "mma.acc_free", "mma.first_full", "mma.commit", "epi.begin", "epi.tfull",
         "epi.h0", "epi.h1", "epi.h2", "epi.h3", "epi.release", "res.last_issue"]


def run(name, n, h, w, cin, cout, k, stride=1, res=False):
    x = torch.randn(n, h, w, cin, device=dev).to(bf)
    wt = (torch.randn(cout, k * k * cin, device=dev) / (cin * k * k) ** 0.5).to(bf)
    b = torch.randn(cout, device=dev)
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    r = torch.randn(n, ho, wo, cout, device=dev).to(bf) if res else None
    call = lambda: ops.conv2d_nhwc(x, wt, b, kh=k, kw=k, stride=stride, pad_h=(k - 1) // 2, pad_w=(k - 1) // 2, residual=r, act=ops.ACT_RELU)
    call(); call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); call(); e1.record(); torch.cuda.synchronize()
    buf = torch.zeros(CTAS * TILES * SLOTS, dtype=torch.int64, device=dev)
    lib.avcer_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    call()
    torch.cuda.synchronize()
    lib.avcer_debug_set_trace(ctypes.c_void_p(0))
    t = buf.view(CTAS, TILES, SLOTS).cpu()
    print(f"== {name}: {e0.elapsed_time(e1) * 1e3:.1f} us")
    for cta in (0, 1):
        base = int(t[cta, 0][t[cta, 0] > 0].min())
        ntile = int((t[cta, :, 6] > 0).sum())
        print(f" CTA {cta}: {ntile} tiles traced; cycles relative to the CTA's first stamp")
        flat = bool((t[cta, :, 15] > 0).any())
        names = NAMES_FLAT if flat else NAMES
        print("   tile " + " ".join(f"{nm[-11:]:>11s}" for nm in names))
        for tl in list(range(0, min(ntile, 8))) + ([ntile - 1] if ntile > 8 else []):
            row = t[cta, tl]
            print(f"   {tl:4d} " + " ".join(f"{(int(v) - base) if v > 0 else -1:11d}" for v in row[:len(names)]))
        if ntile > 3:
            rel = 15 if flat else 12
            per = (int(t[cta, ntile - 2, rel]) - int(t[cta, 1, rel])) / (ntile - 3)
            print(f"   steady-state cycles per tile (epi.release spacing): {per:.0f}")


run("l3.c3 1x1 256->1024 +res (14x14)", 256, 14, 14, 256, 1024, 1, res=True)
run("l3.0.ds 1x1 s2 512->1024 (28->14)", 256, 28, 28, 512, 1024, 1, stride=2)
run("l1.c3 1x1 64->256 +res (55x55)", 256, 55, 55, 64, 256, 1, res=True)

print("trace done")
