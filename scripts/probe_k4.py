"""GPU probe: K4 (fuse_compound) at BASELINE config 4's 1.5 M frames: time and achieved algorithmic GB/s (116 B/frame)."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import ops, get_weights_matrices as gwm
dev = "cuda:0"
for n in (1_500_000, 15_000_000):
    g = torch.Generator(device=dev).manual_seed(7)
    ps = [torch.softmax(torch.randn(n, 7, device=dev, generator=g), 1).contiguous() for _ in range(3)]
    lab = torch.empty((4, n), device=dev, dtype=torch.int64)
    w1, w2 = gwm.class_weights(gwm.weights_3), [1, 1, 1]
    fn = lambda: ops.fuse_compound(ps[0], ps[1], ps[2], w1, w2, False, True, labels=lab)
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(8):
            fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(6):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b) / 8)
    t = statistics.median(ts)
    print(f"K4 n={n}: {t * 1e3:.1f} us  {n * 116 / t / 1e6:.0f} GB/s algorithmic  ({n * 116 / t / 1e6 / 6452.8:.3f} of HBM copy peak)")
    del ps, lab
