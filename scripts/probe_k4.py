import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import ops, get_weights_matrices as gwm
dev = "cuda"
n = 1_500_000
ps = [torch.softmax(torch.randn(n, 7, device=dev), 1).contiguous() for _ in range(3)]
lab = torch.empty((4, n), device=dev, dtype=torch.int64)
flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
def t(fn):
    ts = []
    for i in range(6):
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts[2:])
w1 = gwm.class_weights(gwm.weights_3)
print("f64 path, mask, w2=1     ", t(lambda: ops.fuse_compound(*ps, w1, [1, 1, 1], False, True, labels=lab)))
print("f64 path, rule2+mask, w2 ", t(lambda: ops.fuse_compound(*ps, w1, [0.16, 0.36, 0.01], True, True, labels=lab)))
print("f32 path (no weights)    ", t(lambda: ops.fuse_compound(*ps, None, [1, 1, 1], False, True, labels=lab)))
x = torch.empty(n * 21 // 2, device=dev); y = torch.empty(n * 8, device=dev)
print("torch copy of same bytes ", t(lambda: (y.copy_(x[: n * 8]), x.sum())))
