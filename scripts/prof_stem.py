import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import nets, synthetic as syn, ops
net = nets.VSNet(syn.make_vs_state_dict(0, "default"), "bf16", "cuda:0")
x = net.alloc_input(256)
x.normal_()
for _ in range(3):
    y = net.stem(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = net.stem(x)
e1.record(); torch.cuda.synchronize()
print("stem ms", e0.elapsed_time(e1) / 10)
