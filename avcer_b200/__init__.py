"""avcer_b200: B200-native (sm_100a) implementation of AVCER's batched inference-and-fusion path.

Host code is Python/PyTorch (plumbing: memory, streams, torch.distributed); all computation runs
in hand-written CUDA kernels of libavcer_b200.so, reached through a C ABI (include/avcer_b200.h).
"""
__version__ = "0.1.0"
