"""Summarise an ncu `gpu__time_duration.sum` launch list (CSV) per kernel and per VS / A layer."""
import collections, csv, re, sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]; data = rows[hi + 1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit'); mi = hdr.index('Metric Name')
seq = []
for r in data:
    if len(r) <= vi or not r[mi].startswith('gpu__time_duration'): continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    if u == 'ns': v /= 1000
    elif u == 'ms': v *= 1000
    name = re.sub(r'\(.*', '', r[ki]).replace('void avcer::', '').replace('avcer::', '')
    if name.startswith(('conv3x3_kernel', 'stem_pool_kernel')):   # tcgen05 contraction kernels like tc_gemm*
        name = 'tc_gemm:' + name
    seq.append((name, v))
agg = collections.defaultdict(lambda: [0, 0.0])
for n, t in seq:
    agg[n[:60]][0] += 1; agg[n[:60]][1] += t
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.0f} us over {len(seq)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{v[1]:11.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:5d} avg={v[1] / v[0]:8.1f}  {k}")

def vs_specs(B=256):
    specs = [("stem", B * 112 * 112, 64, 147)]
    cin = 64; hw = {1: 55, 2: 28, 3: 14, 4: 7}
    for li, (P, nb) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3)), start=1):
        for bi in range(nb):
            M = B * hw[li] ** 2
            # block 0: the projection shortcut is folded into conv3 (K = Cin + P, nets.VSNet.fused_shortcut); last block of
            # layer1-3: conv2 / conv3 only at the pixels the next stage samples (nets.VSNet.sampled_tail)
            Mt = B * hw[li + 1] ** 2 if (li < 4 and bi == nb - 1) else M
            specs += [(f"l{li}.{bi}.c1", M, P, cin), (f"l{li}.{bi}.c2" + ("/s2" if Mt != M else ""), Mt, P, 9 * P),
                      (f"l{li}.{bi}.c3" + ("+ds" if bi == 0 else "") + ("/s2" if Mt != M else ""), Mt, 4 * P, P + (cin if bi == 0 else 0))]
            cin = 4 * P
    return specs + [("fc1", B, 512, 2048)]

def a_specs(B=32):
    T = [12799, 6399, 3199, 1599, 799, 399, 199]
    specs = [(f"conv{i}", B * T[i], 512, k * 512) for i, k in zip(range(1, 7), (3, 3, 3, 3, 2, 2))]
    specs += [("proj", B * 199, 1024, 512), ("posconv", B * 199, 1024, 64 * 128)]
    for l in range(12):
        specs += [(f"L{l}.qkv", B * 199, 3072, 1024), (f"L{l}.o", B * 199, 1024, 1024), (f"L{l}.ff1", B * 199, 4096, 1024), (f"L{l}.ff2", B * 199, 1024, 4096)]
    for t in ("tl1", "tl2"):
        specs += [(f"{t}.qkv", B * 199, 3072, 1024), (f"{t}.o", B * 199, 1024, 1024), (f"{t}.ff1", B * 199, 1024, 1024), (f"{t}.ff2", B * 199, 1024, 1024)]
    return specs + [("td0", B * 64, 1024, 5120), ("td4", B * 10, 1024, 3072)]

def section(start_prefix, specs, title, brief):
    idx = [i for i, (n, _) in enumerate(seq) if n.startswith(start_prefix)]
    if len(idx) < 2: return
    part = seq[idx[0]:idx[1]]
    # graph warm-up runs the forward twice back to back: keep the first pass only
    ntc = 0
    for j, (n, _) in enumerate(part):
        if n.startswith('tc_gemm'):
            ntc += 1
            if ntc == len(specs):
                part = part[:j + 1 + next((k for k, (m, _) in enumerate(part[j + 1:]) if m.startswith('tc_gemm')), len(part))]
                break
    tc = [(n, t) for n, t in part if n.startswith('tc_gemm')][:len(specs)]
    print(f"\n== {title}: {sum(t for _, t in part):.0f} us total, {sum(t for _, t in tc):.0f} us in tc_gemm ({len(tc)} launches, {len(specs)} expected)")
    totf = 0
    for (n, t), (name, M, N, K) in zip(tc, specs):
        fl = 2 * M * N * K; totf += fl
        if not brief or not re.match(r"L([1-9]|10)\.", name):
            print(f"  {name:10s} M={M:8d} N={N:5d} K={K:5d} {t:8.1f}us {fl / t / 1e6:8.1f} TF/s  out+res~{(M * N * 2) / t / 1e3:7.0f} GB/s  {n[:26]}")
    print(f"  aggregate {totf / sum(t for _, t in tc) / 1e6:.1f} TF/s")
    other = collections.defaultdict(float)
    for n, t in part:
        if not n.startswith('tc_gemm'): other[n[:40]] += t
    for k, v in sorted(other.items(), key=lambda kv: -kv[1]): print(f"  {v:9.1f} us  {k}")

vb = int(sys.argv[3]) if len(sys.argv) > 3 else 256
section('preprocess', vs_specs(vb), f"VS forward, batch {vb}", False)
ab = int(sys.argv[2]) if len(sys.argv) > 2 else 64
section('audio_normalize', a_specs(ab), f"A forward, {ab} windows", True)
