# A/B of AVCER_PDL on one box (same library)
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nets.py -x -q -m gpu 2>&1 | tail -2
for pdl in 0 1; do
  AVCER_PDL=$pdl python scripts/time_a_forward.py | sed "s/^/PDL=$pdl /"
  AVCER_PDL=$pdl python scripts/time_vs_layers.py quiet | tail -1 | sed "s/^/PDL=$pdl /"
done
for i in 1 2; do
  for pdl in 0 1; do
    AVCER_PDL=$pdl timeout 250 python bench.py --clips-per-gpu 2 --steps 5 --warmup 2 --skip-cpu-baseline --skip-e2e > gpurun_out/ab_pdl$pdl.log 2>&1
    python - <<PY
import json
l=[x for x in open("gpurun_out/ab_pdl$pdl.log") if x.startswith("{")]
d=json.loads(l[-1]); print("PDL=$pdl", round(d["value"]), round(d["ms_per_step"],2), round(d["vs_resnet50_b256"]["ms"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
  done
done
