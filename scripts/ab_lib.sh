# A/B of two builds of the library on the same box: gpurun_ab/libavcer_base.so (previous commit) vs the in-tree build.
BASE=$PWD/gpurun_ab/libavcer_base.so
for i in 1 2; do
  AVCER_LIB=$BASE python scripts/time_a_forward.py | sed 's/^/base: /'
  python scripts/time_a_forward.py | sed 's/^/new:  /'
done
AVCER_LIB=$BASE python scripts/time_vs_layers.py quiet | tail -1 | sed 's/^/base: /'
python scripts/time_vs_layers.py quiet | tail -1 | sed 's/^/new:  /'
for i in 1 2; do
  for which in base new; do
    if [ $which = base ]; then export AVCER_LIB=$BASE; else unset AVCER_LIB; fi
    timeout 250 python bench.py --clips-per-gpu 2 --steps 5 --warmup 2 --skip-cpu-baseline --skip-e2e > gpurun_out/ab_$which.log 2>&1
    python - <<PY
import json
l=[x for x in open("gpurun_out/ab_$which.log") if x.startswith("{")]
d=json.loads(l[-1]); print("$which", round(d["value"]), round(d["ms_per_step"],2), round(d["vs_resnet50_b256"]["ms"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
  done
done
