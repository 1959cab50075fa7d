#!/bin/bash
# compute-sanitizer over a reduced forward of the whole path (scripts/sanitize_target.py), one tool per pass:
#   memcheck  -- out-of-bounds / misaligned global, shared and local accesses
#   racecheck -- shared-memory hazards between the warps of a CTA (the class of bug that shipped mid round 1: a residual
#                ring slot released to the TMA refill without a proxy fence)
#   synccheck -- divergent / mismatched barrier use
# Run on a B200 (gpurun); logs land in gpurun_out/sanitize_<tool>.log, summaries are copied to profiles/ by hand.
# AVCER_PDL=0: programmatic dependent launch overlaps kernels on purpose; the sanitizer should see them serialised.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export AVCER_PDL=0
rc_all=0
for tool in ${SANITIZE_TOOLS:-memcheck racecheck synccheck}; do
  log=gpurun_out/sanitize_${tool}.log
  timeout ${SANITIZE_TIMEOUT:-900} /usr/local/cuda/bin/compute-sanitizer --tool "$tool" --print-limit 20 \
      --launch-timeout 0 --error-exitcode 9 python scripts/sanitize_target.py > "$log" 2>&1
  rc=$?
  echo "== $tool rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok' "$log" | tr '\n' ' ')"
  [ $rc -ne 0 ] && rc_all=$rc
done
exit $rc_all
