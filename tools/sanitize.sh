#!/bin/bash
# compute-sanitizer over the hot path (the GPU counterpart of the reference's `tox -e asan`,
# /root/reference/tox.ini:68-78): memcheck, racecheck (shared-memory hazards: mbarriers, byte
# counters, look-back), initcheck, synccheck over smoke() and the tiny-record / outlier parity cases.
# Usage (on the GPU box): tools/sanitize.sh [outdir]   -> <outdir>/sanitize_<tool>.log + summary
set -uo pipefail
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
SMOKE='import __graft_entry__ as g; g.smoke()'
TESTS="tests/test_gpu_parity.py -k 'tiny_records or outliers or odd_content or escalations_over or newline_runs or invalid_phred'"
: > "$OUT/sanitize_summary.txt"
for tool in memcheck racecheck initcheck synccheck; do
    extra=""
    [ "$tool" = memcheck ] && extra="--leak-check no"
    timeout 1500 $SAN --tool $tool $extra --error-exitcode 9 --print-limit 20 \
        python -c "$SMOKE" > "$OUT/sanitize_${tool}_smoke.log" 2>&1
    rc1=$?
    rc2=skipped
    if [ "$tool" = memcheck ] || [ "$tool" = racecheck ]; then
        eval timeout 1500 $SAN --tool $tool $extra --error-exitcode 9 --print-limit 20 \
            python -m pytest -x -q -m gpu $TESTS > "$OUT/sanitize_${tool}_tests.log" 2>&1
        rc2=$?
    fi
    s1=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$OUT/sanitize_${tool}_smoke.log" | tail -1)
    s2=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$OUT/sanitize_${tool}_tests.log" 2>/dev/null | tail -1)
    echo "$tool smoke rc=$rc1 [$s1] tests rc=$rc2 [$s2]" | tee -a "$OUT/sanitize_summary.txt"
done
