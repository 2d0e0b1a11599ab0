#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the kernels added beside the hot path in round 2: the BAM
# record chain + tiled decode, report aggregation, _seqident, NanoStats name capture, the table-stream fork and
# the block cache (tests that exercise them on small inputs).
# Usage (on the GPU box): tools/sanitize_next_rows.sh [outdir]
set -uo pipefail
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
TESTS="tests/test_gpu_bam_edges.py tests/test_gpu_report.py tests/test_gpu_seqident.py tests/test_gpu_round2.py -k 'device_walk or long_records or nanopore_fastq_headers or nanopore_ubam_tags or odd_durations or length_distribution or known_alignments or other_scores or tag_errors or skipped_reason or pi_warning'"
: > "$OUT/sanitize_next_rows_summary.txt"
for tool in memcheck racecheck synccheck; do
    extra=""
    [ "$tool" = memcheck ] && extra="--leak-check no"
    eval timeout 1500 $SAN --tool $tool $extra --error-exitcode 9 --print-limit 20 --target-processes all \
        python -m pytest -x -q -m gpu $TESTS > "$OUT/sanitize_next_rows_${tool}.log" 2>&1
    rc=$?
    s=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$OUT/sanitize_next_rows_${tool}.log" | sort | uniq -c | tr '\n' ';')
    t=$(grep -E "passed|failed" "$OUT/sanitize_next_rows_${tool}.log" | tail -1)
    echo "$tool rc=$rc [$s] pytest: $t" | tee -a "$OUT/sanitize_next_rows_summary.txt"
done
