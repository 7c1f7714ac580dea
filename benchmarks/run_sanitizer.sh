#!/bin/bash
# compute-sanitizer over the small-shape GPU tests: memcheck on the Davies term builder / dense gap groups /
# jump kernels / C demos, racecheck on the hand-rolled mbarrier/TMA pipelines (ZGEMM, XOR trajectory kernel).
#   gpurun --timeout 900 -- 'bash benchmarks/run_sanitizer.sh'        → gpurun_out/sanitizer_*.log (summaries go to profiles/)
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, tool, time limit, pytest selection…
    local name=$1 tool=$2 lim=$3; shift 3
    echo "== $name ($tool)" | tee -a "$OUT/sanitizer_summary.log"
    timeout "$lim" $CS --tool "$tool" --error-exitcode 86 --print-limit 20 --log-file "$OUT/sanitizer_${name}.log" \
        python -m pytest -x -q -m gpu "$@" > "$OUT/sanitizer_${name}.pytest.log" 2>&1
    echo "   exit code $? (86 = sanitizer errors, 124 = time limit)" | tee -a "$OUT/sanitizer_summary.log"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|Invalid|Hazard" "$OUT/sanitizer_${name}.log" | sort | uniq -c | head -8 | tee -a "$OUT/sanitizer_summary.log"
    tail -2 "$OUT/sanitizer_${name}.pytest.log" | tee -a "$OUT/sanitizer_summary.log"
}
: > "$OUT/sanitizer_summary.log"
run memcheck_davies memcheck 170 tests/test_gpu_davies_groups.py -k "transverse_field_start and 4 or two_davies or gap_indices or library_liouvillian_one_call"
run memcheck_jump memcheck 120 tests/test_gpu_resample.py
run memcheck_parity memcheck 170 tests/test_gpu_parity.py -k "lindblad_full_rhs or davies_closed_form or onesided or sparse_commutator or traj"
run racecheck_zgemm racecheck 170 tests/test_gpu_parity.py -k "test_zgemm and (16-16-16 or 64-64-64 or 130-70-9)"
run racecheck_traj racecheck 170 tests/test_gpu_traj_evolve.py
cat "$OUT/sanitizer_summary.log"
