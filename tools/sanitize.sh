#!/usr/bin/env bash
# compute-sanitizer passes over the small-system parity tests (the reference has no race or
# memory checker; the CUDA library gets memcheck + racecheck + initcheck on the boxes the tests use).
#
#     gpurun --timeout 900 -- bash tools/sanitize.sh [pytest -k expression]
#
# Writes gpurun_out/sanitize_<tool>.log; exits non-zero when a tool reports an error.  The kernels run 10-50x slower
# under the sanitizer: the default selection keeps to the bench systems, the Monte Carlo hooks and the MD steps.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
select="${1:-bench_system or cache_move or wolf_move or ewald_move or integrators_follow or lj_box_cell_list or water_box_cell_list or on_the_cutoff or sorted_resident}"
status=0
for tool in memcheck racecheck initcheck; do
    log="gpurun_out/sanitize_${tool}.log"
    compute-sanitizer --tool "$tool" --error-exitcode 3 --target-processes all \
        python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_parity.py tests/test_gpu_mc.py tests/test_gpu_md.py -q -m gpu -x -k "$select" > "$log" 2>&1
    code=$?
    tail -3 "$log"
    if [ "$code" -ne 0 ]; then
        echo "compute-sanitizer --tool $tool: exit code $code (see $log)"
        status=1
    fi
done
exit $status
