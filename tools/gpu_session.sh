#!/bin/bash
# One gpurun call of a development round: GPU parity tests, the C3 bench line, the
# ncu launch list of the same bench command and a `--set full` capture of the
# pair-sum kernels. Everything lands in gpurun_out/<tag>_*.
#   usage: tools/gpu_session.sh <tag> [parts]   parts: any of t (tests) b (bench) l (launch list) f (full capture) s (smoke)
tag=${1:-r01}
parts=${2:-tsblf}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
has() { [[ "$parts" == *"$1"* ]]; }
if has s; then timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; fi
if has t; then timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log; fi
if has b; then timeout 600 python bench.py > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench_c3.json; fi
if has l; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
  echo "launch list rc=$?"
fi
if has f; then
  # One substep's kernels (boundary extrapolation, the five wall stages, k_rhs), then k_shift_sums.
  # Reports must stay small: gpurun brings back at most 64 MiB.
  timeout 900 ncu --set full --clock-control none --profile-from-start off -k 'regex:k_rhs|k_ws|k_we|k_wc|k_wf|k_setup' -c ${NCU_COUNT:-7} \
    -o gpurun_out/${tag}_full -f python tools/prof_step.py 3 ${NCOL_FULL:-110} 1 > gpurun_out/${tag}_full.log 2>&1
  echo "full capture rc=$?"
  timeout 900 ncu --set full --clock-control none --profile-from-start off -k 'regex:k_shift_sums' -c 1 \
    -o gpurun_out/${tag}_full_shift -f python tools/prof_step.py 3 ${NCOL_FULL:-110} 1 > gpurun_out/${tag}_full_shift.log 2>&1
  echo "full capture (shift) rc=$?"
  du -sh gpurun_out
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.csv 2>&1
ls -la gpurun_out | tail -20
