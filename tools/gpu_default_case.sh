#!/bin/bash
# The reference's default case (2-D dam break, dr = H/80, to t sqrt(g/H) = 10) and a 3-D
# run through the C++ facade drivers, with the physical sanity report of each.
#   usage: tools/gpu_default_case.sh <tag> [n_col_3d=40] [steps_3d=2000]
tag=${1:-r03}; n3=${2:-40}; s3=${3:-2000}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build_examples()"
t0=$(date +%s.%N)
timeout 200 examples/dam_break_2d 80 0 - /tmp/p2.ttdb > gpurun_out/${tag}_default_case_2d.log 2>&1; echo "2d rc=$? $(python -c "import time,sys; print(round(time.time() - float(sys.argv[1]), 1))" $t0) s"
timeout 120 python tools/default_case_report.py /tmp/p2.ttdb gpurun_out/${tag}_default_case_2d.json gpurun_out/${tag}_default_case_2d_thin.ttdb
t0=$(date +%s.%N)
timeout 200 examples/dam_break_3d $n3 $s3 /tmp/p3.ttdb > gpurun_out/${tag}_case_3d.log 2>&1; echo "3d rc=$? $(python -c "import time,sys; print(round(time.time() - float(sys.argv[1]), 1))" $t0) s"
timeout 120 python tools/default_case_report.py /tmp/p3.ttdb gpurun_out/${tag}_case_3d.json
tail -n 3 gpurun_out/${tag}_default_case_2d.log; tail -n 3 gpurun_out/${tag}_case_3d.log
ls -la /tmp/p2.ttdb /tmp/p3.ttdb
