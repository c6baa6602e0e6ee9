#!/bin/bash
# One gpurun call that produces everything profiles/ holds for a round:
#   tools/capture_profiles.sh TAG        (run on the GPU box from the repo root; outputs under gpurun_out/)
# 1. the GPU test suite, 2. bench.py and its reference arm (default arguments), 3. the ncu launch list of a short bench run
# (launch-by-launch path: ncu cannot see kernel nodes of a graph with a conditional node), 4. one `ncu --set full` capture of
# the step's longest kernels.  Numbers printed under ncu are never bench values.
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1
tail -3 $O/${TAG}_tests.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
COOPERMAP_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --no-kernel-pass > $O/${TAG}_launches.log 2>&1
COOPERMAP_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on \
    --kernel-name regex:'sr_ring_kernel|search_kernel|search_hard_kernel|fit_solve_kernel|vox_segment_kernel|solve_warp_kernel' \
    --launch-skip 126 --launch-count 14 -f -o $O/${TAG}_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-kernel-pass > $O/${TAG}_full.log 2>&1
ls -la $O | tail -12
