#!/bin/bash
# One `gpurun --gpus N` call for the multi-GPU evidence of a round:  tools/capture_multigpu.sh TAG N
#   config 4 (ONE 50 M-point map sharded over the ranks, in-library exchange), the sharded-vs-one-GPU identity check on a >= 4 M
#   map, config 3 as quoted (256 VLP-16 streams, stream i on rank i mod N) and the headline config 2 at N GPUs.
TAG=${1:-rXX}; N=${2:-8}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --config 4 --gpus $N --steps 10 --warmup 3 > $O/${TAG}_c4_g$N.json 2> $O/${TAG}_c4_g$N.err
DIST_SPACING=0.18 $TR --master-port 29522 tools/dist_check.py > $O/${TAG}_distcheck_g$N.json 2> $O/${TAG}_distcheck_g$N.err
$TR --master-port 29523 bench.py --config 3 --gpus $N --steps 10 --warmup 3 > $O/${TAG}_c3_g$N.json 2> $O/${TAG}_c3_g$N.err
$TR --master-port 29524 bench.py --gpus $N --no-cpu-baseline --no-latency --no-kernel-pass > $O/${TAG}_c2_g$N.json 2> $O/${TAG}_c2_g$N.err
tail -c 600 $O/${TAG}_c4_g$N.json; tail -3 $O/${TAG}_c4_g$N.err; tail -c 400 $O/${TAG}_distcheck_g$N.json
