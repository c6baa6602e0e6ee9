for f in 1 0; do echo "cluster=$f"; COOPERMAP_VOX_CLUSTER=$f timeout 600 python -m pytest tests/test_voxel_gpu.py tests/test_mapping_gpu.py tests/test_configs_gpu.py -m gpu -x -q 2>&1 | tail -3; done
python tools/latency_breakdown.py 2>&1 | grep -E "graph path|vox_"
COOPERMAP_VOX_CLUSTER=0 python tools/latency_breakdown.py 2>&1 | grep -E "graph path|vox_"
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_voxel_gpu.py -m gpu -x -q 2>&1 | tail -5
