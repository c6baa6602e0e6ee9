for S in 192 256; do python bench.py --streams $S --no-cpu-baseline --no-latency --no-kernel-pass --no-pcl-arm --no-xyz12-arm 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('streams',$S,'value %.3e'%d['value'],'ms %.3f'%d['ms_per_step'],'e2e16 %.3e'%d['e2e']['value'])"; done
