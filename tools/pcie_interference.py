import sys, os, importlib, threading, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
synth = importlib.import_module(bench.PKG + ".synth"); cmb = importlib.import_module(bench.PKG)
S = 64
mc, ms, frames, poses = bench.make_workload(6, synth)
ctx = cmb.Context(device=0, **bench.CFG)
ctx.mapping_create(S, max_corner_points=max(4 * len(mc), 100000), max_surf_points=int(1.6 * len(ms)) + 200000)
eye = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
for o in range(0, len(ms), 1 << 18):
    ctx.map_insert([mc if o == 0 else mc[:0]] * S, [ms[o:o + (1 << 18)]] * S, [eye] * S)
rng = np.random.default_rng(7); dev = torch.device("cuda", 0)
pool = torch.from_numpy(frames).to(dev)
NB = 6
order = [[(3 * s + b) % len(frames) for s in range(S)] for b in range(NB)]
bufs = [pool[torch.tensor(order[b], device=dev)].contiguous() for b in range(NB)]
mapped = np.empty((S, 12), np.float32); stats = (cmb.MatchStats * S)()
def run(n):
    od = [bench.pack_isos([bench.noisy_odom(poses, order[k % NB][s], rng, synth) for s in range(S)]) for k in range(n)]
    ctx.pipeline_prefetch_dev(bufs[0].data_ptr(), 64, 2048)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(n):
        if k + 1 < n: ctx.pipeline_prefetch_dev(bufs[(k + 1) % NB].data_ptr(), 64, 2048)
        ctx.pipeline_step_dev(bufs[k % NB].data_ptr(), 64, 2048, od[k], mapped, stats)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
run(4)
print("device arm, quiet link: %.2f ms/step" % run(10))
stop = False
src = torch.empty(134217728, dtype=torch.uint8).pin_memory(); dst = torch.empty_like(src, device="cuda")
def bg():
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        while not stop:
            dst.copy_(src, non_blocking=True); st.synchronize()
th = threading.Thread(target=bg); th.start(); time.sleep(0.2)
print("device arm, link busy with 134 MB H2D copies: %.2f ms/step" % run(10))
stop = True; th.join()
ctx.close()
