"""Development aid: do the ways a mapping step can be submitted (WHILE graph / launch by launch, warm-started search or not,
estimate-sized or repeated exactly) give identical poses?  Prints the neighbour lists that differ."""
import ctypes as C
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cmb = importlib.import_module("the-cooper-mapper_b200"); synth = importlib.import_module("the-cooper-mapper_b200.synth")
MAP_CFG = dict(filter_corner=0.4, filter_surf=0.8, map_filter_corner=0.4, map_filter_surf=0.4)
def frames(sc, n, cols, seed, speed=0.5):
    return [(R, t, synth.simulate_scan(sc, R, t, "VLP-16", seed=seed + k, cols=cols)) for k, (R, t) in enumerate(synth.trajectory(n, speed=speed))]
sc = synth.make_scene(seed=61, extent=40.0, n_boxes=12, n_poles=10)
S = 2
big = [frames(sc, 6, 900, 300 + 40 * s) for s in range(S)]
small = [frames(sc, 6, 128, 700 + 40 * s) for s in range(S)]
CAP = 16 * 900
def run(env):
    for k in ("COOPERMAP_TEST_UNDERESTIMATE", "COOPERMAP_NO_GRAPH", "COOPERMAP_NO_WARM"):
        os.environ.pop(k, None)
    for k in env: os.environ[k] = "1"
    ctx = cmb.Context(**MAP_CFG); ctx.mapping_create(S, 100000, 600000)
    out = []; dumps = []
    for k in range(6):
        seqs = small if k in (1, 4) else big
        fr = np.stack([seqs[s][k][2] for s in range(S)])
        od = [(seqs[s][k][0].astype(np.float32), seqs[s][k][1].astype(np.float32)) for s in range(S)]
        out.append(ctx.pipeline_step(fr, od))
        cap = fr.shape[1] * fr.shape[2]
        slots = np.empty((S, 2 * cap, 5), np.int32)
        ctx._check(ctx.L.cm_debug_read_slots(ctx.h, slots.ctypes.data_as(C.c_void_p), C.c_size_t(slots.size)))
        q = []
        for cls in (0, 1):
            buf = np.empty((S, cap, 4), np.float32); cnt = np.zeros(S, np.int32)
            ctx._check(ctx.L.cm_debug_read_queries(ctx.h, C.c_int(cls), buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size), cnt.ctypes.data_as(C.c_void_p)))
            q.append((buf, cnt))
        dumps.append((slots, q))
    return out, dumps, ctx
names = (("ref", []), ("und", ["COOPERMAP_TEST_UNDERESTIMATE"]), ("nograph", ["COOPERMAP_NO_GRAPH"]), ("nowarm", ["COOPERMAP_NO_WARM"]), ("nowarm_nograph", ["COOPERMAP_NO_WARM", "COOPERMAP_NO_GRAPH"]))
runs = {n: run(e) for n, e in names}
ref = runs["ref"]
for n, (out, dumps, ctx) in runs.items():
    print(n)
    for k in range(6):
        for s in range(S):
            a = ref[0][k]; b = out[k]
            dt = np.max(np.abs(a[0][s][1] - b[0][s][1])); dR = np.max(np.abs(a[0][s][0] - b[0][s][0]))
            if dt or dR:
                print("  step", k, "stream", s, "dt", dt, "dR", dR, a[1][s], b[1][s])
        sa, qa = ref[1][k]; sb, qb = dumps[k]
        if k != 5:
            continue
        for s in range(S):
            nC, nS = int(qa[0][1][s]), int(qa[1][1][s])
            def geo(cx, sl):
                out = np.empty((nC + nS, 5, 4), np.float32)
                for cls, lo, hi in ((0, 0, nC), (1, nC, nC + nS)):
                    if hi > lo:
                        slc = np.ascontiguousarray(sl[s, lo:hi].reshape(-1), np.int32)
                        buf = np.empty((len(slc), 4), np.float32)
                        cx._check(cx.L.cm_debug_read_map_points(cx.h, C.c_int(s), C.c_int(cls), slc.ctypes.data_as(C.c_void_p), C.c_int(len(slc)), buf.ctypes.data_as(C.c_void_p)))
                        out[lo:hi] = buf.reshape(-1, 5, 4)
                return out
            ga, gb = geo(ref[2], sa), geo(ctx, sb)
            same = np.array([np.array_equal(ga[r].view(np.uint32), gb[r].view(np.uint32)) for r in range(nC + nS)])
            bad = np.where(~same)[0]
            print("  step", k, "stream", s, "rows whose neighbour POINTS differ (set or order):", len(bad), "of", nC + nS)
            qc = np.concatenate([qa[0][0][s, :nC], qa[1][0][s, :nS]])
            for r in bad[:4]:
                print("    row", r, "cls", 0 if r < nC else 1, "query", qc[r, :3].tolist())
                print("      ref  ", ga[r, :, :3].tolist())
                print("      other", gb[r, :, :3].tolist())
