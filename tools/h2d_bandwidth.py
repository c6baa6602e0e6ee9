import torch, time
N = 134217728
a = torch.empty(N, dtype=torch.uint8).pin_memory()
b = torch.empty_like(a, device="cuda")
def run(nstreams, reps=10):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    chunk = N // nstreams
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                b[i*chunk:(i+1)*chunk].copy_(a[i*chunk:(i+1)*chunk], non_blocking=True)
    torch.cuda.synchronize()
    return reps * N / (time.perf_counter() - t0) / 1e9
for n in (1, 2, 4, 8):
    run(n, 2)
    print("streams", n, "H2D GB/s %.1f" % run(n))
# numa / cpu info
import os
print(os.popen("lscpu | grep -i 'numa\|model name\|socket' | head -8").read())
print(os.popen("nvidia-smi topo -m | head -6").read())
