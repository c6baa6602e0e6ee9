#!/usr/bin/env python
"""What the GPU box looks like (development aid): topology, NUMA, NCCL library, CUDA IPC between processes."""
import glob
import os
import subprocess
import sys


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:
        return "ERR %s" % e


def child(q, r):
    import torch
    t = q.get()
    t += 1
    torch.cuda.synchronize()
    r.put(float(t.sum().item()))


def main():
    print("== nvidia-smi -L\n" + sh("nvidia-smi -L"))
    print("== topo\n" + sh("nvidia-smi topo -m"))
    print("== lscpu\n" + sh("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)|Thread'"))
    print("== mem\n" + sh("free -g | head -2"))
    print("== affinity", sorted(os.sched_getaffinity(0))[:8], "...", len(os.sched_getaffinity(0)))
    for d in glob.glob("/sys/bus/pci/devices/*/numa_node"):
        try:
            cls = open(os.path.dirname(d) + "/class").read().strip()
            if cls.startswith("0x0302") or cls.startswith("0x0300"):
                print("gpu pci", os.path.dirname(d)[-12:], "numa", open(d).read().strip(), "cpus", open(os.path.dirname(d) + "/local_cpulist").read().strip())
        except Exception:
            pass
    print("== nccl libs\n" + sh("ldconfig -p | grep -i nccl; python -c \"import nvidia.nccl, os; print(os.path.dirname(nvidia.nccl.__file__)); print(os.listdir(os.path.join(os.path.dirname(nvidia.nccl.__file__), 'lib')))\"; ls /usr/include/nccl.h /usr/lib/x86_64-linux-gnu/libnccl* 2>&1"))
    import torch
    print("torch", torch.__version__, "cuda", torch.version.cuda, "nccl", torch.cuda.nccl.version(), "devices", torch.cuda.device_count())
    p = torch.cuda.get_device_properties(0)
    print("pci", p.pci_domain_id, p.pci_bus_id, p.pci_device_id, "sm", p.multi_processor_count)
    import torch.multiprocessing as mp
    mp.set_start_method("spawn", force=True)
    q, r = mp.Queue(), mp.Queue()
    pr = mp.Process(target=child, args=(q, r))
    pr.start()
    t = torch.zeros(1024, device="cuda")
    q.put(t)
    try:
        print("== cuda ipc (torch tensor to a spawned process): child sum", r.get(timeout=120), "parent sees", float(t.sum().item()))
    except Exception as e:
        print("== cuda ipc FAILED", e)
    pr.join(10)


if __name__ == "__main__":
    main()
