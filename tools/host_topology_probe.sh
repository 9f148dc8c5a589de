#!/bin/bash
# What the GPU box's host looks like to this container (for the e2e path's pinned-memory placement).
nproc; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
nvidia-smi topo -m 2>&1 | head -20
python - <<'PY'
import os, pynvml
pynvml.nvmlInit()
for i in range(pynvml.nvmlDeviceGetCount()):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    try:
        m = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        print("gpu", i, "cpu affinity words", [hex(x) for x in m])
    except Exception as e:
        print("gpu", i, "affinity query failed", e)
    try:
        print("  numa node id", pynvml.nvmlDeviceGetNumaNodeId(h))
    except Exception as e:
        print("  numa id query failed", e)
print("allowed cpus", sorted(os.sched_getaffinity(0)))
PY
