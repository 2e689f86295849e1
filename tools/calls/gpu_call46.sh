#!/bin/bash
nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core|^CPU\(s\)" | head -12
nvidia-smi topo -m 2>&1 | head -20
python - <<'PY'
import os, pynvml as n
n.nvmlInit()
print('affinity now', sorted(os.sched_getaffinity(0)))
for i in range(n.nvmlDeviceGetCount()):
    h = n.nvmlDeviceGetHandleByIndex(i)
    try:
        m = n.nvmlDeviceGetCpuAffinity(h, 4)
        print(i, [hex(x) for x in m])
    except Exception as e:
        print(i, 'err', e)
PY
cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /proc/loadavg
