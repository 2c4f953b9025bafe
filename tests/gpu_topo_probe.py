"""Prints what a rank can see of the box's host topology (NUMA nodes, allowed cores, where each
GPU hangs) -- the inputs of nanocompore_b200/numa.py.  Run on a GPU box; not a test."""
import json
import sys
import os
import subprocess

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nanocompore_b200 import numa


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as e:
        return f"<{e}>"


def main():
    out = {
        "cpu_count": os.cpu_count(),
        "affinity": len(os.sched_getaffinity(0)),
        "affinity_list": sorted(os.sched_getaffinity(0))[:8] + ['...'] + sorted(os.sched_getaffinity(0))[-4:],
        "nodes_online": sh("cat /sys/devices/system/node/online"),
        "node_cpulists": sh("for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist); done"),
        "node_meminfo": sh("grep -h MemTotal /sys/devices/system/node/node*/meminfo"),
        "cgroup_cpuset": sh("cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null"),
        "cpu_max": sh("cat /sys/fs/cgroup/cpu.max 2>/dev/null"),
        "mems_allowed": sh("grep -E 'Mems_allowed_list|Cpus_allowed_list' /proc/self/status"),
        "gpus": [],
    }
    for i in range(torch.cuda.device_count()):
        addr = numa.gpu_pci_address(i)
        node, cpus = numa.gpu_numa_node(addr)
        out["gpus"].append({"index": i, "pci": addr, "node": node, "node_cpus": len(cpus),
                            "bind": numa.bind_rank_to_gpu_node(i, apply=False)})
    print(json.dumps(out, indent=1))
    print(sh("nvidia-smi topo -m"))


if __name__ == '__main__':
    main()
