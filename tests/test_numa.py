"""Host-memory placement helper (nanocompore_b200/numa.py) against a fake sysfs tree."""
import os

from nanocompore_b200 import numa


def _fake_sysfs(root, nodes, gpu_node, allowed):
    """Two-node box: node k owns cores [k*4, k*4+4); one GPU at 0000:1b:00.0 on ``gpu_node``."""
    dev = root / 'bus' / 'pci' / 'devices' / '0000:1b:00.0'
    dev.mkdir(parents=True)
    (dev / 'numa_node').write_text(f"{gpu_node}\n")
    (dev / 'local_cpulist').write_text("0-7\n")
    sysnode = root / 'devices' / 'system' / 'node'
    sysnode.mkdir(parents=True)
    (sysnode / 'online').write_text(f"0-{nodes - 1}\n" if nodes > 1 else "0\n")
    for k in range(nodes):
        d = sysnode / f'node{k}'
        d.mkdir()
        (d / 'cpulist').write_text(f"{k * 4}-{k * 4 + 3}\n")
    return str(root)


def test_parse_cpulist():
    assert numa.parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    assert numa.parse_cpulist('') == set()
    assert numa.parse_cpulist('5') == {5}


def test_gpu_node_and_cpus_from_sysfs(tmp_path):
    sysfs = _fake_sysfs(tmp_path, 2, 1, None)
    node, cpus = numa.gpu_numa_node('0000:1b:00.0', sysfs)
    assert node == 1 and cpus == {4, 5, 6, 7}      # the node's cpulist wins over local_cpulist
    assert numa.n_numa_nodes(sysfs) == 2
    assert numa.gpu_numa_node('0000:ff:00.0', sysfs) == (None, set())


def test_bind_reports_without_applying(tmp_path):
    sysfs = _fake_sysfs(tmp_path, 2, 0, None)
    before = os.sched_getaffinity(0)
    info = numa.bind_rank_to_gpu_node(0, sysfs=sysfs, pci_address='0000:1b:00.0', apply=False)
    assert os.sched_getaffinity(0) == before
    assert info['node'] == 0 and info['nodes'] == 2 and not info['applied']
    assert info['cpus'] == len(before & {0, 1, 2, 3}) or info['why']


def test_bind_is_a_noop_on_one_node_or_unknown_node(tmp_path):
    one = _fake_sysfs(tmp_path / 'a', 1, 0, None)
    info = numa.bind_rank_to_gpu_node(0, sysfs=one, pci_address='0000:1b:00.0')
    assert not info['applied'] and info['why'] == 'one NUMA node'
    unknown = _fake_sysfs(tmp_path / 'b', 2, -1, None)
    info = numa.bind_rank_to_gpu_node(0, sysfs=unknown, pci_address='0000:1b:00.0')
    assert not info['applied'] and info['node'] is None


def test_bind_applies_and_restores(tmp_path):
    before = os.sched_getaffinity(0)
    if len(before) < 2:
        return
    # a tree whose node 0 holds exactly one of the cores this process may use
    core = min(before)
    root = tmp_path
    dev = root / 'bus' / 'pci' / 'devices' / '0000:1b:00.0'
    dev.mkdir(parents=True)
    (dev / 'numa_node').write_text("0\n")
    sysnode = root / 'devices' / 'system' / 'node'
    (sysnode / 'node0').mkdir(parents=True)
    (sysnode / 'node1').mkdir(parents=True)
    (sysnode / 'online').write_text("0-1\n")
    (sysnode / 'node0' / 'cpulist').write_text(f"{core}\n")
    (sysnode / 'node1' / 'cpulist').write_text("4096\n")
    try:
        info = numa.bind_rank_to_gpu_node(0, sysfs=str(root), pci_address='0000:1b:00.0')
        assert info['applied'] and info['cpus'] == 1
        assert os.sched_getaffinity(0) == {core}
    finally:
        os.sched_setaffinity(0, before)   # (the memory policy now prefers node 0, which exists)
