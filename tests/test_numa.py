"""Host placement helper of the end-to-end path (openvr_fsr_b200/numa.py) on a fake sysfs: cpulist parsing, PCI id
normalisation, and binding that only ever narrows the current affinity."""
import os

from openvr_fsr_b200 import numa


def test_parse_cpulist():
    assert numa.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert numa.parse_cpulist("") == [] and numa.parse_cpulist("5") == [5]
    assert numa.parse_cpulist("32-63,96-127")[:2] == [32, 33] and len(numa.parse_cpulist("32-63,96-127")) == 64


def test_normalise_pci_bus_id():
    assert numa.normalise_pci_bus_id("00000000:1B:00.0") == "0000:1b:00.0"
    assert numa.normalise_pci_bus_id("0000:9a:00.0") == "0000:9a:00.0"


def test_bind_uses_the_gpus_node_and_only_narrows(tmp_path):
    mine = sorted(os.sched_getaffinity(0))
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices/system/node/node1"
    node.mkdir(parents=True)
    keep = mine[: max(1, len(mine) // 2)]
    (node / "cpulist").write_text(",".join(str(c) for c in keep + [4096]))  # a CPU this process may not use is ignored
    assert numa.gpu_numa_node(0, str(tmp_path), "00000000:1B:00.0") == 1
    try:
        info = numa.bind_to_gpu_node(0, str(tmp_path), "0000:1b:00.0")
        assert info == {"node": 1, "cpus": len(keep), "bound": True}
        assert sorted(os.sched_getaffinity(0)) == keep
    finally:
        os.sched_setaffinity(0, mine)
    # unknown device / node -1: nothing changes
    (dev / "numa_node").write_text("-1\n")
    assert numa.bind_to_gpu_node(0, str(tmp_path), "0000:1b:00.0")["bound"] is False
    assert numa.bind_to_gpu_node(0, str(tmp_path), "0000:ff:00.0")["bound"] is False
    assert sorted(os.sched_getaffinity(0)) == mine
