"""Host placement helper (pandasarrow_b200/numa.py) against a fake sysfs tree: the process is bound to the CPUs of the
device's NUMA node for the duration of the block and gets its old mask back."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    # by path: importing the package would load libpa_b200.so, which this test does not need
    spec = importlib.util.spec_from_file_location("pa_numa", os.path.join(ROOT, "pandasarrow_b200", "numa.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _fake_sysfs(tmp_path, pci, node, cpulist):
    d = tmp_path / "bus" / "pci" / "devices" / pci
    d.mkdir(parents=True)
    (d / "numa_node").write_text(f"{node}\n")
    if node >= 0:
        nd = tmp_path / "devices" / "system" / "node" / f"node{node}"
        nd.mkdir(parents=True)
        (nd / "cpulist").write_text(cpulist + "\n")
    return str(tmp_path)


def test_cpulist_parser():
    m = _load()
    assert m._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert m._parse_cpulist("") == set()


def test_bind_and_restore(tmp_path):
    m = _load()
    before = os.sched_getaffinity(0)
    one = min(before)
    sysfs = _fake_sysfs(tmp_path, "0000:1b:00.0", 1, f"{one}")
    with m.bind_to_device_numa(0, sysfs=sysfs, pci="0000:1b:00.0") as b:
        assert b.info["bound"] and b.info["numa_node"] == 1 and b.info["cpus"] == 1
        assert os.sched_getaffinity(0) == {one}
    assert os.sched_getaffinity(0) == before


def test_no_node_or_foreign_cpus_leave_the_mask_alone(tmp_path):
    m = _load()
    before = os.sched_getaffinity(0)
    sysfs = _fake_sysfs(tmp_path, "0000:1b:00.0", -1, "")
    with m.bind_to_device_numa(0, sysfs=sysfs, pci="0000:1b:00.0") as b:
        assert not b.info["bound"] and os.sched_getaffinity(0) == before
    sysfs2 = _fake_sysfs(tmp_path / "b", "0000:1c:00.0", 0, str(max(before) + 1000))
    with m.bind_to_device_numa(0, sysfs=sysfs2, pci="0000:1c:00.0") as b:
        assert not b.info["bound"] and "available" in b.info["why"]
    assert os.sched_getaffinity(0) == before
