"""bench.py on the CPU: the reference arm (`--impl reference`) prints ONE JSON line with the keys the
driver reads, on a tiny configuration; helpers of the GPU arm that need no device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--binsize", "1000000",
                          "--nnz", "2e6", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ICE pixels/s" and d["unit"] == "pixels/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] >= 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["binsize"] == 1000000 and "workload" in d["config"]


def test_cpulist_parser_and_numa_binding_is_a_no_op_without_topology():
    sys.path.insert(0, ROOT)
    import bench
    assert bench._cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert bench._cpulist("") == set()

    class _Props:
        pci_domain_id, pci_bus_id, pci_device_id = 0xffff, 0xff, 0x1f      # no such device under /sys

    class _Torch:
        class cuda:
            @staticmethod
            def get_device_properties(i):
                return _Props

    before = os.sched_getaffinity(0)
    assert bench.bind_host_to_gpu_node(_Torch, 0) is None
    assert os.sched_getaffinity(0) == before
