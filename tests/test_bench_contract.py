"""CPU: the reference arm of bench.py prints exactly one JSON line with the contract's keys (on a workload shrunk with
--bodies, which marks the line as reduced — never a bench value). The own arm needs a GPU and is covered by the
driver's bench run itself."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--bodies", "20000",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "body-steps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "REDUCED" in d["config"]["workload"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--bodies", "20000", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout[-500:], r.stderr[-500:])


import pytest


@pytest.mark.gpu
def test_own_arm_prints_one_contract_line():
    """The own arm on a shrunk workload: one JSON line carrying every key of the contract, device and end-to-end
    numbers both present, kernels actually launched."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--bodies", "100000", "--steps", "3",
                        "--warmup", "3", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "roofline_hbm", "roofline_step", "e2e",
              "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["value"] > 0 and d["gpu_launches"] >= 3 * 20
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 40 * 100000
    assert d["e2e"]["d2h_bytes_per_step"] == 32 * 100000
    assert 0 < d["roofline"]["frac"] < 1 and "REDUCED" in d["config"]["workload"]


def test_reference_arm_never_maps_the_product_library():
    """The reference arm's process loads oracle/ and the workload generators only — not liblpe_bh.so."""
    code = ("import sys, json; sys.argv=['bench.py']; import bench; "
            "w=dict(bench.WORKLOADS['c2']); w['n']=5000; bench.reference_sample(w, 1, 0); "
            "maps=open('/proc/self/maps').read(); "
            "print(json.dumps({'product': 'liblpe_bh.so' in maps, 'workloads': 'libworkloads.so' in maps, "
            "'oracle': ('libref_bh.so' in maps) or ('liboracle_bh.so' in maps)}))")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d == {"product": False, "workloads": True, "oracle": True}, d
