"""CPU-only checks of the measurement tooling: the reference arm of bench.py (no GPU, no product code on its path) prints
the contract's JSON line, and the host-side build products the boundary needs exist and export what they should."""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "products",
                        "--scale", "0.02", "--batch", "1000", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(line) == 1
    j = json.loads(line[0])
    assert j["impl"] == "reference" and j["metric"] == "sampled+gathered seeds/sec" and j["unit"] == "seeds/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["steps"] == 2 and j["warmup"] == 1
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "seeds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["workload"].startswith("products-shaped") and j["config"]["scale"] == 0.02
    # the CPU arm never loads the product library
    assert "liblegion_b200" not in r.stderr


def test_workload_label_is_shared_by_both_arms():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.workload_label("ukunion") == "ukunion-shaped synthetic graph (BASELINE.json configs[3] shape, 128-d)"
    args = bench.argparse.Namespace(workload="ukunion", scale=1.0, batch=0)
    shape = bench.shape_of(args)
    assert shape["N"] == 133_633_040 and shape["D"] == 128 and shape["fanout"] == [25, 10] and shape["batch"] == 8000


def test_server_library_and_pybind_module_are_built():
    """build/lib/libserver.so (reference Makefile: -lserver) exports the in-process API; the pybind module exposes Run"""
    lib = os.path.join(ROOT, "sampling_server", "build", "lib", "libserver.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g
        g.build()
    L = ctypes.CDLL(lib)
    assert hasattr(L, "_Z12NewGPUServerv") and hasattr(L, "_Z12NewGPURunnerv")
    code = ("import sys; sys.path.insert(0, %r); import sampling_server as s; "
            "assert callable(s.Run) and 'fanout' in s.Run.__doc__; print('ok')" % os.path.join(ROOT, "sampling_server"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
