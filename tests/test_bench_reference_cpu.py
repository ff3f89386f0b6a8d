"""The reference arm of bench.py (the reference algorithm on host cores) runs without a GPU and prints one JSON line
with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-sample-mb", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "compress_MBps_level9_byte_identical" and d["unit"] == "MB/s"
    assert d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("text100m level 9")


def test_traffic_file_names_kernels_the_library_reports():
    """profiles/traffic.json (ncu DRAM bytes per launch) is keyed by the kernel names bz2b200_kernel_stats reports."""
    src = open(os.path.join(ROOT, "bzip2_rust_b200", "csrc", "common.cuh")).read()
    names = set(json.loads("[" + src.split("kKernelNames[] = {")[1].split("};")[0] + "]"))
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for k in t:
        if k.startswith("_") or "(" in k:
            continue
        assert k in names, k
