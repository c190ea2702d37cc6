"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints ONE JSON line with the keys
the driver reads; bench.py never imports the oracle outside its CPU-baseline legs."""
import json
import os
import re
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "conformations/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_oracle_is_only_used_by_checker_code():
    """The product package never imports oracle/ (tier rule: only tests/, smoke() and bench.py's CPU-baseline legs may)."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for base, _, files in os.walk(os.path.join(ROOT, "molearn_b200")):
        for f in files:
            if f.endswith(".py"):
                assert not pat.search(open(os.path.join(base, f)).read()), os.path.join(base, f)
    src = open(os.path.join(ROOT, "bench.py")).read()
    for m in pat.finditer(src):
        # every oracle import in bench.py sits inside one of the baseline legs (CPU port, or the same port on the GPU)
        head = src[:m.start()]
        fn = re.findall(r"^def (\w+)\(", head, re.M)[-1]
        assert fn in ("cpu_port_rate", "run_reference", "run_scan", "run_config2", "run_config4", "run_config5", "torch_cuda_rate"), fn
