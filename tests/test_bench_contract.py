"""CPU: the driver-facing contract of bench.py that can be checked without a GPU -- the reference arm prints ONE JSON
line with the agreed keys (the unmodified reference from oracle/_ref -- or the oracle port when it is not staged --
timed on the host cores), and the product arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--cpu-sample-batch", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train samples/s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    staged = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "models")) or os.path.isdir("/root/reference/models")
    assert cb["kind"] == ("reference" if staged else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "batch 2" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_configs():
    """BASELINE.json configs[3] (inference): same line shape, its own metric name."""
    r = _run("--impl", "reference", "--config", "infer_bs512", "--steps", "1", "--cpu-sample-batch", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["metric"] == "infer samples/s" and d["config"]["name"] == "infer_bs512" and d["value"] > 0


def test_reference_arm_only_rank0_prints_under_torchrun_env():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_cuda():
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "1", timeout=300)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_product_arm_never_imports_the_oracle():
    """Only the baseline legs (cpu_baseline / --impl reference / gpu_eager_baseline: the reference's own modules being
    TIMED as the incumbent, never the product) may touch oracle/."""
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = [n for n in ast.walk(fn) if (isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] == "oracle")
                or (isinstance(n, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in n.names))]
        if uses:
            assert fn.name in ("cpu_reference_run", "_reference_model"), fn.name
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any((getattr(n, "module", "") or "").startswith("oracle") for n in top)
