"""The GPU test modules must at least import and collect on a CPU-only box (they run
unattended on the GPU box at the end of a round)."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_tests_collect_without_a_gpu():
    out = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu",
                          "--collect-only", "-q", "-p", "no:cacheprovider"], cwd=REPO,
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    n_tests = sum(1 for ln in out.stdout.splitlines() if "::" in ln)
    assert n_tests >= 120, n_tests
