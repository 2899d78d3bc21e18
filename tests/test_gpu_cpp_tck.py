"""GPU suite: the reference's QueryTest restated in C++ against the header-only host mirror (cpp/colq.hpp), linked to
libcolq.so through the C ABI only (java-columnar-query-engine_b200/cpp/tck_main.cpp)."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
TCK = ROOT / "java-columnar-query-engine_b200" / "lib" / "colq_tck"


@pytest.mark.gpu
def test_cpp_host_mirror_runs_the_reference_query_tests():
    assert TCK.exists(), "build it with `make -C java-columnar-query-engine_b200`"
    out = subprocess.run([str(TCK)], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "PASSED (0 failure(s))" in out.stdout


def test_cpp_host_mirror_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert TCK.exists(), "build it with `make -C java-columnar-query-engine_b200`"
    out = subprocess.run([str(TCK)], capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "no CPU fallback" in out.stdout
