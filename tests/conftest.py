import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "java-columnar-query-engine_b200", ROOT / "oracle", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")
    # a fresh checkout has no built artefacts (they are git-ignored): build libcolq.so (nvcc cross-compiles sm_100a
    # without a GPU) and the oracle once, exactly as __graft_entry__.build() does
    if not (ROOT / "java-columnar-query-engine_b200" / "lib" / "libcolq.so").exists():
        import subprocess
        subprocess.run(["make", "-C", str(ROOT / "java-columnar-query-engine_b200")], check=True, capture_output=True)


@pytest.fixture(scope="session")
def expected():
    return json.loads((ROOT / "tests" / "golden" / "expected.json").read_text())


@pytest.fixture(scope="session")
def base_geography():
    from colq import geography
    return geography.load_base()
