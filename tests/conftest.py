import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_sessionstart(session):
    """The native libraries are build artefacts (git-ignored): build whatever is missing or stale before any test loads
    them - nvcc cross-compiles sm_100a without a GPU, gcc builds the host layer and the oracle's C restatement.  On the
    GPU box the prebuilt files travel with the snapshot and nothing is rebuilt."""
    import subprocess
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    if os.path.exists("/dev/nvidia0") and os.path.exists(os.path.join(root, "iac_b200", "libiamf_b200.so")) \
            and os.path.exists(os.path.join(root, "iac_b200", "libiamf.so")) and os.path.exists(os.path.join(root, "oracle", "liboracle.so")):
        return    # GPU box: the prebuilt libraries of the snapshot are the ones under test
    try:
        from iac_b200 import build
        build.build_all(force=False)
        subprocess.run(["make", "-s", "-C", os.path.join(root, "oracle"), "liboracle.so"], check=True)
    except Exception as e:  # noqa: BLE001 - the tests that need the libraries will fail with the real reason
        print(f"[conftest] native build skipped: {e}", file=sys.stderr)
