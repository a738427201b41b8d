"""The oracle is pinned again ON THE GPU BOX, in the same run as the GPU parity tests: the stage-by-stage and whole-pipeline
comparisons of oracle/ against the compiled, unmodified reference (oracle/_ref, shipped with the snapshot) are CPU tests
(`-m "not gpu"`), which the driver runs in the authoring container only - this wrapper repeats them under `-m gpu`, so the
checker the GPU tests compare against is itself checked where they run."""
import os
import subprocess
import sys

import pytest

import refbind

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_equals_compiled_reference_on_this_box():
    if not refbind.have_ref():
        pytest.skip("compiled reference not shipped")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "not gpu", "-p", "no:cacheprovider",
                        os.path.join(HERE, "test_oracle_vs_ref.py"), os.path.join(HERE, "test_oracle_pipeline_vs_ref.py"),
                        os.path.join(HERE, "test_hrtf_oracle.py")], capture_output=True, text=True, cwd=os.path.dirname(HERE))
    tail = (r.stdout or "")[-600:]
    assert r.returncode == 0, tail
    assert " passed" in tail and "skipped" not in tail.splitlines()[-1], tail
