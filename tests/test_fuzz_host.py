"""Host layer under AddressSanitizer / UBSan: a short run of the mutation fuzzer (tests/fuzz/) over the five configurations'
bitstreams and two MP4 files - damaged descriptors, OBU headers, parameter blocks, audio frames, box trees; single handles
and the batch entry point.  tests/fuzz/engine_stub.c stands in for the engine and holds the host layer to the buffer
contract of include/iamf_b200.h (it reads / writes every byte the contract names).  Longer runs: tests/fuzz/run.sh."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_mutated_streams_do_not_break_the_host_layer(tmp_path):
    env = dict(os.environ, FUZZ_DIR=str(tmp_path))
    r = subprocess.run(["bash", os.path.join(ROOT, "tests", "fuzz", "run.sh"), "400", "11", "stub"], env=env, capture_output=True, text=True, timeout=600)
    if r.returncode != 0 and "sanitizer" in (r.stderr + r.stdout).lower() and "cannot find" in (r.stderr + r.stdout).lower():
        pytest.skip("no sanitizer runtime for this gcc")
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert r.stderr.count("inputs ok") == 7, r.stderr[-2000:]
