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


def _configure(blob):
    """IAMF_decoder_configure of the drop-in library on `blob`; returns (return code, bytes consumed)"""
    import ctypes as C
    import iamfapi
    api = iamfapi.Api(os.path.join(ROOT, "iac_b200", "libiamf.so"))
    h = api.L.IAMF_decoder_open()
    api.L.IAMF_decoder_output_layout_set_sound_system(h, 0)
    used = C.c_uint32(0)
    rc = api.L.IAMF_decoder_configure(h, blob, len(blob), C.byref(used))
    api.L.IAMF_decoder_close(h)
    return rc, used.value


def _c1_stream():
    import refstreams
    import scenarios as S
    sc, st, api_kw, unit_kw = refstreams.case("c1")
    inputs = S.synth_inputs(sc, 1, 1, seed=3)
    P, _, _ = S.synth_params(sc, 1, 1, seed=4)
    return st.descriptors(), refstreams.temporal_units(sc, st, inputs, P, unit_kw, 0)[0]


def test_configure_never_steps_past_the_buffer():
    # (found by the fuzzer) a stream whose first OBU is no sequence header: the reference re-splits that OBU until the
    # position passes the end of the buffer and then parses `size - pos` bytes from there; this library stops at the end
    desc, unit = _c1_stream()
    rc, used = _configure(unit[:200])          # an audio frame OBU first, cut short
    assert rc != 0 and used <= 200
    rc, used = _configure(desc[6:])            # descriptors without the sequence header in front
    assert rc != 0 and used <= len(desc) - 6


def test_absurd_codec_configs_are_refused():
    # (found by the fuzzer) 128 Hz in the ipcm decoder config: one 960-sample frame would be 360 000 samples at the 48 kHz
    # output rate - far beyond what max_frame_size tells the caller to allocate
    desc, unit = _c1_stream()
    i = desc.index(b"ipcm")
    rate_at = i + 4 + 2 + 2 + 2                # fourcc, frame size (leb128 x 2), roll distance, format flags + sample size
    assert int.from_bytes(desc[rate_at:rate_at + 4], "big") == 48000
    bad = desc[:rate_at] + (128).to_bytes(4, "big") + desc[rate_at + 4:]
    rc_good, _ = _configure(desc + unit)       # 0 on a GPU box, IAMF_ERR_INTERNAL (-3: no device) on the CPU box
    rc, _ = _configure(bad + unit)
    assert rc_good in (0, -3)
    assert rc not in (0, -3)                   # the damaged config never completes the descriptor set: no plan is ever built from it
