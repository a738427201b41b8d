"""The repo's own command-line player (iac_b200/player: WAV writer, MP4 / fragmented-MP4 reader, .met records) against the
reference's stock iamfplayer (compiled unmodified from the reference tree into oracle/_ref/iamfplayer_ref, linked against
the compiled reference decoder): same options in, byte-identical .wav and .met files out."""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mp4gen  # noqa: E402
import refstreams  # noqa: E402
import scenarios as S  # noqa: E402

PLAYER = os.path.join(ROOT, "iac_b200", "iamfplayer_b200")
PLAYER_REF = os.path.join(ROOT, "oracle", "_ref", "iamfplayer_ref")
needs_ref = pytest.mark.skipif(not os.path.exists(PLAYER_REF), reason="stock player not built (make -C oracle ref)")


def stream(case, F=12, seed=77):
    sc, st, api_kw, unit_kw = refstreams.case(case)
    inputs = S.synth_inputs(sc, 1, F, seed=seed)
    P, _, _ = S.synth_params(sc, 1, F, seed=seed + 1)
    refstreams.no_param_gaps(sc, P)
    return sc, st.descriptors(), refstreams.temporal_units(sc, st, inputs, P, unit_kw, 0)


def run(player, args, files, expect_ok=True):
    """runs a player in a scratch directory holding `files` {name: bytes}; returns {output file name: bytes}"""
    d = tempfile.mkdtemp(prefix="player_")
    try:
        for name, blob in files.items():
            with open(os.path.join(d, name), "wb") as f:
                f.write(blob)
        r = subprocess.run([player] + args, cwd=d, capture_output=True, text=True, timeout=300)
        if expect_ok:
            assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
        return {os.path.basename(p): open(p, "rb").read() for p in glob.glob(os.path.join(d, "*.wav")) + glob.glob(os.path.join(d, "*.met"))}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_mp4_reader_finds_the_samples_the_muxer_wrote():
    # no device needed: -probe prints the sample table
    rng = np.random.default_rng(5)
    samples = [bytes(rng.integers(0, 256, int(n), dtype=np.uint8)) for n in rng.integers(5, 400, 23)]
    deltas = [960] * 22 + [500]
    for kw in (dict(), dict(skip=312, samples_per_chunk=5), dict(fragmented=True, per_fragment=4), dict(fragmented=True, per_fragment=1, skip=7),
               dict(desc_of_sample=[1] * 9 + [2] * 14, samples_per_chunk=3)):
        descs = [b"D" * 33, b"E" * 57] if "desc_of_sample" in kw else [b"D" * 33]
        blob, table = mp4gen.mux(descs, samples, deltas, timescale=48000, **kw)
        d = tempfile.mkdtemp(prefix="probe_")
        try:
            p = os.path.join(d, "x.mp4")
            open(p, "wb").write(blob)
            r = subprocess.run([PLAYER, "-probe", p], capture_output=True, text=True, timeout=60)
        finally:
            shutil.rmtree(d, ignore_errors=True)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.strip().splitlines()
        head = lines[0].split()
        assert int(head[1]) == len(samples) and int(head[5]) == 48000 and int(head[7]) == kw.get("skip", 0) and int(head[9]) == len(descs)
        got = [tuple(int(v) for v in ln.split()) for ln in lines[1:]]
        assert got == table, kw
        for (off, size, _, _), s in zip(got, samples):
            assert blob[off:off + size] == s


def test_mp4_reader_rejects_damaged_files():
    blob, _ = mp4gen.mux([b"D" * 20], [b"x" * 50] * 6, [960] * 6)
    d = tempfile.mkdtemp(prefix="probe_")
    try:
        for k, bad in enumerate((blob[:len(blob) - 120], blob[:40], b"", blob.replace(b"iamf", b"mp4a"))):
            p = os.path.join(d, f"bad{k}.mp4")
            open(p, "wb").write(bad)
            r = subprocess.run([PLAYER, "-probe", p], capture_output=True, text=True, timeout=60)
            assert r.returncode != 0 and "error" in r.stdout, (k, r.stdout)
    finally:
        shutil.rmtree(d, ignore_errors=True)


@needs_ref
def test_muxer_is_accepted_by_the_stock_player():
    # the reference's own player gives the same samples from the MP4 as from the plain bitstream (checks the test muxer)
    sc, desc, units = stream("c1")
    a = run(PLAYER_REF, ["-o2", "-s0", "in.iamf"], {"in.iamf": desc + b"".join(units)})
    blob, _ = mp4gen.mux(desc, units, [sc.frame_size] * len(units))
    b = run(PLAYER_REF, ["-i1", "-o2", "-s0", "in.mp4"], {"in.mp4": blob})
    assert len(a["ss0_in.wav"]) > 44 + 1000 and a["ss0_in.wav"] == b["ss0_in.wav"]


CASES = [("c1", ["-s0"]), ("c2", ["-s1"]), ("c3", ["-s7", "-d", "24"]), ("c4", ["-sb"]), ("c5", ["-s0", "-r", "48000", "-l", "-24"]),
         ("c2", ["-s9", "-disable_limiter", "-d", "32"]), ("c1", ["-s1", "-p", "-6"])]


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("case,args", CASES, ids=[c + "".join(a) for c, a in CASES])
def test_bitstream_wav_and_met_match_the_stock_player(case, args):
    sc, desc, units = stream(case)
    files = {"in.iamf": desc + b"".join(units)}
    ref = run(PLAYER_REF, ["-o2", "-m"] + args + ["in.iamf"], files)
    got = run(PLAYER, ["-o2", "-m"] + args + ["in.iamf"], files)
    assert sorted(ref) == sorted(got) and len(ref) == 2, (sorted(ref), sorted(got))
    for name in ref:
        assert len(ref[name]) > 44 and ref[name] == got[name], f"{name} differs ({len(ref[name])} vs {len(got[name])} bytes)"


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("kw,extra", [(dict(), []), (dict(skip=480, samples_per_chunk=5), []), (dict(samples_per_chunk=1), ["-ts", "0"]),
                                      (dict(skip=100), ["-m"])], ids=["plain", "edit-list", "chunk-per-sample", "met"])
def test_mp4_matches_the_stock_player(kw, extra):
    sc, desc, units = stream("c2", F=14)
    blob, _ = mp4gen.mux(desc, units, [sc.frame_size] * len(units), **kw)
    args = ["-i1", "-o2", "-s1"] + extra + ["in.mp4"]
    ref = run(PLAYER_REF, args, {"in.mp4": blob})
    got = run(PLAYER, args, {"in.mp4": blob})
    assert sorted(ref) == sorted(got) and ref, (sorted(ref), sorted(got))
    for name in ref:
        assert len(ref[name]) > 44 and ref[name] == got[name], f"{name} differs ({len(ref[name])} vs {len(got[name])} bytes)"


@pytest.mark.gpu
def test_mp4_seek_starts_at_the_sample_the_reference_rule_names():
    # -ts: whole samples are consumed until start * movie time scale + edit-list skip is used up, the sample that crosses it
    # included (mp4iamfpar.c:203-234).  (The stock player itself crashes on -ts with these files, so the expectation is the
    # rule: 48000 + 200 time units over samples of 960 -> 51 samples consumed; decoding restarts from a fresh decoder.)
    sc, desc, units = stream("c1", F=120)      # 2.4 s
    blob, _ = mp4gen.mux(desc, units, [sc.frame_size] * len(units), skip=200)
    got = run(PLAYER, ["-i1", "-o2", "-s0", "-ts", "1", "in.mp4"], {"in.mp4": blob})
    want = run(PLAYER, ["-o2", "-s0", "in.iamf"], {"in.iamf": desc + b"".join(units[51:])})
    assert len(want["ss0_in.wav"]) > 44 + 1000 and want["ss0_in.wav"] == got["ss0_in.wav"]


@pytest.mark.gpu
def test_fragmented_mp4_gives_the_bitstream_samples():
    # (the stock player's fragment path needs its own box order; here the fragmented file is checked against the plain bitstream)
    sc, desc, units = stream("c2", F=14)
    a = run(PLAYER, ["-o2", "-s1", "in.iamf"], {"in.iamf": desc + b"".join(units)})
    blob, _ = mp4gen.mux(desc, units, [sc.frame_size] * len(units), fragmented=True, per_fragment=4)
    b = run(PLAYER, ["-i1", "-o2", "-s1", "in.mp4"], {"in.mp4": blob})
    assert len(a["ss1_in.wav"]) > 44 + 1000 and a["ss1_in.wav"] == b["ss1_in.wav"]


@pytest.mark.gpu
def test_several_files_at_once_match_one_by_one():
    # several inputs step together through IAMF_decoder_decode_batch_units (ragged lengths); same files as one at a time
    files, names = {}, []
    for k, F in enumerate((9, 17, 12, 30)):
        sc, desc, units = stream("c2", F=F, seed=100 + k)
        files[f"in{k}.iamf"] = desc + b"".join(units)
        names.append(f"in{k}.iamf")
    together = run(PLAYER, ["-o2", "-s1"] + names, files)
    assert len(together) == 4
    for n in names:
        alone = run(PLAYER, ["-o2", "-s1", n], {n: files[n]})
        w = "ss1_" + n.replace(".iamf", ".wav")
        assert len(alone[w]) > 44 + 1000 and alone[w] == together[w], w
    # inputs of different pipelines fall back to one after the other
    sc, desc, units = stream("c1", F=8)
    files["other.iamf"] = desc + b"".join(units)
    mixed = run(PLAYER, ["-o2", "-s1", "in0.iamf", "other.iamf"], files)
    assert mixed["ss1_in0.wav"] == together["ss1_in0.wav"] and len(mixed["ss1_other.wav"]) > 44 + 1000
