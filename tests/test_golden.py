"""Known-answer tests against tests/golden/*.npz - PCM produced by the UNMODIFIED reference decoder through its public
API (tests/golden/make_golden.py, run in the authoring container).  They need neither /root/reference nor oracle/_ref:

  * CPU (`-m "not gpu"`): the oracle (oracle/liboracle.so, our C restatement) must reproduce every fixture bit for bit;
  * GPU (`-m gpu`): so must the CUDA path behind the C ABI (include/iamf_b200.h), through host-buffer submits split
    raggedly over several calls.
"""
import importlib.util
import os

import numpy as np
import pytest

import scenarios as S

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
MG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MG)

IDS = [c[0] for c in MG.CASES]


def load(entry):
    name, case, kw, n_streams, n_frames, seed_in, seed_p = entry
    fx = np.load(os.path.join(HERE, "golden", name + ".npz"))
    sc, st, api_kw, unit_kw, inputs, P, ramps, oramp = MG.build_case(case, kw, n_streams, n_frames, seed_in, seed_p)
    assert MG.input_digest(inputs, P) == str(fx["digest"]), "synthetic input generator drifted from the fixture"
    assert int(fx["out_channels"]) == sc.out_channels and int(fx["bit_depth"]) == sc.bit_depth
    return fx, sc, inputs, P, ramps, oramp


def check(fx, got, n_streams, what):
    for s in range(n_streams):
        counts, raw = got[s]
        assert list(counts) == [int(c) for c in fx[f"counts{s}"]], f"{what}: stream {s} per-call sample counts"
        ref = fx[f"pcm{s}"]
        assert raw.shape == ref.shape, f"{what}: stream {s} {raw.shape} vs {ref.shape} bytes"
        assert np.array_equal(raw, ref), f"{what}: stream {s} PCM differs from the reference decoder's output"


@pytest.mark.parametrize("entry", MG.CASES, ids=IDS)
def test_oracle_reproduces_reference_fixture(entry):
    fx, sc, inputs, P, ramps, oramp = load(entry)
    check(fx, S.run_oracle(sc, inputs, P, ramps, oramp), entry[3], "oracle")


@pytest.mark.gpu
@pytest.mark.parametrize("entry", MG.CASES, ids=IDS)
def test_cuda_path_reproduces_reference_fixture(entry):
    from gpu_harness import run_product
    fx, sc, inputs, P, ramps, oramp = load(entry)
    F = entry[4]
    splits = [1, F - 3, 2] if F >= 5 else [F]
    got, launches = run_product(sc, inputs, P, ramps, oramp, splits=splits)
    assert launches > 0
    check(fx, got, entry[3], "cuda")


def test_reference_facts_from_the_survey():
    """HOA -> sound system H leaves channels 3 (LFE1) and 23 at zero"""
    fx = np.load(os.path.join(HERE, "golden", "c3_toa_to_H.npz"))
    pcm = np.frombuffer(fx["pcm0"].tobytes(), np.int16).reshape(-1, 24)
    assert not pcm[:, 3].any() and not pcm[:, 23].any()
    assert pcm[:, 9].any()          # the LFE2 slot carries shifted speaker data (h2m_rdr.c:1010,1114-1135)
