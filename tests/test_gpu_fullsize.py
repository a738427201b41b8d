"""Full-size launches: the stream counts BASELINE.json names for each configuration, checked through properties that
do not need the oracle to render thousands of streams:

  * K distinct synthetic streams are replicated over the whole batch (stream s carries stream s mod K); every replica
    must produce byte-identical PCM wherever it sits in the batch (no dependence on the position in the grid, no
    interference between neighbours) and the K distinct results must equal the oracle bit for bit;
  * splitting the same frames over two submits must give the same bytes as one submit (state carried between
    submits: de-mixer walk, recon smoothing, resampler phase, limiter delay line and gain state).
"""
import hashlib

import numpy as np
import pytest

import scenarios as S

pytestmark = pytest.mark.gpu

K = 16
FULL = [
    (S.c1_stereo, 1024, 6),               # configs[0] is one stream on the CPU; the batched form of the same pipeline
    (S.c1_stereo, 10656, 3),              # ... at the batch size bench.py runs it with (several waves of k_stream blocks)
    (S.c2_714_to_B, 1024, 6),             # configs[1]
    (S.c3_toa_to_H, 4096, 3),             # configs[2]
    (S.c4_714_foa_binaural, 2048, 4),     # configs[3]
    (S.c5_resample, 2048, 4),             # configs[4]: 16384 streams over 8 GPUs = 2048 per GPU
]


def _tile(a, n):
    reps = (n + a.shape[0] - 1) // a.shape[0]
    return np.concatenate([a] * reps, axis=0)[:n]


@pytest.mark.parametrize("mk,n_streams,F", FULL, ids=[f"{f[0].__name__}-{f[1]}" for f in FULL])
def test_full_size_replicas_and_split_invariance(mk, n_streams, F):
    from gpu_harness import run_product
    sc = mk()
    base_in = S.synth_inputs(sc, K, F, seed=0x1A3F + 77)
    P0, ramps, oramp = S.synth_params(sc, K, F, seed=0x77 + 7)
    assert ramps is None and oramp is None
    inputs = [_tile(x, n_streams) for x in base_in]
    P = _tile(P0, n_streams)
    got, launches = run_product(sc, inputs, P, None, None, splits=[F])
    assert launches > 0
    ref = S.run_oracle(sc, base_in, P0, None, None)
    digests = {}
    for s in range(n_streams):
        k = s % K
        if s < K:
            assert got[s][0] == ref[k][0], f"{sc.name}: stream {s} per-call sample counts"
            assert np.array_equal(got[s][1], ref[k][1]), f"{sc.name}: stream {s} differs from the oracle"
            digests[k] = (got[s][0], hashlib.sha256(got[s][1].tobytes()).digest())
        else:
            assert got[s][0] == digests[k][0], f"{sc.name}: replica {s} of stream {k}: sample counts"
            assert hashlib.sha256(got[s][1].tobytes()).digest() == digests[k][1], f"{sc.name}: replica {s} of stream {k} differs"
    # the same frames over two submits
    a = F // 2
    got2, _ = run_product(sc, inputs, P, None, None, splits=[a, F - a])
    for s in range(n_streams):
        assert sum(got2[s][0]) == sum(got[s][0])
        assert hashlib.sha256(got2[s][1].tobytes()).digest() == digests[s % K][1], f"{sc.name}: stream {s}: split submits differ"
