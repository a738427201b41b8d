#!/usr/bin/env python3
"""Generates tests/golden/*.npz: outputs of the UNMODIFIED reference decoder (oracle/_ref/libiamf_ref.so, compiled from
/root/reference by oracle/Makefile) driven end to end through its public API
(IAMF_decoder_open/configure/decode, include/IAMF_decoder.h:60-239) on synthetic ipcm bitstreams of the five
BASELINE.json configurations plus a few setter variants.

The reference tree holds no golden vectors of its own, so these fixtures - produced by running
the reference itself in the authoring container - are what pins both the oracle (tests/test_golden.py, CPU) and the
CUDA path (tests/test_gpu_golden.py) wherever /root/reference and oracle/_ref are absent.

    python tests/golden/make_golden.py          # needs oracle/_ref (make -C oracle ref)

Each fixture stores: the case name, number of streams / frames, the seeds, a SHA-256 of the synthesised decoded input
(so that a change of the generator is caught rather than silently compared), the per-call sample counts returned by
IAMF_decoder_decode (including the final flush call) and the interleaved PCM bytes per stream.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import iamfapi  # noqa: E402
import refbind  # noqa: E402
import refstreams  # noqa: E402
import scenarios as S  # noqa: E402

# (fixture name, refstreams case, scenario kwargs, streams, frames, input seed, parameter seed)
CASES = [
    ("c1_stereo_to_A", "c1", {}, 3, 8, 101, 201),
    ("c1_stereo_to_A_hot_24bit", "c1", dict(bit_depth=24, peak_db=(0.0, 4.0)), 2, 8, 102, 202),
    ("c1_stereo_to_A_nolimiter_32bit", "c1", dict(bit_depth=32, limiter=False, peak_db=(-1.0, 0.0)), 2, 4, 103, 203),
    ("c2_714_scalable_to_B", "c2", {}, 3, 14, 104, 204),
    ("c2_714_scalable_to_B_hot", "c2", dict(peak_db=(-1.0, 3.0)), 2, 8, 105, 205),
    ("c3_toa_to_H", "c3", {}, 2, 5, 106, 206),
    ("c4_714_foa_to_binaural", "c4", {}, 2, 6, 107, 207),
    ("c5_resample_loud_lim", "c5", {}, 3, 9, 108, 208),
    ("c1_trims", "c1", dict(trims={0: (312, 0), 2: (960, 0), 5: (0, 100), 7: (0, 960)}, peak_db=(0.0, 2.0)), 2, 9, 109, 209),
]


def input_digest(inputs, P):
    h = hashlib.sha256()
    for x in inputs:
        h.update(np.ascontiguousarray(x).tobytes())
    h.update(np.ascontiguousarray(P).tobytes())
    return h.hexdigest()


def build_case(case, kw, n_streams, n_frames, seed_in, seed_p):
    sc, st, api_kw, unit_kw = refstreams.case(case, **kw)
    inputs = S.synth_inputs(sc, n_streams, n_frames, seed=seed_in)
    P, ramps, oramp = S.synth_params(sc, n_streams, n_frames, seed=seed_p)
    refstreams.no_param_gaps(sc, P)
    return sc, st, api_kw, unit_kw, inputs, P, ramps, oramp


def main():
    if not refbind.have_ref():
        sys.exit("oracle/_ref/libiamf_ref.so is missing: run `make -C oracle ref` where /root/reference exists")
    api = iamfapi.Api(refbind.REF_SO)
    for name, case, kw, n_streams, n_frames, seed_in, seed_p in CASES:
        sc, st, api_kw, unit_kw, inputs, P, _, _ = build_case(case, kw, n_streams, n_frames, seed_in, seed_p)
        desc = st.descriptors()
        out = dict(case=case, n_streams=n_streams, n_frames=n_frames, seed_in=seed_in, seed_p=seed_p,
                   digest=input_digest(inputs, P), out_channels=sc.out_channels, bit_depth=sc.bit_depth)
        for s in range(n_streams):
            units = refstreams.temporal_units(sc, st, inputs, P, unit_kw, s)
            pcm, counts = api.render(desc, units, **api_kw)
            out[f"counts{s}"] = np.asarray(counts, np.int32)
            out[f"pcm{s}"] = np.frombuffer(pcm.tobytes(), np.uint8)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {n_streams} streams x {n_frames} frames -> {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
