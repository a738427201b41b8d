"""Writes the seed inputs of the host-layer fuzzer (tests/fuzz/fuzz_host.c): the five BASELINE configurations as
descriptor OBUs + two temporal units each, and two MP4 files (plain and fragmented) of configuration 1."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import mp4gen      # noqa: E402
import refstreams  # noqa: E402
import scenarios as S  # noqa: E402


def main(out):
    os.makedirs(out, exist_ok=True)
    for name in ["c1", "c2", "c3", "c4", "c5"]:
        sc, st, api_kw, unit_kw = refstreams.case(name)
        inputs = S.synth_inputs(sc, 1, 3, seed=3)
        P, _, _ = S.synth_params(sc, 1, 3, seed=4)
        refstreams.no_param_gaps(sc, P)
        desc = st.descriptors()
        tus = refstreams.temporal_units(sc, st, inputs, P, unit_kw, 0)
        with open(os.path.join(out, name + ".bin"), "wb") as f:
            f.write(desc + b"".join(tus[:2]))
        if name == "c1":
            for frag in (False, True):
                blob, _table = mp4gen.mux([desc], tus, [sc.frame_size] * len(tus), timescale=48000, skip=0, fragmented=frag)
                with open(os.path.join(out, "c1_frag.mp4" if frag else "c1.mp4"), "wb") as f:
                    f.write(blob)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/iamfb_fuzz_corpus")
