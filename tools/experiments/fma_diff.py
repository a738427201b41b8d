"""experiment: where does IAMFB_ARITH_FMA differ from the exact oracle, and by how much (run on a GPU box)"""
import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenarios as S  # noqa: E402
from gpu_harness import run_product  # noqa: E402

for base, bits in ((S.c3_toa_to_H(), 16), (S.c3_toa_to_H(), 0), (S.c3_toa_to_H(limiter=False), 16), (S.c5_resample(), 16), (S.c5_resample(), 0)):
    sc = dataclasses.replace(base, arithmetic=1, bit_depth=bits)
    n, F = 21, 10
    inputs = S.synth_inputs(sc, n, F, seed=0x1A3F + 300)
    P, ramps, oramp = S.synth_params(sc, n, F, seed=0x77 + 300)
    ref = S.run_oracle(sc, inputs, P, ramps, oramp)
    got, _ = run_product(sc, inputs, P, ramps, oramp, splits=[4, 6])
    dt = np.float32 if bits == 0 else np.int16
    hist = {}
    for s in range(n):
        a, b = got[s][1].view(dt).astype(np.float64), ref[s][1].view(dt).astype(np.float64)
        d = np.abs(a - b)
        if bits == 0:
            k = "max %.3g" % d.max()
            hist[k] = hist.get(k, 0) + 1
        else:
            for v in np.unique(d):
                hist[int(v)] = hist.get(int(v), 0) + int((d == v).sum())
            bad = np.nonzero(d >= 2)[0]
            if len(bad):
                co = sc.out_channels
                print("  stream", s, "first >=2 at sample", bad[0] // co, "ch", bad[0] % co, "values", a[bad[0]], b[bad[0]], "n", len(bad))
    print(sc.name, "limiter", sc.limiter, "bits", bits, hist)
