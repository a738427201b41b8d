"""Runs a Scenario through the product's C ABI (host-resident submits) and returns, per stream, the per-call sample
counts and the concatenated PCM bytes in the same form scenarios.run_oracle() returns them."""
import numpy as np

import scenarios as S
from iac_b200 import Engine


def run_product(sc, inputs, P, ramps=None, oramp=None, splits=None, flush=True, device=0, s16=False, expect_path=None, kernels=None):
    n_streams, F = P.shape
    splits = splits or [F]
    assert sum(splits) == F
    eng = Engine(S.plan_desc(sc), n_streams, max(splits), device=device)
    if expect_path is not None:
        path = eng.kernel_path_s16 if s16 else eng.kernel_path
        assert path == expect_path, f"{sc.name}: kernel path {path}, expected {expect_path}"
    if kernels is not None:   # the caller wants to know which kernels ran (name -> launches)
        eng.set_timing(True)
    co = eng.out_channels
    bps = eng.bytes_per_sample
    counts = [[] for _ in range(n_streams)]
    chunks = [[] for _ in range(n_streams)]
    f0 = 0
    for n in splits:
        sl = slice(f0, f0 + n)
        ins = [x[:, sl] for x in inputs]
        if s16:   # hand the int16 the codec produced instead of its float scaling (IAMFB_IN_S16)
            ins = [np.rint(x.astype(np.float64) * 32768.0).astype(np.int16) for x in ins]
        pcm, cnt = eng.submit_host(ins, P[:, sl],
                                   [r[:, sl] if r is not None else None for r in ramps] if ramps else None,
                                   oramp[:, sl] if oramp is not None else None)
        for s in range(n_streams):
            counts[s].extend(int(c) for c in cnt[s])
            tot = int(cnt[s].sum())
            chunks[s].append(pcm[s, : tot * co * bps].copy())
        f0 += n
    if flush:
        pcm, cnt = eng.flush_host()
        for s in range(n_streams):
            counts[s].append(int(cnt[s]))
            chunks[s].append(pcm[s, : int(cnt[s]) * co * bps].copy())
    launches = eng.launch_count()
    if kernels is not None:
        kernels.update({k: int(v[1]) for k, v in eng.get_timing().items()})
    eng.close()
    return {s: (counts[s], np.concatenate(chunks[s]) if chunks[s] else np.zeros(0, np.uint8)) for s in range(n_streams)}, launches
