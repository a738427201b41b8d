#!/usr/bin/env python3
"""bench.py - rendered audio-seconds per second of the post-decode rendering path on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3                 # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # the reference's own CPU implementation

A "step" is one pass of the hot path over one batch of synthetic input: S streams x F consecutive frames of the
configuration BASELINE.json's metric is quoted on (configs[1]: 1024 concurrent 7.1.4-scalable streams with recon-gain
demixing rendered to sound system B).  Streams shard independently over the GPUs (weak scaling, no collective on the
data path; torch.distributed only provides the barrier and the max-over-ranks of the device time).

  value     whole-job audio-s/s with the decoded PCM already resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the host-buffer entry point of the C ABI (pinned host -> H2D -> kernels -> D2H)
  roofline  dominant kernel: algorithmic bytes per launch / its CUDA-event time, vs the measured HBM copy peak
  cpu_baseline  the reference decoder (oracle/_ref, else our C port) on a bounded sample, all host cores

Input sets are larger than L2 (126 MB), so consecutive steps never find their input cached.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

CONFIGS = {
    # (configs[0] is ONE stream on the CPU; the batch is ours to size: 72 streams per SM.  The submit lasts as long as its
    # loudest streams' serial limiter recurrence, so a batch of several waves keeps the SMs busy behind the quiet streams:
    # 1024 x 32 frames 1.21 M audio-s/s, 5328 x 32 2.64 M, 10656 x 16 2.98 M, 21312 x 16 3.20 M - profiles/r2_c1_c2_batch_sizes.log)
    "c1": dict(streams=10656, frames=16, desc="10656 simple-profile stereo streams -> sound system A"),
    "c2": dict(streams=1024, frames=16, desc="1024 base-profile streams, 7.1.4 scalable (2.0 -> 7.1.4) with recon-gain demixing -> sound system B (0+5+0)"),
    "c3": dict(streams=4096, frames=16, desc="4096 streams of 3rd-order ambisonics (16 ch) -> sound system H (9+10+3)"),
    "c4": dict(streams=2048, frames=16, desc="2048 streams, 7.1.4 + FOA mix presentation -> binaural (as built: stereo matrices)"),
    "c4h": dict(streams=2048, frames=16, desc="2048 streams, 7.1.4 + FOA mix presentation -> binaural with HRTF convolution (256-tap in-repo HRIR set, "
                                             "exact int8 tensor-core contraction; self-oracle, parity unpinned by the reference)",
                in_format="s16"),   # the decoded PCM is handed over as the int16 the codecs produce: two limbs per sample instead of three
    "c5": dict(streams=2048, frames=16, desc="2048 streams/GPU, stereo 44.1->48 kHz resample, loudness -24 LKFS, limiter, 16-bit"),
    # the FP32-bound configuration again in the engine's tolerance mode (IAMFB_ARITH_FMA, include/iamf_b200.h): the HOA matrix
    # fuses multiply and add; PCM within +-1 LSB of the reference instead of bit-identical
    # (tests/test_gpu_parity.py::test_fma_arithmetic_stays_within_the_stated_tolerance).  Never the headline.
    "c3f": dict(streams=4096, frames=16, base="c3", arithmetic=1,
                desc="c3 in tolerance mode (IAMFB_ARITH_FMA: fused multiply-add in the HOA matrix, +-1 LSB instead of bit-exact)"),
}


def alg_bytes_per_audio_second(sc):
    """f32 planar decoded input + integer interleaved output, every byte once"""
    bps = sc.bit_depth // 8 if sc.bit_depth else 4
    return sum(el.n_in for el in sc.elements) * 4 * sc.in_rate + sc.out_channels * bps * sc.out_rate


# ---------------------------------------------------------------------------------------------------------------------
# CPU leg: the reference's own implementation on host cores
# ---------------------------------------------------------------------------------------------------------------------
CPU_DISTINCT = 2      # distinct synthetic streams per worker (bounds host memory: ~20 MB per worker on c2)
CPU_RENDERS = 16      # stream renders per worker per step (the distinct streams are cycled)


def _cpu_worker(cfg, kind, n_frames, idx, reps, barrier, q):
    """one worker = one host core.  Set-up (synthesis, bitstream packing) happens once and untimed; every rep renders
    CPU_RENDERS whole streams start to finish through the reference's public API between two barriers."""
    import refbind
    import refstreams
    import scenarios as S
    try:
        sc, st, api_kw, unit_kw = refstreams.case(CONFIGS[cfg].get("base", cfg))
        n = CPU_DISTINCT
        inputs = S.synth_inputs(sc, n, n_frames, seed=0x1A3F + 1000 * idx)
        P, ramps, oramp = S.synth_params(sc, n, n_frames, seed=0x77 + idx)
        refstreams.no_param_gaps(sc, P)
        bps = sc.bit_depth // 8 if sc.bit_depth else 4
        if kind == "reference":
            import iamfapi
            api = iamfapi.Api(refbind.REF_SO)
            desc = st.descriptors()
            units = [refstreams.temporal_units(sc, st, inputs, P, unit_kw, s) for s in range(n)]
        out = []
        for _ in range(reps):
            samples = 0
            barrier.wait()
            t0 = time.monotonic()
            for r in range(CPU_RENDERS):
                s = r % n
                if kind == "reference":
                    pcm, counts = api.render(desc, units[s], **api_kw)
                    samples += pcm.shape[0]
                else:
                    res = S.run_oracle(sc, [x[s:s + 1] for x in inputs], P[s:s + 1],
                                       [g[s:s + 1] for g in ramps] if ramps else None,
                                       oramp[s:s + 1] if oramp is not None else None)
                    samples += len(res[0][1]) // (bps * sc.out_channels)
            t1 = time.monotonic()
            out.append((t0, t1, samples / float(sc.out_rate)))
        q.put((idx, out))
    except Exception as e:  # noqa: BLE001
        try:
            barrier.abort()
        except Exception:
            pass
        q.put((idx, repr(e)))


def cpu_reference_c(cfg, n_frames, reps=1, warm=0, distinct=8):
    """the unmodified reference through its public API from the C harness oracle/ref_harness.c (one worker process per host
    core, every worker renders whole streams start to finish) - no Python in the timed region"""
    import ctypes as C
    import refbind
    import refstreams
    import scenarios as S
    so = os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")
    H = C.CDLL(so, mode=os.RTLD_LOCAL | os.RTLD_NOW)

    class RefStream(C.Structure):
        _fields_ = [("desc", C.c_char_p), ("desc_len", C.c_int), ("n_units", C.c_int), ("units", C.POINTER(C.c_char_p)),
                    ("unit_len", C.POINTER(C.c_int))]

    class RefJob(C.Structure):
        _fields_ = [("n_streams", C.c_int), ("streams", C.POINTER(RefStream)), ("renders", C.c_int), ("threads", C.c_int),
                    ("sound_system", C.c_int), ("bit_depth", C.c_int), ("rate", C.c_int), ("limiter", C.c_int),
                    ("loudness", C.c_float), ("threshold_db", C.c_float), ("out_channels", C.c_int)]
    H.ref_harness_run.argtypes = [C.POINTER(RefJob), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    sc, st, api_kw, unit_kw = refstreams.case(CONFIGS[cfg].get("base", cfg))
    inputs = S.synth_inputs(sc, distinct, n_frames, seed=0x1A3F)
    P, _, _ = S.synth_params(sc, distinct, n_frames, seed=0x77)
    refstreams.no_param_gaps(sc, P)
    desc = st.descriptors()
    keep, streams = [], (RefStream * distinct)()
    for s in range(distinct):
        units = refstreams.temporal_units(sc, st, inputs, P, unit_kw, s)
        ua = (C.c_char_p * len(units))(*units)
        ul = (C.c_int * len(units))(*[len(u) for u in units])
        keep += [units, ua, ul]
        streams[s] = RefStream(desc, len(desc), len(units), ua, ul)
    workers = max(1, len(os.sched_getaffinity(0)))
    job = RefJob(distinct, streams, workers * CPU_RENDERS, workers, -1 if api_kw.get("binaural") else api_kw.get("sound_system", 0),
                 api_kw.get("bit_depth", 16), api_kw.get("rate", 0), 1 if api_kw.get("limiter", True) else 0,
                 api_kw.get("loudness", 0.0), api_kw.get("threshold_db", -1.0), sc.out_channels)
    vals, secs = [], []
    for k in range(warm + reps):
        sec, smp = C.c_double(0), C.c_longlong(0)
        rc = H.ref_harness_run(C.byref(job), C.byref(sec), C.byref(smp))
        if rc:
            raise RuntimeError("reference harness: a stream failed to render")
        if k >= warm:
            vals.append(smp.value / float(sc.out_rate) / sec.value)
            secs.append(sec.value)
    what = (f"{workers} worker processes (one per host core) x {CPU_RENDERS} streams x {n_frames} frames of the same workload per step, "
            "unmodified reference decoder (oracle/_ref/libiamf_ref.so) through IAMF_decoder_configure/decode on ipcm-coded "
            "streams, driven from C (oracle/ref_harness.c)")
    return dict(value=float(np.mean(vals)), unit="audio-s/s", cores=workers, kind="reference", sample=what,
                seconds=float(np.mean(secs)), per_rep=vals)


def cpu_reference(cfg, n_frames, reps=1, warm=0):
    """times the reference CPU implementation of the path on ALL host cores: one process per core, each rendering
    CPU_RENDERS streams x n_frames frames per rep.  Returns dict(value audio-s/s (mean over reps), per_rep, cores,
    kind, sample, seconds)."""
    import multiprocessing as mp
    import refbind
    kind = "reference" if (refbind.have_ref() and cfg != "c4h") else "port"   # (the reference is built without its binauraliser)
    if kind == "reference" and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_harness.so")):
        return cpu_reference_c(cfg, n_frames, reps, warm)
    import orcbind
    orcbind.lib()   # make sure liboracle.so exists before the workers start (refstreams uses its scalar helpers)
    workers = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(workers)
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(cfg, kind, n_frames, i, warm + reps, barrier, q), daemon=True)
             for i in range(workers)]
    for p in procs:
        p.start()
    res = [q.get() for _ in range(workers)]
    for p in procs:
        p.join(timeout=30)
    bad = [r for r in res if isinstance(r[1], str)]
    if bad:
        raise RuntimeError(f"cpu reference worker failed: {bad[0][1]}")
    vals, secs = [], []
    for k in range(warm, warm + reps):
        t0 = min(r[1][k][0] for r in res)
        t1 = max(r[1][k][1] for r in res)
        audio = sum(r[1][k][2] for r in res)
        vals.append(audio / (t1 - t0))
        secs.append(t1 - t0)
    what = (f"{workers} processes (one per host core) x {CPU_RENDERS} streams x {n_frames} frames of the same workload per step, "
            + ("unmodified reference decoder (oracle/_ref) through IAMF_decoder_configure/decode on ipcm-coded streams"
               if kind == "reference" else "C port of the path (oracle/)"))
    return dict(value=float(np.mean(vals)), unit="audio-s/s", cores=workers, kind=kind, sample=what,
                seconds=float(np.mean(secs)), per_rep=vals)


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.marks = []

    def mark(self):
        """brackets the timed region: samples taken between the first two marks are reported apart"""
        self.marks.append(len(self.samples))

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 8 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 8 and s[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            if len(s) >= 8:
                for n, v in zip(names, s[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        timed = None
        if len(self.marks) >= 2:
            tm = [float(x[0]) for x in self.samples[self.marks[0]:self.marks[1]] if len(x) >= 8 and x[0].replace(".", "").isdigit()]
            timed = dict(sm_mhz=float(np.median(tm)) if tm else None, samples=len(tm))
        return dict(sm_mhz=(timed["sm_mhz"] if timed and timed["sm_mhz"] else (float(np.median(sm)) if sm else None)),
                    sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(sm),
                    timed_region=timed, whole_run_sm_mhz=float(np.median(sm)) if sm else None)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def make_inputs_torch(sc, S, F, seed, device, render_peak=None):
    """device-resident decoded PCM: 3 sines + noise per channel, per-stream peak uniformly in sc.peak_db dBFS,
    quantised to int16 and scaled by 1/32768 (the codec glue's contract).  [S][F][C][N] float32 per element.

    The level is the peak the LIMITER sees ("threshold is -1 dBFS, so ~20 % of streams exercise the limiter's
    active path"): render_peak(inputs) -> [S] returns the pre-limiter peak of every stream's rendered mix for a trial
    input, and - everything before the limiter being linear - the channels are rescaled so that this peak lands on the
    stream's target.  render_peak=None scales the decoded channels themselves to the target instead (--peak-ref input:
    after a 12 -> 6 channel down-mix that puts ~80 % of the tiles above the threshold).
    Returns (inputs, per-stream target peak)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    N = sc.frame_size
    T = F * N
    t = torch.arange(T, device=device, dtype=torch.float32) / sc.in_rate
    outs = []
    peak = 10.0 ** ((sc.peak_db[0] + (sc.peak_db[1] - sc.peak_db[0]) * torch.rand(S, generator=g, device=device)) / 20.0)
    trial = 0.125 if render_peak is not None else None      # trial level of the calibration pass (no clipping anywhere)
    for el in sc.elements:
        C = el.n_in
        x = torch.empty((S, C, T), device=device, dtype=torch.float32)
        step = max(1, 256 // C)
        for s0 in range(0, S, step):
            s1 = min(S, s0 + step)
            n = s1 - s0
            fr = 50.0 + 11950.0 * torch.rand((n, C, 3, 1), generator=g, device=device)
            ph = 6.2831853 * torch.rand((n, C, 3, 1), generator=g, device=device)
            y = torch.sin(6.2831853 * fr * t.view(1, 1, 1, T) + ph).sum(dim=2)
            y += 0.6 * torch.rand((n, C, T), generator=g, device=device) - 0.3
            if trial is None:
                y *= (peak[s0:s1] / y.abs().amax(dim=(1, 2))).view(n, 1, 1)
                x[s0:s1] = torch.clamp(torch.round(y * 32768.0), -32768, 32767) / 32768.0
            else:
                x[s0:s1] = y * (trial / y.abs().amax(dim=(1, 2))).view(n, 1, 1)
        outs.append(x.view(S, C, F, N).permute(0, 2, 1, 3).contiguous())
    if trial is not None:
        rendered = render_peak(outs).clamp_min(1e-9)        # [S] pre-limiter peak of the mix at the trial level
        for x in outs:
            x *= (peak / rendered).view(S, 1, 1, 1)
            torch.clamp(torch.round(x * 32768.0), -32768, 32767, out=x)
            x /= 32768.0
    return outs, peak


def rendered_peak_fn(sc, S_, F, P, local, stream, dev):
    """pre-limiter peak of every stream's rendered mix: the same plan with the limiter off and the float debug output
    (bit_depth 0), run once through the engine (untimed, before the benchmark's own engine exists)"""
    import dataclasses
    import torch
    import scenarios as S
    from iac_b200 import Engine

    def fn(inputs):
        cal = dataclasses.replace(sc, limiter=False, bit_depth=0)
        eng = Engine(S.plan_desc(cal), S_, F, device=local, cuda_stream=stream.cuda_stream)
        d_params = torch.from_numpy(P.view(np.uint8).reshape(S_, F * 48).copy()).to(dev)
        stride = eng.out_stride_bytes(F)
        d_pcm = torch.zeros((S_, stride), dtype=torch.uint8, device=dev)
        d_counts = torch.zeros((S_, F), dtype=torch.int32, device=dev)
        eng.submit_device([x.data_ptr() for x in inputs], d_params.data_ptr(), d_pcm.data_ptr(), d_counts.data_ptr(), F)
        torch.cuda.synchronize()
        n = d_counts.sum(dim=1).to(torch.int64) * cal.out_channels          # floats written per stream
        pcm = d_pcm.view(torch.float32).view(S_, -1)
        mask = torch.arange(pcm.shape[1], device=dev).view(1, -1) < n.view(-1, 1)
        pk = torch.where(mask, pcm.abs(), torch.zeros_like(pcm)).amax(dim=1)
        eng.close()
        return pk
    return fn


class DeviceWorkload:
    """one configuration resident in HBM: synthetic decoded PCM, per-frame parameters, an engine and its output buffers"""

    def __init__(self, cfg, S_, F, rank, local, dev, stream, peak_ref="rendered", peak_db="", in_format="f32"):
        import torch
        import scenarios as S
        import refstreams
        from iac_b200 import Engine
        self.cfg, self.S, self.F, self.dev, self.stream = cfg, S_, F, dev, stream
        sc, _, _, _ = refstreams.case(CONFIGS[cfg].get("base", cfg))
        sc.arithmetic = CONFIGS[cfg].get("arithmetic", 0)
        if peak_db:   # experiment knob (not the benchmark workload): per-stream peak range in dBFS
            sc.peak_db = tuple(float(v) for v in peak_db.split(","))
        self.sc = sc
        P, _, _ = S.synth_params(sc, S_, F, seed=0x77 + rank)
        refstreams.no_param_gaps(sc, P)
        self.P = P
        cal = rendered_peak_fn(sc, S_, F, P, local, stream, dev) if (peak_ref == "rendered" and sc.limiter) else None
        self.calibrated = cal is not None
        inputs, target_peak = make_inputs_torch(sc, S_, F, seed=0x1A3F + 7919 * rank, device=dev, render_peak=cal)
        thr_lin = 10.0 ** (sc.threshold_db / 20.0)
        self.active_frac = float((target_peak > thr_lin).float().mean().item()) if cal is not None else None
        self.inputs_f32 = inputs
        self.in_format = in_format
        if in_format == "s16":
            # what core decode hands over BEFORE the codec glue's 1/32768 (opus/IAMF_opus_decoder.c:133-135): int16
            inputs = [torch.round(x * 32768.0).to(torch.int16).contiguous() for x in inputs]
        self.inputs = inputs
        self.eng = Engine(S.plan_desc(sc), S_, F, device=local, cuda_stream=stream.cuda_stream)
        self.d_params = torch.from_numpy(P.view(np.uint8).reshape(S_, F * 48).copy()).to(dev)
        self.stride = self.eng.out_stride_bytes(F)
        self.d_pcm = torch.zeros((S_, self.stride), dtype=torch.uint8, device=dev)
        self.d_counts = torch.zeros((S_, F), dtype=torch.int32, device=dev)
        self.in_ptrs = [x.data_ptr() for x in self.inputs]
        self.input_bytes = int(sum(x.numel() * x.element_size() for x in self.inputs))

    def submit(self):
        self.eng.submit_device(self.in_ptrs, self.d_params.data_ptr(), self.d_pcm.data_ptr(), self.d_counts.data_ptr(), self.F,
                               in_format=1 if self.in_format == "s16" else 0)

    def out_per_submit(self):
        return int(self.d_counts.sum().item())

    def kernel_timing(self, submits):
        """per-kernel CUDA-event timing (a separate pass, so that the events do not perturb the throughput number)"""
        import torch
        self.eng.set_timing(True)
        for _ in range(submits):
            self.submit()
        torch.cuda.synchronize()
        timing = self.eng.get_timing()
        self.eng.set_timing(False)
        # per launch: the MEDIAN of the launches for a kernel launched once per submit (a launch that was pre-empted or met a
        # clock dip does not move it); the mean for kernels with several, differently sized launches per submit
        out = {}
        for k, v in timing.items():
            once = v[1] <= submits
            per = v[2] if once else v[0] / max(v[1], 1)
            out[k] = dict(ms_per_launch=per, launches_per_submit=v[1] / submits, ms_per_submit=per * v[1] / submits,
                          mean_ms_per_launch=v[0] / max(v[1], 1), statistic="median" if once else "mean")
        return out

    def roofline(self, kernels, ms_per_submit, out_per_submit):
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        dom = max(kernels, key=lambda k: kernels[k]["ms_per_submit"])
        alg = alg_bytes_per_audio_second(self.sc) * (out_per_submit / self.sc.out_rate)
        dom_launches = max(kernels[dom]["launches_per_submit"], 1e-9)
        achieved = alg / dom_launches / (kernels[dom]["ms_per_launch"] / 1e3) / 1e9
        traffic = None   # DRAM traffic of the dominant kernel from the committed ncu capture (same workload shape), else null
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(self.cfg, {}).get(dom)
            if t and t["streams"] == self.S and t["frames"] == self.F:
                traffic = t["bytes_per_launch"]
        except Exception:
            pass
        pipe = alg / (ms_per_submit / 1e3) / 1e9
        if dom == "k_hrtf_gemm":
            # the HRTF contraction is tensor-bound: int8 MACs ISSUED (every limb pair of every tile, Toeplitz padding included)
            # against twice the measured dense bf16 rate (kind::i8 runs at twice the 16-bit rate); "useful" counts one
            # multiply-add per sample, tap and ear
            audio = out_per_submit / self.sc.out_rate
            ch = sum(el.n_in for el in self.sc.elements if getattr(el, "hrtf", False))
            useful = audio * self.sc.in_rate * ch * 256 * 2
            limb_pairs = 2 * (2 if self.in_format == "s16" else 3)
            issued = useful * limb_pairs * (320.0 / 256.0)
            t = kernels[dom]["ms_per_submit"] / 1e3
            peak_t = 2.0 * float(peaks.get("bf16_tflops", 2250.0))
            return dict(bound="tensor", kernel=dom, achieved=2 * issued / t / 1e12, peak=peak_t, unit="int8 TOP/s", frac=2 * issued / t / 1e12 / peak_t,
                        traffic=traffic, peak_source="2 x MEASURED_PEAKS.json bf16_tflops (int8 dense rate)" if "bf16_tflops" in peaks else "2 x 2250 nominal",
                        useful_fir_tmacs=useful / t / 1e12, limb_pairs=limb_pairs,
                        pipeline=dict(achieved=pipe, frac=pipe / peak_gbs, note="all kernels of a submit, algorithmic bytes / submit time, vs HBM peak"),
                        kernels=kernels)
        ro = dict(bound="hbm", kernel=dom, achieved=achieved, peak=peak_gbs, unit="GB/s", frac=achieved / peak_gbs,
                  traffic=traffic, peak_source=peak_src, algorithmic_bytes_per_launch=alg / dom_launches,
                  pipeline=dict(achieved=pipe, frac=pipe / peak_gbs, note="all kernels of a submit, algorithmic bytes / submit time"),
                  kernels=kernels)
        if dom in ("k_resample_ls", "k_pipe_rs") and self.sc.in_rate != self.sc.out_rate:
            # the resampler FIR is bound by the FP32 pipe, not by HBM (the contract's roofline above stays the HBM one): the
            # reference's sums are 4 accumulators x filt_len taps per output and channel, every product rounded before it is
            # added (2 lane-operations per tap), against 128 FP32 lanes per SM at the device's clock
            import torch
            pr = torch.cuda.get_device_properties(self.dev)
            nf = 64 if self.sc.out_rate >= self.sc.in_rate else (((64 * self.sc.in_rate // self.sc.out_rate) - 1) & ~7) + 8
            ops = out_per_submit * self.sc.out_channels * 4 * nf * 2
            peak_ops = pr.multi_processor_count * 128 * float(getattr(pr, "clock_rate", 1965000)) * 1e3
            t = kernels[dom]["ms_per_submit"] / 1e3
            ro["fp32_pipe"] = dict(lane_ops_per_submit=ops, peak_lane_ops_per_s=peak_ops, floor_ms=ops / peak_ops * 1e3,
                                   frac=ops / t / peak_ops, note="exact (separately rounded) multiply-adds of the resampler FIR vs 128 FP32 lanes per SM")
        return ro

    def close(self):
        self.eng.close()


def timed_submits(w, n_submits, barrier):
    """device time of n_submits back-to-back submits (CUDA events on the launching stream)"""
    import torch
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(w.stream)
    for _ in range(n_submits):
        w.submit()
    e1.record(w.stream)
    barrier()
    return e0.elapsed_time(e1)


def copy_ceiling(h2d_bytes, d2h_bytes, dev, barrier, reps=5):
    """the same bytes as plain pinned copies, both directions at once on two streams (what the box's host path can move).
    Every rep starts behind a barrier, so that under torchrun all ranks copy at the same time; the median rep counts."""
    import torch
    up_h = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    dn_h = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    up_d = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    dn_d = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    times = []
    for k in range(reps + 1):
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_up):
            up_d.copy_(up_h, non_blocking=True)
        with torch.cuda.stream(s_dn):
            dn_h.copy_(dn_d, non_blocking=True)
        torch.cuda.synchronize()
        if k:
            times.append(time.perf_counter() - t0)
    return float(np.median(times))


def api_leg(cfg, n_handles, K, rank, w, shard, dev, distinct=32, calls=6):
    """audio-s/s through IAMF_decoder_decode_batch_units: n_handles handles, K temporal units per handle and call"""
    import ctypes as C
    import iamfapi
    import refstreams
    import scenarios as S
    sc, st, api_kw, unit_kw = refstreams.case(CONFIGS[cfg].get("base", cfg))
    inputs = S.synth_inputs(sc, distinct, K, seed=0x1A3F + 977 * rank)
    P, _, _ = S.synth_params(sc, distinct, K, seed=0x99 + rank)
    refstreams.no_param_gaps(sc, P)
    desc = st.descriptors()
    blobs = [b"".join(refstreams.temporal_units(sc, st, inputs, P, unit_kw, s)) for s in range(distinct)]
    first = [refstreams.temporal_units(sc, st, inputs, P[:, :1], unit_kw, s)[0] for s in range(distinct)]
    api = iamfapi.Api(os.path.join(ROOT, "iac_b200", "libiamf.so"))
    L = api.L
    vp = C.c_void_p
    L.IAMF_decoder_decode_batch_units.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int32),
                                                  C.POINTER(C.c_uint32), C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
    os.environ["IAMF_B200_DEVICE"] = str(dev.index or 0)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if "IAMF_B200_HOST_THREADS" not in os.environ:    # the ranks of a box share its host cores
        os.environ["IAMF_B200_HOST_THREADS"] = str(max(2, len(os.sched_getaffinity(0)) // world))
    n = n_handles
    hs = (vp * n)()
    for i in range(n):
        hs[i] = api.open_configured(desc + first[i % distinct], **api_kw)
    ch = sc.out_channels
    bps = sc.bit_depth // 8 if sc.bit_depth else 4
    per = bps * ch * (sc.frame_size * (sc.out_rate // sc.in_rate + 1) + 64) * K
    big = C.create_string_buffer(per * n)
    base = C.addressof(big)
    pcm = (vp * n)(*[base + i * per for i in range(n)])
    data = (C.c_char_p * n)(*[blobs[i % distinct] for i in range(n)])
    size = (C.c_int32 * n)(*[len(blobs[i % distinct]) for i in range(n)])
    rs = (C.c_uint32 * n)()
    ret = (C.c_int * n)()
    in_bytes = sum(len(blobs[i % distinct]) for i in range(n))

    def step():
        rc = L.IAMF_decoder_decode_batch_units(hs, n, data, size, rs, pcm, ret, K, None)
        assert rc == 0, rc
    for _ in range(2):
        step()
    t0 = time.perf_counter()
    for _ in range(calls):
        step()
    dt = time.perf_counter() - t0
    out = float(sum(ret[i] for i in range(n)))
    assert all(rs[i] == size[i] for i in range(n))
    for i in range(n):
        L.IAMF_decoder_close(hs[i])
    ms_max, out_total = shard.aggregate(dt * 1e3, out, device=dev)
    value = shard.job_throughput(ms_max, out_total, sc.out_rate, steps=calls)
    return dict(value=value, unit="audio-s/s", ms_per_step=ms_max / calls, handles_per_gpu=n, units_per_call=K,
                bitstream_bytes_per_step=in_bytes, host_threads=int(os.environ.get("IAMF_B200_HOST_THREADS", "0")) or len(os.sched_getaffinity(0)),
                entry="IAMF_decoder_decode_batch_units (include/IAMF_decoder.h): ipcm-coded temporal units in, interleaved int16 PCM out")


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import scenarios as S

    from iac_b200 import Engine, shard
    rank, world, local = shard.rank_world()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = args.config
    S_, F = args.streams or CONFIGS[cfg]["streams"], args.frames or CONFIGS[cfg]["frames"]

    # CPU baseline first (rank 0, N=1 only), before the GPU is busy
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        cpu = cpu_reference(cfg, n_frames=args.cpu_frames, reps=1, warm=1)
        cpu.pop("per_rep", None)

    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.in_format == "f32" and "in_format" in CONFIGS[cfg]:
        args.in_format = CONFIGS[cfg]["in_format"]
    w = DeviceWorkload(cfg, S_, F, rank, local, dev, stream, args.peak_ref, args.peak_db, args.in_format)
    sc = w.sc

    # clocks / throttle reasons are sampled from the warm-up to the end of the end-to-end leg
    sampler = ClockSampler(local)
    sampler.start()
    # A step = R consecutive submits of the same S x F batch (the streams simply go on: limiter / de-mixer state carries
    # over), R chosen so that one step lasts ~args.step_ms and the timed region of K steps is >= 1 s: the clocks the
    # number was taken at are sustained ones, not a 5 ms burst
    for _ in range(3):
        w.submit()
    probe = timed_submits(w, 8, barrier) / 8.0
    R = args.submits_per_step or max(1, int(round(args.step_ms / max(probe, 1e-3))))
    for _ in range(args.warmup):
        for _ in range(R):
            w.submit()
    l0 = w.eng.launch_count()
    sampler.mark()
    ms = timed_submits(w, args.steps * R, barrier)
    sampler.mark()
    launches = w.eng.launch_count() - l0
    out_per_submit = w.out_per_submit()                 # samples produced by one steady-state submit (this rank)
    ms_max, out_total = shard.aggregate(ms, out_per_submit * R, device=dev)     # MAX of device time, SUM of samples
    value = shard.job_throughput(ms_max, out_total, sc.out_rate, steps=args.steps)
    ms_per_submit = ms_max / (args.steps * R)

    if args.quick:
        sampler.stop_flag = True
        if rank == 0:
            extra = {}
            if os.environ.get("IAMFB_BENCH_KERNELS"):      # per-kernel CUDA-event times of the quick run (development aid)
                extra["kernels"] = {k: round(v["ms_per_submit"], 5) for k, v in w.kernel_timing(40).items()}
            print(json.dumps({**extra, "metric": "rendered audio-sec/sec", "value": value, "ms_per_step": ms_max / args.steps,
                              "ms_per_submit": ms_per_submit, "submits_per_step": R,
                              "gpu_launches": int(launches), "quick": True, "peak_ref": args.peak_ref,
                              "streams_above_limiter_threshold": w.active_frac, "kernel_path": w.eng.kernel_path_s16 if w.in_format == "s16" else w.eng.kernel_path}))
        w.close()
        return

    # per-kernel times come from a separate pass (events around every launch).  A kernel cannot take longer than the whole
    # submit it is part of took in the sustained region above: a pass whose dominant kernel says otherwise met a clock dip
    # (seen once: 0.265 ms against a 0.229 ms submit) and is repeated, at most twice; the passes taken are reported
    passes = []
    for _ in range(3):
        kernels = w.kernel_timing(min(args.steps * R, 120))
        dom_ms = max(v["ms_per_submit"] for v in kernels.values())
        passes.append(round(dom_ms, 6))
        if dom_ms <= ms_per_submit * 1.02:
            break
    roofline = w.roofline(kernels, ms_per_submit, out_per_submit)
    roofline["timing_passes_dominant_ms"] = passes

    # ---- end to end through the host-buffer entry point of the C ABI: pinned host memory -> H2D -> kernels -> D2H every
    # step.  The host buffers hold what core decode produces for this workload - int16 PCM (Opus / AAC / 16-bit ipcm),
    # IAMFB_IN_S16 - and the interleaved int16 PCM comes back; the call pipelines groups of streams over the two copy
    # engines and the compute stream.
    Fe = min(F, args.e2e_frames)
    eng_e = Engine(S.plan_desc(sc), S_, Fe, device=local, cuda_stream=stream.cuda_stream)
    h_in = [torch.round(x[:, :Fe] * 32768.0).to(torch.int16).contiguous().cpu().pin_memory() for x in w.inputs_f32]
    h_params = torch.from_numpy(np.ascontiguousarray(w.P[:, :Fe]).view(np.uint8).reshape(S_, Fe * 48).copy()).pin_memory()
    stride_e = eng_e.out_stride_bytes(Fe)
    h_pcm = torch.zeros((S_, stride_e), dtype=torch.uint8).pin_memory()
    h_counts = torch.zeros((S_, Fe), dtype=torch.int32).pin_memory()
    import ctypes as C
    from iac_b200.binding import Io, _check
    io = Io()
    for e, x in enumerate(h_in):
        io.in_[e] = x.data_ptr()
    io.in_format = 1
    io.params = h_params.data_ptr()
    io.pcm = h_pcm.data_ptr()
    io.out_counts = h_counts.data_ptr()

    def step_e2e():
        _check(eng_e.L.iamfb_batch_submit_host(eng_e.batch, C.byref(io), Fe), "iamfb_batch_submit_host")

    for _ in range(3):
        step_e2e()
    barrier()
    ke = max(3, min(args.steps, 10))
    le0 = eng_e.launch_count()
    t0 = time.perf_counter()
    for _ in range(ke):
        step_e2e()          # synchronous: returns when the PCM of every stream is back in host memory
    te = time.perf_counter() - t0
    e2e_launches = eng_e.launch_count() - le0
    out_e = float(h_counts.sum().item())
    te_max_ms, oe_total = shard.aggregate(te * 1e3, out_e, device=dev)
    e2e_value = shard.job_throughput(te_max_ms, oe_total, sc.out_rate, steps=ke)
    h2d = sum(x.numel() * 2 for x in h_in) + h_params.numel()
    d2h = S_ * stride_e + h_counts.numel() * 4
    eng_e.close()
    # the same bytes as two plain pinned copies running against each other on this rank (all ranks at once under
    # torchrun): the ceiling the host path of this box sets for the end-to-end step
    barrier()
    ceil_s = copy_ceiling(int(h2d), int(d2h), dev, barrier)
    ceil_ms_max, _ = shard.aggregate(ceil_s * 1e3, 0.0, device=dev)
    del h_in, h_pcm

    # ---- the same through the PUBLIC API of the drop-in library (include/IAMF_decoder.h + the additive batch call): one
    # IAMF_DecoderHandle per stream fed ipcm-coded IAMF temporal units (OBU parsing, parameter time lines and core decode on
    # the host thread pool, int16 hand-over, one device pass per call for the whole group)
    e2e_api = None
    if not args.no_api_leg:
        try:
            e2e_api = api_leg(cfg, S_, min(Fe, 8), rank, w, shard, dev)
        except Exception as ex:  # noqa: BLE001
            e2e_api = {"error": repr(ex)}

    # ---- the other BASELINE configurations and the input-referred level of this one, device-resident, ~0.3 s each
    others = {}
    value_input_referred = None
    if not args.no_other_configs:
        w_keep = w
        if sc.limiter and args.peak_ref == "rendered":
            w2 = DeviceWorkload(cfg, S_, F, rank, local, dev, stream, "input", "", args.in_format)
            for _ in range(3):
                w2.submit()
            n2 = max(4, int(300.0 / max(ms_per_submit, 1e-3)))
            t2 = timed_submits(w2, n2, barrier)
            t2_max, o2 = shard.aggregate(t2, w2.out_per_submit() * n2, device=dev)
            value_input_referred = dict(value=shard.job_throughput(t2_max, o2, sc.out_rate, steps=1),
                                        ms_per_submit=t2_max / n2,
                                        note="per-stream peak applied to the decoded input channels instead of the rendered mix")
            w2.close()
            del w2
        for oc in sorted(CONFIGS):
            if oc == cfg:
                continue
            torch.cuda.empty_cache()
            wo = DeviceWorkload(oc, CONFIGS[oc]["streams"], CONFIGS[oc]["frames"], rank, local, dev, stream, args.peak_ref, "",
                                CONFIGS[oc].get("in_format", args.in_format))
            for _ in range(3):
                wo.submit()
            pr = timed_submits(wo, 4, barrier) / 4.0
            no = max(4, int(300.0 / max(pr, 1e-3)))
            to = timed_submits(wo, no, barrier)
            ops = wo.out_per_submit()
            to_max, oo = shard.aggregate(to, ops * no, device=dev)
            ko = wo.kernel_timing(min(no, 20))
            ro = wo.roofline(ko, to_max / no, ops)
            others[oc] = dict(value=shard.job_throughput(to_max, oo, wo.sc.out_rate, steps=1), unit="audio-s/s",
                              ms_per_submit=to_max / no, submits_timed=no, streams_per_gpu=wo.S, frames_per_submit=wo.F,
                              workload=CONFIGS[oc]["desc"], kernel_path=wo.eng.kernel_path_s16 if wo.in_format == "s16" else wo.eng.kernel_path, input_format=wo.in_format,
                              streams_above_limiter_threshold=wo.active_frac,
                              roofline=dict(bound=ro["bound"], kernel=ro["kernel"], achieved=ro["achieved"], peak=ro["peak"], unit=ro["unit"],
                                            frac=ro["frac"], traffic=ro["traffic"], pipeline_frac=ro["pipeline"]["frac"],
                                            kernels={k: round(v["ms_per_submit"], 5) for k, v in ko.items()},
                                            **({"fp32_pipe": ro["fp32_pipe"]} if "fp32_pipe" in ro else {})))
            wo.close()
            del wo
        w = w_keep

    sampler.stop_flag = True
    sampler.join(timeout=2)
    if rank == 0:
        clocks = sampler.summary()
        line = {
            "metric": "rendered audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{cfg}: {CONFIGS[cfg]['desc']}", "streams_per_gpu": S_, "frames_per_submit": F,
                       "submits_per_step": R, "ms_per_submit": ms_per_submit,
                       "frame_size": sc.frame_size, "input_bytes_per_submit_per_gpu": w.input_bytes,
                       "input_format": ("int16 PCM as core decode produces it, scaled by 1/32768 on the device (IAMFB_IN_S16)"
                                        if args.in_format == "s16" else "float32 (after the codec glue's 1/32768)"),
                       "l2_policy": "the inputs of one submit are larger than L2 (126 MB); nothing is re-read from cache across submits",
                       "limiter_active_peak_range_db": list(sc.peak_db),
                       "peak_reference": ("pre-limiter peak of the rendered mix" if w.calibrated else "decoded input channels"),
                       "streams_above_limiter_threshold": w.active_frac, "kernel_path": w.eng.kernel_path_s16 if w.in_format == "s16" else w.eng.kernel_path},
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "frames_per_step": Fe, "steps": ke, "ms_per_step": te_max_ms / ke, "input": "int16 PCM as decoded (IAMFB_IN_S16)",
                    "gpu_launches": int(e2e_launches),
                    "copy_ceiling_ms": ceil_ms_max, "frac_of_copy_ceiling": ceil_ms_max / (te_max_ms / ke),
                    "copy_ceiling_note": "the step's H2D and D2H bytes as two plain pinned copies running against each other, all ranks at once behind a barrier (median of 5), max over ranks"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "e2e_api": e2e_api,
            "value_input_referred": value_input_referred,
            "configs": others,
        }
        print(json.dumps(line))
    w.close()
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """the reference's own CPU implementation of the path on all host cores (oracle/_ref when it was compiled,
    else our C port); rank 0 only under torchrun"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    res = cpu_reference(cfg, n_frames=args.cpu_frames, reps=max(1, args.steps), warm=max(0, args.warmup))
    value = res["value"]
    line = {
        "impl": "reference", "metric": "rendered audio-sec/sec", "value": value, "unit": "audio-s/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{cfg}: {CONFIGS[cfg]['desc']}", "sample_streams_per_step": res["cores"] * CPU_RENDERS,
                   "sample_frames": args.cpu_frames},
        "cpu_baseline": dict(value=value, unit="audio-s/s", cores=res["cores"], kind=res["kind"], sample=res["sample"]),
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    # rank 0 prints exactly ONE line on stdout: NCCL's version banner (printed at NCCL_DEBUG=VERSION and above) and
    # anything else it logs go to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--e2e-frames", type=int, default=8)
    ap.add_argument("--cpu-frames", type=int, default=250)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-resident timing only (used under ncu)")
    ap.add_argument("--step-ms", type=float, default=60.0, help="target duration of one step (a step = R back-to-back submits)")
    ap.add_argument("--submits-per-step", type=int, default=0, help="R; 0 = derive it from --step-ms")
    ap.add_argument("--in-format", default="f32", choices=["f32", "s16"], help="device-resident decoded input format")
    ap.add_argument("--no-api-leg", action="store_true", help="skip the leg through the drop-in library's public API")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short runs of the other BASELINE configurations")
    ap.add_argument("--peak-db", default="", help="experiment: override the per-stream peak range, e.g. -40,-30")
    ap.add_argument("--peak-ref", default="rendered", choices=["rendered", "input"],
                    help="what the per-stream peak level refers to: the pre-limiter rendered mix (SURVEY 8d: ~20 %% of "
                         "streams above the limiter threshold) or the decoded input channels")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
