"""Builds the in-tree native libraries of iac_b200 (nvcc, sm_100a only; cross-compiles without a GPU).

    python -m iac_b200.build            # build everything that is stale
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libiamf_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--fmad=false",          # the reference is built without FMA contraction (x86-64 baseline): keep mul and add apart
    "--expt-relaxed-constexpr",   # constexpr table look-ups (iamfb_stream.cuh) are evaluated at compile time on both sides
    "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False):
    """every csrc/*.cu is one translation unit (kernels + their launchers), compiled in parallel into build/*.o and
    linked into ONE shared library; headers (*.cuh, *.inc, include/*.h) are dependencies of every unit"""
    files = sorted(os.listdir(CSRC))
    units = [f for f in files if f.endswith(".cu")]
    hdrs = [os.path.join(CSRC, f) for f in files if not f.endswith(".cu")] + [os.path.join(ROOT, "include", "iamf_b200.h")]
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    procs, objs = [], []
    for u in units:
        src, obj = os.path.join(CSRC, u), os.path.join(objdir, u[:-3] + ".o")
        objs.append(obj)
        if not force and not _stale(obj, [src] + hdrs):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
            "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-c", "-o", obj, src]
        print("[iac_b200.build]", " ".join(cmd), file=sys.stderr)
        procs.append((u, subprocess.Popen(cmd)))
    failed = [u for u, p in procs if p.wait() != 0]
    if failed:
        raise RuntimeError("nvcc failed for " + ", ".join(failed))
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        print("[iac_b200.build]", " ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return LIB


HOST = os.path.join(PKG, "host")
DROPIN = os.path.join(PKG, "libiamf.so")
REF_OPUS = "/root/reference/dep_codecs/lib/libopus.a"
REF_FLAC = "/root/reference/dep_codecs/lib/libFLAC.a"


def build_dropin(force=False):
    """the drop-in libiamf.so: plain-C host layer (include/IAMF_decoder.h API) on top of libiamf_b200.so.
    Opus / FLAC entropy decode is linked from the reference tree's prebuilt libopus.a / libFLAC.a when they exist
    (authoring container); without them the library is built ipcm-only."""
    srcs = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".c")]
    deps = srcs + [os.path.join(HOST, "iamf_host.h"), os.path.join(HOST, "libiamf.map"), os.path.join(ROOT, "include", "IAMF_decoder.h"),
                   os.path.join(ROOT, "include", "iamf_b200.h"), LIB]
    if not force and not _stale(DROPIN, deps):
        return DROPIN
    cmd = [os.environ.get("CC", "gcc"), "-std=c99", "-O2", "-Wall", "-fPIC", "-shared", "-fvisibility=default",
           "-I" + os.path.join(ROOT, "include"), "-I" + HOST, "-o", DROPIN] + srcs
    if os.path.exists(REF_OPUS):
        cmd += ["-DIH_HAVE_OPUS", REF_OPUS]
    if os.path.exists(REF_FLAC):
        cmd += ["-DIH_HAVE_FLAC", REF_FLAC]
    cmd += ["-L" + PKG, "-l:libiamf_b200.so", "-Wl,-rpath,$ORIGIN", "-Wl,--exclude-libs,ALL",
            "-Wl,--version-script=" + os.path.join(HOST, "libiamf.map"), "-lm"]
    print("[iac_b200.build]", " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return DROPIN


PLAYER_SRC = os.path.join(PKG, "player")
PLAYER = os.path.join(PKG, "iamfplayer_b200")


def build_player(force=False):
    """the command-line player (IAMF bitstream / MP4 in, WAV + .met out) on top of libiamf.so: own WAV writer and MP4 reader"""
    srcs = [os.path.join(PLAYER_SRC, f) for f in sorted(os.listdir(PLAYER_SRC)) if f.endswith(".c")]
    deps = srcs + [os.path.join(PLAYER_SRC, f) for f in os.listdir(PLAYER_SRC) if f.endswith(".h")] + [
        os.path.join(ROOT, "include", "IAMF_decoder.h"), DROPIN]
    if not force and not _stale(PLAYER, deps):
        return PLAYER
    cmd = [os.environ.get("CC", "gcc"), "-std=c99", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + PLAYER_SRC, "-o", PLAYER] + srcs + [
        "-L" + PKG, "-l:libiamf.so", "-Wl,-rpath,$ORIGIN"]
    print("[iac_b200.build]", " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return PLAYER


def build_all(force=False):
    build_cuda(force)
    lib = build_dropin(force)
    build_player(force)
    return lib


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
