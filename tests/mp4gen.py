"""Test infrastructure: a minimal MP4 muxer for IAMF (one audio track, sample entry `iamf` whose payload behind the 28
AudioSampleEntry bytes is the descriptor OBUs - the container draft the reference's player reads,
test/tools/iamfplayer/src/mp4demux.c:512-574), plain (stsc/stsz/stco) or fragmented (moof/traf/tfhd/trun)."""
import struct


def box(kind, payload):
    return struct.pack(">I4s", 8 + len(payload), kind) + payload


def full(kind, version, flags, payload):
    return box(kind, struct.pack(">I", (version << 24) | flags) + payload)


MATRIX = struct.pack(">9I", 0x10000, 0, 0, 0, 0x10000, 0, 0, 0, 0x40000000)


def _sample_entry(desc, channels=2, rate=48000):
    return box(b"iamf", b"\0" * 6 + struct.pack(">H", 1) + b"\0" * 8 + struct.pack(">HHHHI", channels, 16, 0, 0, rate << 16) + desc)


def _moov(descs, sizes, deltas, desc_of_chunk, chunk_offsets, timescale, skip, fragmented, samples_per_chunk):
    n = len(sizes)
    duration = sum(deltas)
    mvhd = full(b"mvhd", 0, 0, struct.pack(">IIII", 0, 0, timescale, duration) + struct.pack(">IH", 0x10000, 0x100) + b"\0" * 10 + MATRIX +
                b"\0" * 24 + struct.pack(">I", 2))
    tkhd = full(b"tkhd", 0, 3, struct.pack(">IIIII", 0, 0, 1, 0, duration) + b"\0" * 8 + struct.pack(">HHHH", 0, 0, 0x100, 0) + MATRIX +
                struct.pack(">II", 0, 0))
    edts = box(b"edts", full(b"elst", 0, 0, struct.pack(">IIiHH", 1, duration, skip, 1, 0))) if skip else b""
    mdhd = full(b"mdhd", 0, 0, struct.pack(">IIIIHH", 0, 0, timescale, duration, 0x55C4, 0))
    hdlr = full(b"hdlr", 0, 0, struct.pack(">I4s", 0, b"soun") + b"\0" * 12 + b"\0")
    stsd = full(b"stsd", 0, 0, struct.pack(">I", len(descs)) + b"".join(_sample_entry(d) for d in descs))
    if fragmented:
        stts = full(b"stts", 0, 0, struct.pack(">I", 0))
        stsc = full(b"stsc", 0, 0, struct.pack(">I", 0))
        stsz = full(b"stsz", 0, 0, struct.pack(">II", 0, 0))
        stco = full(b"stco", 0, 0, struct.pack(">I", 0))
    else:
        runs = []
        for d in deltas:
            if runs and runs[-1][1] == d:
                runs[-1][0] += 1
            else:
                runs.append([1, d])
        stts = full(b"stts", 0, 0, struct.pack(">I", len(runs)) + b"".join(struct.pack(">II", c, d) for c, d in runs))
        n_chunks = len(chunk_offsets)
        ent = []
        for c in range(n_chunks):
            per = min(samples_per_chunk, n - c * samples_per_chunk)
            if not ent or ent[-1][1:] != (per, desc_of_chunk[c]):
                ent.append((c + 1, per, desc_of_chunk[c]))
        stsc = full(b"stsc", 0, 0, struct.pack(">I", len(ent)) + b"".join(struct.pack(">III", *e) for e in ent))
        stsz = full(b"stsz", 0, 0, struct.pack(">II", 0, n) + b"".join(struct.pack(">I", s) for s in sizes))
        stco = full(b"stco", 0, 0, struct.pack(">I", n_chunks) + b"".join(struct.pack(">I", o) for o in chunk_offsets))
    stbl = box(b"stbl", stsd + stts + stsc + stsz + stco)
    dinf = box(b"dinf", full(b"dref", 0, 0, struct.pack(">I", 1) + full(b"url ", 0, 1, b"")))
    minf = box(b"minf", full(b"smhd", 0, 0, b"\0" * 4) + dinf + stbl)
    trak = box(b"trak", tkhd + edts + box(b"mdia", mdhd + hdlr + minf))
    mvex = box(b"mvex", full(b"trex", 0, 0, struct.pack(">IIIII", 1, 1, 0, 0, 0))) if fragmented else b""
    return box(b"moov", mvhd + trak + mvex)


def mux(descs, samples, deltas, timescale=48000, skip=0, fragmented=False, samples_per_chunk=3, desc_of_sample=None, per_fragment=4):
    """descs: list of descriptor-OBU blobs (sample entries); samples: list of bytes (one temporal unit each); deltas: their
    durations; desc_of_sample: 1-based sample entry per sample (switches only at chunk starts).  Returns (file bytes,
    [(offset, size, delta, desc_index)] for checking a reader)."""
    if isinstance(descs, (bytes, bytearray)):
        descs = [bytes(descs)]
    n = len(samples)
    sizes = [len(s) for s in samples]
    desc_of_sample = desc_of_sample or [1] * n
    ftyp = box(b"ftyp", b"isom" + struct.pack(">I", 0x200) + b"isomiamf")
    table = []
    if not fragmented:
        n_chunks = (n + samples_per_chunk - 1) // samples_per_chunk
        doc = [desc_of_sample[c * samples_per_chunk] for c in range(n_chunks)]
        moov = _moov(descs, sizes, deltas, doc, [0] * n_chunks, timescale, skip, False, samples_per_chunk)
        base = len(ftyp) + len(moov) + 8
        offs, pos = [], base
        for c in range(n_chunks):
            offs.append(pos)
            for k in range(c * samples_per_chunk, min(n, (c + 1) * samples_per_chunk)):
                table.append((pos, sizes[k], deltas[k], doc[c]))
                pos += sizes[k]
        moov = _moov(descs, sizes, deltas, doc, offs, timescale, skip, False, samples_per_chunk)
        return ftyp + moov + box(b"mdat", b"".join(samples)), table
    out = ftyp + _moov(descs, sizes, deltas, [], [], timescale, skip, True, samples_per_chunk)
    seq = 1
    for f0 in range(0, n, per_fragment):
        idx = list(range(f0, min(n, f0 + per_fragment)))
        di = desc_of_sample[f0]
        # tfhd: default-base-is-moof | sample-description-index; trun: data-offset | duration | size
        def moof(data_offset):
            tfhd = full(b"tfhd", 0, 0x020002, struct.pack(">II", 1, di))
            trun = full(b"trun", 0, 0x000301, struct.pack(">Ii", len(idx), data_offset) + b"".join(struct.pack(">II", deltas[k], sizes[k]) for k in idx))
            return box(b"moof", full(b"mfhd", 0, 0, struct.pack(">I", seq)) + box(b"traf", tfhd + trun))
        m = moof(0)
        m = moof(len(m) + 8)
        pos = len(out) + len(m) + 8
        for k in idx:
            table.append((pos, sizes[k], deltas[k], di))
            pos += sizes[k]
        out += m + box(b"mdat", b"".join(samples[k] for k in idx))
        seq += 1
    return out, table
