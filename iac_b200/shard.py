"""Multi-GPU plumbing of the rendering path: streams shard independently, there is no collective on the data path.

Every decoder handle / stream is independent in the reference (no shared mutable state, IAMF_decoder_private.h:312-345),
and all frames of one stream must stay on one GPU because the limiter, resampler and de-mixer carry state from frame to
frame.  So stream i lives on rank i mod world for its whole life; torch.distributed (NCCL on the GPU box, gloo in the
CPU tests) is only used for the barrier around the timed region and to reduce two scalars: the MAX of the per-rank
device time and the SUM of the audio produced.
"""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_streams(n_streams_total, rank, world):
    """global stream ids rendered by `rank` (round robin, like a front-end assigning new sessions)"""
    return list(range(rank, n_streams_total, world))


def stream_seed(base_seed, global_stream_id):
    """the seed of a synthetic stream depends on its GLOBAL id only, so the job renders the same audio for every world size"""
    return base_seed + global_stream_id


def aggregate(ms_local, out_samples_local, device=None):
    """(max over ranks of the device time in ms, sum over ranks of the output samples); identity when not distributed"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms_local), float(out_samples_local)
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    n = torch.tensor([float(out_samples_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), float(n.item())


def job_throughput(ms_max, out_samples_total, out_rate, steps=1):
    """whole-job rendered audio-seconds per second: all ranks' audio / the slowest rank's time"""
    return (out_samples_total / float(out_rate)) * steps / (ms_max / 1e3)
