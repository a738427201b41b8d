"""In-process multi-GPU through the drop-in library: the bench's public-API leg (IAMF_decoder_decode_batch_units on ipcm-coded
configuration-2 streams, 1024 handles per GPU) with IAMF_B200_DEVICES naming 1, 2, ... GPUs of the box - ONE process, one host
thread + context + pinned buffers per device (iac_b200/host/iamf_decoder.c).  Prints one JSON line per device count.

    python tools/api_devices.py [max devices]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch
import bench
from iac_b200 import shard
d = int(os.environ["API_DEVICES_N"])
r = bench.api_leg("c2", 1024 * d, 8, 0, None, shard, torch.device("cuda", 0))
r["devices"] = d
r["handles_total"] = r.pop("handles_per_gpu")
print("RESULT " + json.dumps(r))
"""


def main():
    import torch
    have = torch.cuda.device_count()
    top = min(have, int(sys.argv[1]) if len(sys.argv) > 1 else have)
    d = 1
    while d <= top:
        env = dict(os.environ, IAMF_B200_DEVICES=",".join(str(i) for i in range(d)), API_DEVICES_N=str(d))
        r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT)], capture_output=True, text=True, env=env, timeout=1200)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
        print(line[-1][7:] if line else json.dumps({"devices": d, "error": r.stderr[-800:]}), flush=True)
        d *= 2


if __name__ == "__main__":
    main()
