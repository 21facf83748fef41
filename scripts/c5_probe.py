"""One C5 bucket through decode + merge, for profiling: python scripts/c5_probe.py [none|zstd]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    import torch
    torch.cuda.init()
    peak = 6568.4
    out = bench.extra_c5(0, peak, steps=2, codecs=tuple(sys.argv[1:]) or ("none", "zstd"))
    print(json.dumps(out))
