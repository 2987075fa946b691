#!/usr/bin/env python
"""Convert checkpoints between this framework's flat-shard layout and the reference's torch.distributed.checkpoint (DCP)
layout (open_diloco/ckpt_utils.py:48-156).  See opendiloco_b200/utils/dcp_interop.py.

    python scripts/convert_ckpt.py to-dcp   <our checkpoint rank dir>  <out dir>
    python scripts/convert_ckpt.py from-dcp <reference checkpoint dir> <out dir> --model 150m
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opendiloco_b200.models.config import LlamaConfig  # noqa: E402
from opendiloco_b200.utils.dcp_interop import export_dcp, import_dcp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("direction", choices=["to-dcp", "from-dcp"])
ap.add_argument("src")
ap.add_argument("dst")
ap.add_argument("--model", default="150m", help="preset name or config directory (from-dcp only)")
a = ap.parse_args()
if a.direction == "to-dcp":
    export_dcp(a.src, a.dst)
else:
    import_dcp(a.src, a.dst, LlamaConfig.from_pretrained(a.model))
print(f"wrote {a.dst}")
