"""generate_stream overlap: the codec step of frame i next to a simulated producer of frame i+1 —
`bench.llm_overlap` on its own.  usage: python tools/stream_overlap_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from fireredtts2_b200.codec import RedCodecB200
from fireredtts2_b200.config import C0
from fireredtts2_b200.weights import synthetic_state_dict

codec = RedCodecB200(C0, synthetic_state_dict(C0, 0), device="cuda:0")
codec.stream_max_tokens = 1200
print(json.dumps(bench.llm_overlap(codec, C0, torch.device("cuda", 0))))
