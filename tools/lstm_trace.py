"""PROFILING ONLY: per-item %globaltimer stamps of CTA 0 of the last LSTM layer launch (FCB_LSTM_TRACE=1).
usage: FCB_LSTM_TRACE=1 python tools/lstm_trace.py <config> <B> <L>"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from funcodec_b200 import get_config, init_state_dict
from funcodec_b200.encodec import B200Encodec
name, B, L = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = get_config(name)
m = B200Encodec(cfg, init_state_dict(cfg, 0), "cuda:0")
x = 0.1 * torch.randn(B, L, device="cuda")
for _ in range(3):
    m.inference(x, need_sub_quants=False)
torch.cuda.synchronize()
del m
