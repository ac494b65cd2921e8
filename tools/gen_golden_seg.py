"""Generate tests/golden/model_small_ds320_segmented.npz: the UNMODIFIED reference Encodec (/root/reference) run with
segment_dur != None (codec_basic.py:287-298,334-359,382-396: per-segment normalise / encode / quantize / decode and
_linear_overlap_add).  Build container only:  python tools/gen_golden_seg.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from ref_harness import build_reference_encodec  # noqa: E402
from funcodec_b200 import get_config, init_state_dict  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CASES = [  # (tag, segment_dur [s], overlap_ratio, B, L, wav_seed)
    ("a", 0.1, 0.3, 2, 5000, 11),      # 1600-sample segments (5 frames), stride 1120: 4 full segments + one 520-sample tail
    ("b", 0.11, 0.1, 1, 4000, 12),     # 1760 samples (not a multiple of hop 320: decoded frames are 1920 long), stride 1584
]


def main():
    cfg = get_config("small_ds320")
    sd = init_state_dict(cfg, 0)
    model = build_reference_encodec(cfg)
    model.load_state_dict(sd)
    out = dict(cfg_name="small_ds320", seed=0)
    for tag, dur, ov, B, L, ws in CASES:
        model.segment_dur, model.overlap_ratio = dur, ov
        g = torch.Generator().manual_seed(ws)
        wav = 0.1 * torch.randn(B, L, generator=g)
        with torch.no_grad():
            r = model.inference(wav, need_recon=True, bit_width=None, use_scale=True)
            frames = model._encode(wav.unsqueeze(1))
        out[f"{tag}.meta"] = np.array([dur, ov, B, L, ws, model.segment_length, model.segment_stride, len(r["code_indices"])])
        out[f"{tag}.wav"] = wav.numpy()
        out[f"{tag}.recon"] = r["recon_speech"].numpy()
        for i, (codes, (quant, scale), (emb, _)) in enumerate(zip(r["code_indices"], r["code_embeddings"], frames)):
            out[f"{tag}.codes{i}"] = codes.numpy().astype(np.int16)
            out[f"{tag}.quant{i}"] = quant.numpy()
            out[f"{tag}.scale{i}"] = scale.numpy()
            out[f"{tag}.encoder_out{i}"] = emb.numpy()
        print(tag, "segments", len(r["code_indices"]), [tuple(c.shape) for c in r["code_indices"]], tuple(r["recon_speech"].shape))
    path = os.path.join(OUT, "model_small_ds320_segmented.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
