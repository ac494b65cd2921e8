"""Glue for the drop-in: derive a CodecConfig from an already-built reference `Encodec`
module (funcodec/models/codec_basic.py:119) and wrap it with the CUDA-backed B200Encodec.

Only the configurations of DESIGN.md §9 are accepted; anything else raises (no silent fallback to the PyTorch path).
"""
from typing import Dict

import torch

from .config import CodecConfig


class UnsupportedReferenceModel(ValueError):
    pass


def config_from_reference_model(model) -> CodecConfig:
    """Reads the attributes SEANetEncoder / SEANetDecoder / CostumeQuantizer / Encodec keep
    (seanet_encoder.py:98-106, seanet_decoder.py:100-106, costume_quantizer.py:49-53, codec_basic.py:249-254)."""
    enc, dec, q = model.encoder, model.decoder, model.quantizer
    ratios = tuple(int(r) for r in dec.ratios)
    if tuple(reversed(ratios)) != tuple(int(r) for r in enc.ratios):
        raise UnsupportedReferenceModel("encoder/decoder ratios differ")
    sd: Dict[str, torch.Tensor] = model.state_dict()
    if any(k.endswith("weight_g") or k.endswith("weight_v") for k in sd):
        raise UnsupportedReferenceModel("weight_norm parametrisation is not supported (norm must be time_group_norm)")
    if "encoder.model.0.conv.norm.weight" not in sd:
        raise UnsupportedReferenceModel("norm must be time_group_norm")
    if getattr(model, "segment_dur", None) is not None:
        raise UnsupportedReferenceModel("segment_dur must be null (whole-utterance processing)")
    if getattr(q, "input_proj", None) is not None or getattr(q, "input_act", None) is not None:
        raise UnsupportedReferenceModel("quantizer projections / codec_range are not supported")
    if "quantizer.rq.model.embed" not in sd:
        raise UnsupportedReferenceModel("quantizer must use use_ddp: true (stacked codebook buffers)")
    if getattr(model, "codec_domain", "time") not in ("time", None):
        raise UnsupportedReferenceModel("only the time-domain Encodec is supported")
    embed = sd["quantizer.rq.model.embed"]
    w0 = sd["encoder.model.0.conv.conv.weight"]
    n_lstm = len([k for k in sd if k.startswith("decoder.model.1.lstm.weight_ih_l")])
    last_idx = max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.model."))
    rb_k = sd["encoder.model.1.block.1.conv.conv.weight"].shape[-1]
    cfg = CodecConfig(name="from_reference", ratios=ratios, n_filters=int(w0.shape[0]), dimension=int(embed.shape[2]),
                      kernel_size=int(w0.shape[-1]),
                      last_kernel_size=int(sd[f"decoder.model.{last_idx}.conv.conv.weight"].shape[-1]),
                      residual_kernel_size=int(rb_k), lstm_layers=n_lstm, codebook_size=int(embed.shape[1]),
                      num_quantizers=int(embed.shape[0]), sample_rate=int(q.sampling_rate),
                      audio_normalize=bool(model.audio_normalize))
    if cfg.hop_length != int(q.encoder_hop_length):
        raise UnsupportedReferenceModel("quantizer.encoder_hop_length does not match prod(ratios)")
    return cfg


def wrap_reference_encodec(model, device: str = "cuda:0"):
    from .encodec import B200Encodec
    cfg = config_from_reference_model(model)
    return B200Encodec(cfg, model.state_dict(), device)
