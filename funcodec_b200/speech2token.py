"""`Speech2Token` mirror (funcodec/bin/codec_inference.py:41-151): the run_mod dispatch the CLI performs on the model,
on top of B200Encodec.  Same call signature and return tuple, same `decode` bit-width arithmetic."""
import math
from typing import Optional, Union

import numpy as np
import torch

from .config import CodecConfig
from .encodec import B200Encodec


class Speech2Token:
    def __init__(self, model: B200Encodec, device: str = "cuda:0", need_sub_quants: bool = True):
        self.model = model
        self.device = device
        # the reference always materialises sub_quants [n_q, B, D, T']; the CLI only reads them under --need_sub_quants
        # (codec_inference.py:282-286), so the driver may switch the extra tensor off
        self.need_sub_quants = need_sub_quants

    @classmethod
    def from_state_dict(cls, cfg: CodecConfig, state_dict, device: str = "cuda:0"):
        return cls(B200Encodec(cfg, state_dict, device), device)

    @classmethod
    def from_pretrained(cls, model_tag: Optional[str] = None, config_file: Optional[str] = None, model_file: Optional[str] = None,
                        device: str = "cuda:0", **kwargs):
        """codec_inference.py:136-150 without the model hub: `model_tag` cannot be downloaded here (no network), so the
        YAML and the checkpoint have to be given as files."""
        if model_tag is not None:
            raise NotImplementedError("model_tag (model hub download) is not available; pass config_file / model_file")
        from .bin.codec_inference import build_speech2token
        return build_speech2token(config_file, model_file, device, need_sub_quants=bool(kwargs.get("need_sub_quants", True)))

    @torch.no_grad()
    def __call__(self, speech: Union[torch.Tensor, np.ndarray], ppg=None, need_recon: bool = True,
                 bit_width: Optional[int] = None, use_scale: bool = True, run_mod: str = "inference"):
        """codec_inference.py:86-134.  `ppg` (CodecSemanticAug only) is not supported."""
        if ppg is not None:
            raise ValueError("ppg input belongs to codec_semantic_aug models, which are out of scope")
        if isinstance(speech, np.ndarray):
            speech = torch.from_numpy(speech)
        m = self.model
        if run_mod == "inference":
            ret = m.inference(speech, need_recon=need_recon, bit_width=bit_width, use_scale=use_scale,
                              need_sub_quants=self.need_sub_quants)
        elif run_mod == "encode":
            ret = m.inference_encoding(speech, need_recon=False, bit_width=bit_width, need_sub_quants=self.need_sub_quants)
        elif run_mod == "decode_emb":
            ret = m.inference_decoding_emb(speech)
        else:
            q = m.quantizer
            bit_per_quant = (q.sampling_rate // q.encoder_hop_length) * int(math.log2(q.codebook_size))
            nq = None
            if bit_width is not None:
                nq = int(max(bit_width // bit_per_quant, 1))
            ret = m.inference_decoding(speech[:, :, :nq])
        return ret["code_indices"], ret["code_embeddings"], ret["recon_speech"], ret["sub_quants"]
