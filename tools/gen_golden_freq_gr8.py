"""Golden vectors for the GROUPED config-4 architecture ("gr8": conv_group_ratio = tr_conv_group_ratio = 8, BASELINE config 4 as
named) from the UNMODIFIED reference FreqCodec (build container only).  Weights = funcodec_b200.weights.init_state_dict(cfg, 0)
loaded into the reference module (not stored in the fixture)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from gen_golden_freq import OUT, build  # noqa: E402

if __name__ == "__main__":
    torch.set_num_threads(8)
    from funcodec_b200 import get_config, init_state_dict
    cfg = get_config("freqcodec_magphase_16k_n32_ds320_gr8")
    sd = init_state_dict(cfg, 0)
    ratios = [[f, t] for f, t in zip(cfg.ratios_f, cfg.ratios)]
    m = build(cfg.n_filters, cfg.dimension, cfg.codebook_size, cfg.num_quantizers, ratios, cfg.conv_group_ratio, cfg.tr_conv_group_ratio)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.split(".")[-1] in ("cluster_size", "embed_avg", "inited", "window") for k in missing), (missing, unexpected)
    m.quantizer.rq.model.inited.fill_(1)
    g = torch.Generator().manual_seed(14)
    wav = 0.1 * torch.randn(2, 8000 + 91, generator=g)
    with torch.no_grad():
        r = m.inference(wav, need_recon=True, bit_width=None, use_scale=True)
        emb, scale = m._encode(wav.unsqueeze(1))[0]
    out = dict(cfg_name=cfg.name, seed=0, wav=wav.numpy(), codes=r["code_indices"][0].numpy().astype(np.int16),
               quant=r["code_embeddings"][0][0].numpy(), scale=r["code_embeddings"][0][1].numpy(), recon=r["recon_speech"].numpy(),
               encoder_out=emb.numpy(), sd_checksum=float(sum(v.double().abs().sum().item() for v in sd.values())))
    path = os.path.join(OUT, "freq_magphase_config4_gr8_arch.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", "codes", out["codes"].shape, "recon", out["recon"].shape)
