"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the build
container.  Run:  python tools/gen_golden.py      (needs /root/reference; NOT run on the GPU box).

Fixtures are small: model weights are NOT stored, they are re-derived from
funcodec_b200.weights.init_state_dict(cfg, seed) (a checksum is stored to detect drift) and loaded
into the reference modules with load_state_dict (reference names).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from ref_harness import build_reference_encodec, import_reference  # noqa: E402
from funcodec_b200 import get_config, init_state_dict  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def sd_checksum(sd):
    return float(sum(v.double().abs().sum().item() for v in sd.values()))


def model_case(cfg_name, seed, B, L, wav_seed, bit_widths=(None,), tag=""):
    cfg = get_config(cfg_name)
    sd = init_state_dict(cfg, seed)
    model = build_reference_encodec(cfg)
    model.load_state_dict(sd)
    g = torch.Generator().manual_seed(wav_seed)
    wav = 0.1 * torch.randn(B, L, generator=g)
    out = dict(cfg_name=cfg_name, seed=seed, wav_seed=wav_seed, B=B, L=L, sd_checksum=sd_checksum(sd),
               wav=wav.numpy())
    with torch.no_grad():
        for bw in bit_widths:
            key = "full" if bw is None else f"bw{bw}"
            r = model.inference(wav, need_recon=True, bit_width=bw, use_scale=True)
            out[f"{key}.codes"] = r["code_indices"][0].numpy().astype(np.int16)
            out[f"{key}.quant"] = r["code_embeddings"][0][0].numpy()
            out[f"{key}.scale"] = r["code_embeddings"][0][1].numpy()
            out[f"{key}.recon"] = r["recon_speech"].numpy()
            if bw is None:
                out["full.sub_quants_sum"] = r["sub_quants"][0].double().sum(dim=(2, 3)).numpy()
                emb, scale = model._encode(wav.unsqueeze(1))[0]
                out["encoder_out"] = emb.numpy()
                # decode from codes (inference_decoding) and from embeddings (inference_decoding_emb)
                toks = r["code_indices"][0].permute(1, 2, 0)
                out["decode_codes.recon"] = model.inference_decoding(toks)["recon_speech"].numpy()
                out["decode_emb.recon"] = model.inference_decoding_emb(r["code_embeddings"][0][0])["recon_speech"].numpy()
                r2 = model.inference(wav, need_recon=True, bit_width=None, use_scale=False)
                out["noscale.recon"] = r2["recon_speech"].numpy()
    path = os.path.join(OUT, f"model_{cfg_name}{tag}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def layer_cases():
    """Layer-level vectors straight from the reference modules (conv.py / lstm.py / ddp_core_vq.py)."""
    import_reference()
    from funcodec.modules.normed_modules.conv import SConv1d, SConvTranspose1d
    from funcodec.modules.normed_modules.lstm import SLSTM
    from funcodec.modules.quantization.ddp_core_vq import DistributedResidualVectorQuantization
    from funcodec.models.encoder.seanet_encoder import SEANetResnetBlock
    out = {}
    g = torch.Generator().manual_seed(7)
    torch.manual_seed(7)
    idx = 0
    # (cin, cout, k, s, d, T): every (k, s) family of the named configs, odd lengths (extra padding),
    # and a short input (T <= max_pad -> zero-extend-then-reflect branch, conv.py:89-97)
    for (cin, cout, k, s, d, T) in [(1, 8, 7, 1, 1, 50), (8, 4, 3, 1, 1, 37), (4, 8, 1, 1, 1, 33), (8, 16, 4, 2, 1, 41),
                                    (8, 16, 8, 4, 1, 43), (8, 16, 10, 5, 1, 52), (8, 16, 16, 8, 1, 77), (6, 5, 7, 1, 1, 3),
                                    (8, 1, 7, 1, 1, 64), (8, 16, 16, 8, 1, 5), (4, 4, 3, 1, 2, 30)]:
        m = SConv1d(cin, cout, k, stride=s, dilation=d, norm="time_group_norm").eval()
        with torch.no_grad():
            m.conv.norm.weight.copy_(1 + 0.1 * torch.randn(cout, generator=g))
            m.conv.norm.bias.copy_(0.1 * torch.randn(cout, generator=g))
            x = torch.randn(2, cin, T, generator=g)
            y = m(x)
        pre = f"conv{idx}."
        out[pre + "meta"] = np.array([cin, cout, k, s, d, T])
        out[pre + "x"] = x.numpy(); out[pre + "y"] = y.numpy()
        out[pre + "w"] = m.conv.conv.weight.detach().numpy(); out[pre + "b"] = m.conv.conv.bias.detach().numpy()
        out[pre + "gw"] = m.conv.norm.weight.detach().numpy(); out[pre + "gb"] = m.conv.norm.bias.detach().numpy()
        idx += 1
    out["n_conv"] = np.array(idx)
    idx = 0
    for (cin, cout, s, T) in [(16, 8, 8, 9), (16, 8, 5, 11), (8, 4, 4, 13), (8, 4, 2, 17), (4, 2, 2, 1)]:
        m = SConvTranspose1d(cin, cout, 2 * s, stride=s, norm="time_group_norm").eval()
        with torch.no_grad():
            m.convtr.norm.weight.copy_(1 + 0.1 * torch.randn(cout, generator=g))
            m.convtr.norm.bias.copy_(0.1 * torch.randn(cout, generator=g))
            x = torch.randn(2, cin, T, generator=g)
            y = m(x)
        pre = f"convtr{idx}."
        out[pre + "meta"] = np.array([cin, cout, 2 * s, s, T])
        out[pre + "x"] = x.numpy(); out[pre + "y"] = y.numpy()
        out[pre + "w"] = m.convtr.convtr.weight.detach().numpy(); out[pre + "b"] = m.convtr.convtr.bias.detach().numpy()
        out[pre + "gw"] = m.convtr.norm.weight.detach().numpy(); out[pre + "gb"] = m.convtr.norm.bias.detach().numpy()
        idx += 1
    out["n_convtr"] = np.array(idx)
    # residual block
    rbm = SEANetResnetBlock(8, kernel_sizes=[3, 1], dilations=[1, 1], norm="time_group_norm", true_skip=False).eval()
    with torch.no_grad():
        x = torch.randn(2, 8, 45, generator=g)
        out["rb.x"] = x.numpy(); out["rb.y"] = rbm(x).numpy()
    for k, v in rbm.state_dict().items():
        out["rb.sd." + k] = v.numpy()
    # SLSTM
    lm = SLSTM(16, num_layers=2).eval()
    with torch.no_grad():
        x = torch.randn(3, 16, 21, generator=g)
        out["lstm.x"] = x.numpy(); out["lstm.y"] = lm(x).numpy()
    for k, v in lm.state_dict().items():
        out["lstm.sd." + k] = v.numpy()
    # RVQ eval forward / encode / decode
    rq = DistributedResidualVectorQuantization(num_quantizers=6, dim=16, codebook_size=32, kmeans_init=False,
                                               decay=0.99, kmeans_iters=10, threshold_ema_dead_code=2).eval()
    with torch.no_grad():
        rq.embed.copy_(torch.randn(6, 32, 16, generator=g) * (0.8 ** torch.arange(6.)).view(6, 1, 1))
        x = torch.randn(2, 16, 19, generator=g)
        quant, codes, _, sub = rq(x, n_q=6)
        q4, codes4, _, _ = rq(x, n_q=4)
        dec = rq.decode(codes)
        enc_codes = rq.encode(x, n_q=6)
    assert torch.equal(enc_codes, codes)
    out["rvq.embed"] = rq.embed.numpy(); out["rvq.x"] = x.numpy(); out["rvq.quant"] = quant.numpy()
    out["rvq.codes"] = codes.numpy(); out["rvq.sub"] = sub.numpy(); out["rvq.quant4"] = q4.numpy()
    out["rvq.codes4"] = codes4.numpy(); out["rvq.decode"] = dec.numpy()
    path = os.path.join(OUT, "layers.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    layer_cases()
    model_case("tiny_ds40", 0, 3, 40 * 23 + 17, 11, bit_widths=(None,), tag="_ragged")       # L % hop != 0
    model_case("tiny_ds40", 0, 2, 40 * 30, 12)
    model_case("small_ds320", 1, 2, 320 * 12, 13, bit_widths=(None, 1000, 2000))
    model_case("encodec_16k_n32_ds640", 0, 2, 16000, 1235, bit_widths=(None, 4000))
    model_case("encodec_16k_n32_ds320", 0, 1, 16000, 1236, bit_widths=(None, 1000))
