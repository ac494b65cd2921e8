"""Golden vectors for the conf/soundstream_noncausal_16k_n32_600k_step.yaml topology (time_group_norm, non-causal,
n_residual_layers 3 with dilations 1 / 2 / 4, seq_model none, wide embedding) and for conf/soundstream_16k_n32_600k_step.yaml's
(the same stacks with norm weight_norm and causal true) from the UNMODIFIED reference SEANetEncoder / SEANetDecoder, at small
widths.  Build container only:  python tools/gen_golden_soundstream.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from ref_harness import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    one("soundstream_noncausal_small.npz", "time_group_norm", False, 21)
    one("soundstream_causal_small.npz", "weight_norm", True, 31)


def one(fname, norm, causal, seed):
    _, SEANetEncoder, SEANetDecoder, _ = import_reference()
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed + 1)
    ratios = [5, 4, 2]
    kw = dict(n_filters=4, ratios=ratios, norm=norm, causal=causal, n_residual_layers=3, dilation_base=2,
              seq_model="none", kernel_size=7, last_kernel_size=7, residual_kernel_size=3)
    enc = SEANetEncoder(input_size=1, dimension=48, **kw).eval()
    dec = SEANetDecoder(input_size=48, channels=1, **kw).eval()
    with torch.no_grad():
        for m in (enc, dec):
            for name, par in m.named_parameters():
                if name.endswith("norm.weight"):
                    par.copy_(1 + 0.1 * torch.randn(par.shape, generator=g))
                elif name.endswith("norm.bias"):
                    par.copy_(0.1 * torch.randn(par.shape, generator=g))
                elif name.endswith("weight_g"):     # weight_norm starts at g = ||v||: move it, with a gain that keeps the
                    par.mul_(1.3 * (1 + 0.1 * torch.randn(par.shape, generator=g)))   # un-normalised activations O(1)
        x = 0.3 * torch.randn(2, 1, 40 * 13 + 7, generator=g)
        emb = enc(x)                      # [B, T', D]
        y = dec(emb)
    out = dict(ratios=np.array(ratios), x=x.numpy(), emb=emb.numpy(), y=y.numpy())
    for k, v in enc.state_dict().items():
        out["sd.encoder." + k] = v.numpy()
    for k, v in dec.state_dict().items():
        out["sd.decoder." + k] = v.numpy()
    path = os.path.join(OUT, fname)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", tuple(emb.shape), tuple(y.shape),
          "emb rms %.3g y rms %.3g" % (emb.pow(2).mean().sqrt().item(), y.pow(2).mean().sqrt().item()))


if __name__ == "__main__":
    main()
