"""Layer-level golden vectors for the 2-D path straight from the reference modules SConv2d / SConvTranspose2d /
SEANetResnetBlock2d (funcodec/modules/normed_modules/conv.py:317-447, models/encoder/seanet_encoder.py:188-249).
Build container only:  python tools/gen_golden_freq_layers.py  -> tests/golden/freq_layers.npz"""
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
    import_reference()
    from funcodec.modules.normed_modules.conv import SConv2d, SConvTranspose2d
    from funcodec.models.encoder.seanet_encoder import SEANetResnetBlock2d
    g = torch.Generator().manual_seed(11)
    torch.manual_seed(11)
    out = {}
    nk = {"num_groups": 1}

    def randomise_norm(m):
        with torch.no_grad():
            for name, par in m.named_parameters():
                if name.endswith("norm.weight"):
                    par.copy_(1 + 0.1 * torch.randn(par.shape, generator=g))
                elif name.endswith("norm.bias"):
                    par.copy_(0.1 * torch.randn(par.shape, generator=g))

    # (cin, cout, (kf, kt), (sf, st), F, T): every kernel / stride family of the FreqCodec stacks, odd T (extra padding
    # on the LEFT of the time axis, conv.py:368), F not a multiple of the stride
    convs = [(3, 4, (7, 7), (1, 1), 20, 33), (4, 2, (3, 3), (1, 1), 9, 17), (2, 4, (1, 1), (1, 1), 5, 11),
             (4, 8, (8, 2), (4, 1), 37, 21), (8, 16, (8, 4), (4, 2), 16, 31), (8, 16, (8, 4), (4, 2), 18, 30)]
    for i, (cin, cout, k, s, F, T) in enumerate(convs):
        m = SConv2d(cin, cout, k, stride=s, norm="time_group_norm", norm_kwargs=nk).eval()
        randomise_norm(m)
        x = torch.randn(2, cin, F, T, generator=g)
        with torch.no_grad():
            y = m(x)
        out[f"conv{i}.meta"] = np.array([cin, cout, k[0], k[1], s[0], s[1]])
        out[f"conv{i}.x"] = x.numpy(); out[f"conv{i}.y"] = y.numpy()
        for kk, v in m.state_dict().items():
            out[f"conv{i}.sd.{kk}"] = v.numpy()
    # transposed: k = 2 s per axis; the decoder's last stage uses out_padding [(0, 1), (0, 0)]
    convtrs = [(16, 8, (4, 1), ((0, 0), (0, 0)), 1, 13), (8, 4, (4, 2), ((0, 0), (0, 0)), 4, 9), (4, 2, (4, 1), ((0, 1), (0, 0)), 16, 10)]
    for i, (cin, cout, s, op, F, T) in enumerate(convtrs):
        k = (2 * s[0], 2 * s[1])
        m = SConvTranspose2d(cin, cout, k, stride=s, norm="time_group_norm", norm_kwargs=nk, out_padding=[op[0], op[1]]).eval()
        randomise_norm(m)
        x = torch.randn(2, cin, F, T, generator=g)
        with torch.no_grad():
            y = m(x)
        out[f"convtr{i}.meta"] = np.array([cin, cout, s[0], s[1], op[0][0], op[0][1], op[1][0], op[1][1]])
        out[f"convtr{i}.x"] = x.numpy(); out[f"convtr{i}.y"] = y.numpy()
        for kk, v in m.state_dict().items():
            out[f"convtr{i}.sd.{kk}"] = v.numpy()
    # residual block (true_skip = False: 1x1 conv shortcut), as the stacks build it
    m = SEANetResnetBlock2d(8, kernel_sizes=[(3, 3), (1, 1)], dilations=[(1, 1), (1, 1)], norm="time_group_norm",
                            norm_params=nk, causal=False, compress=2, true_skip=False, conv_group_ratio=-1).eval()
    randomise_norm(m)
    x = torch.randn(2, 8, 7, 15, generator=g)
    with torch.no_grad():
        y = m(x)
    out["rb.x"] = x.numpy(); out["rb.y"] = y.numpy()
    for kk, v in m.state_dict().items():
        out[f"rb.sd.{kk}"] = v.numpy()
    path = os.path.join(OUT, "freq_layers.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", len(out), "arrays")


if __name__ == "__main__":
    main()
