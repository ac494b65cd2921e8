"""Layer-level GPU parity: every conv family of the named configs through the fcb_debug_conv1d hook, on BOTH
kernels (tensor-core 3xTF32 implicit GEMM and fp32 SIMT), against a float64 torch-CPU evaluation of the same
layer (reflect pad + conv1d / conv_transpose1d + bias, then GroupNorm statistics).

Bar: max-abs error <= 2e-5 x rms(output) (fp32-faithful; plain TF32 would be ~1e-3), statistics to 1e-5 relative.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from funcodec_b200 import get_config, init_state_dict
from oracle import encodec_oracle as O

pytestmark = pytest.mark.gpu

_M = {}


def _models():
    from funcodec_b200.encodec import B200Encodec
    if not _M:
        cfg = get_config("encodec_16k_n32_ds640")
        sd = init_state_dict(cfg, 0)
        _M["cfg"], _M["sd"] = cfg, sd
        _M["tc"] = B200Encodec(cfg, sd, "cuda:0")
        _M["simt"] = B200Encodec(cfg, sd, "cuda:0", options={"use_tc": 0})
    return _M


def _truth(sd, layer, x_bct, elu):
    """float64 evaluation: returns (raw full output [B, C, T_full], row_off, kept_len)."""
    x = x_bct.double()
    if elu:
        x = F.elu(x)
    if ".lstm.ih" in layer:
        pre, l = layer.split(".lstm.ih")
        w = sd[f"{pre}.lstm.weight_ih_l{l}"].double()
        b = (sd[f"{pre}.lstm.bias_ih_l{l}"] + sd[f"{pre}.lstm.bias_hh_l{l}"]).double()
        H = w.shape[1]
        y = torch.einsum("bct,nc->bnt", x, w) + b.view(1, -1, 1)          # rows gate-major g*H + j
        y = y.view(x.shape[0], 4, H, -1).permute(0, 2, 1, 3).reshape(x.shape[0], 4 * H, -1)   # -> 4*j + g
        return y, 0, y.shape[-1]
    if layer + ".convtr.convtr.weight" in sd:
        w = sd[layer + ".convtr.convtr.weight"].double()
        b = sd[layer + ".convtr.convtr.bias"].double()
        s = w.shape[-1] // 2
        y = F.conv_transpose1d(x, w, b, stride=s)
        pt = w.shape[-1] - s
        pr = pt // 2
        return y, pt - pr, x.shape[-1] * s
    w = sd[layer + ".conv.conv.weight"].double()
    b = sd[layer + ".conv.conv.bias"].double()
    k = w.shape[-1]
    stride = _STRIDES.get(layer, 1)
    pl, pr = O.conv_paddings(x.shape[-1], k, stride, 1)
    y = F.conv1d(O.pad1d_reflect(x, (pl, pr)), w, b, stride=stride)
    return y, 0, y.shape[-1]


_STRIDES = {"encoder.model.3": 2, "encoder.model.6": 2, "encoder.model.9": 4, "encoder.model.12": 5, "encoder.model.15": 8}

CASES = [  # (layer, C_in, T_in, elu)
    ("encoder.model.1.block.1", 32, 300, True), ("encoder.model.1.shortcut", 32, 129, False),
    ("encoder.model.3", 32, 2 * 131 + 1, True), ("encoder.model.4.block.1", 64, 260, True),
    ("encoder.model.4.block.3", 32, 260, True), ("encoder.model.4.shortcut", 64, 128, False),
    ("encoder.model.6", 64, 2 * 140, True), ("encoder.model.9", 128, 4 * 150 + 3, True),
    ("encoder.model.12", 256, 5 * 133 + 2, True), ("encoder.model.15", 512, 8 * 130 + 5, True),
    ("encoder.model.18", 1024, 250, True), ("encoder.model.16.lstm.ih0", 1024, 250, False),
    ("decoder.model.0", 128, 250, False), ("decoder.model.3", 1024, 250, True), ("decoder.model.6", 512, 131, True),
    ("decoder.model.9", 256, 140, True), ("decoder.model.12", 128, 257, True), ("decoder.model.15", 64, 300, True),
    ("decoder.model.16.block.1", 32, 515, True), ("encoder.model.18", 1024, 5, True), ("decoder.model.3", 1024, 1, True),
    ("encoder.model.1.block.3", 16, 700, True), ("decoder.model.18", 32, 1500, True), ("decoder.model.18", 32, 3, True),
    ("encoder.model.0", 1, 1000, False),
]


@pytest.mark.parametrize("path", ["tc", "simt"])
@pytest.mark.parametrize("layer,cin,T,elu", CASES)
def test_conv_layer(layer, cin, T, elu, path):
    m = _models()
    model, sd = m[path], m["sd"]
    g = torch.Generator().manual_seed(abs(hash((layer, T))) % (2 ** 31))
    B = 2
    x_btc = torch.randn(B, T, cin, generator=g)
    y, stats, row_off = model.debug_conv(layer, x_btc, elu=elu, want_stats=".lstm." not in layer)
    y = y.cpu().double()
    ref, ref_off, kept = _truth(sd, layer, x_btc.permute(0, 2, 1), elu)
    ref_btc = ref.permute(0, 2, 1)
    assert row_off == ref_off
    assert y.shape == ref_btc.shape, (y.shape, ref_btc.shape)
    rms = ref_btc.pow(2).mean().sqrt().item()
    err = (y - ref_btc).abs().max().item()
    print(f"{path:5s} {layer:28s} T={T:5d} max-abs err {err:.3e} rms {rms:.3e} rel {err / rms:.3e}")
    assert err <= 2e-5 * rms, (path, layer, err, rms)
    if stats is not None:
        st = stats.cpu().double()
        mean = ref.mean(dim=(1, 2))
        var = ref.var(dim=(1, 2), unbiased=False)
        rstd = 1.0 / torch.sqrt(var + 1e-5)
        assert (st[:, 0] - mean).abs().max().item() <= 1e-5 * rms
        assert ((st[:, 1] - rstd) / rstd).abs().max().item() <= 1e-5
