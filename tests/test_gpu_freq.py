"""GPU parity for the FreqCodec mag_phase variant (BASELINE config 4) through the C ABI:
golden vectors from the unmodified reference + the CPU oracle (oracle/freqcodec_oracle.py) on seeded inputs."""
import os

import numpy as np
import pytest
import torch

from funcodec_b200 import get_config, init_state_dict
from oracle.freqcodec_oracle import OracleFreqCodec
from parity_utils import assert_codes_parity

pytestmark = pytest.mark.gpu

WAV_TOL = 1e-4
EMB_TOL = 5e-5
MARGIN = 2e-3
_M = {}


def _small(golden_dir):
    from funcodec_b200.encodec import B200Encodec
    if "small" not in _M:
        z = np.load(os.path.join(golden_dir, "freq_magphase_small.npz"))
        sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
        cfg = get_config("freq_small")
        _M["small"] = (z, cfg, sd, B200Encodec(cfg, sd, "cuda:0"), OracleFreqCodec(sd, list(zip(cfg.ratios_f, cfg.ratios))))
    return _M["small"]


def test_freq_golden(golden_dir):
    z, cfg, sd, model, oracle = _small(golden_dir)
    wav = torch.from_numpy(z["wav"])
    ora = oracle.inference(wav, want_margin=True)
    r = model.inference(wav, need_recon=True, need_encoder_out=True)
    enc = r["encoder_out"].cpu().numpy()
    assert enc.shape == z["encoder_out"].shape
    assert np.abs(enc - z["encoder_out"]).max() <= EMB_TOL, np.abs(enc - z["encoder_out"]).max()
    codes = r["code_indices"][0].cpu().numpy()
    res = assert_codes_parity(codes, z["codes"], ora["margins"].numpy(), MARGIN, min_exact_rate=0.9, what="freq golden")
    ok_clip = ~(res["first_stage"] >= 0).any(axis=1)
    recon = r["recon_speech"].cpu().numpy()
    assert recon.shape == z["recon"].shape
    for b in np.nonzero(ok_clip)[0]:
        assert np.abs(recon[b] - z["recon"][b]).max() <= WAV_TOL, np.abs(recon[b] - z["recon"][b]).max()
    # decode-only parity from the reference's quantized embeddings
    d = model.inference_decoding_emb(torch.from_numpy(z["quant"]))
    dref = oracle.decode_frame(torch.from_numpy(z["quant"]), None)
    assert tuple(d["recon_speech"].shape) == tuple(dref.shape)
    assert (d["recon_speech"].cpu() - dref).abs().max().item() <= WAV_TOL * 10      # un-scaled output (~10x amplitude)


@pytest.mark.parametrize("B,L", [(3, 160 * 21 + 5), (1, 160 * 40), (2, 160 * 37 + 159)])
def test_freq_seeded(golden_dir, B, L):
    """Ragged lengths incl. the case where the iSTFT yields fewer than L samples (even STFT frame count)."""
    z, cfg, sd, model, oracle = _small(golden_dir)
    g = torch.Generator().manual_seed(B * 1000 + L)
    wav = 0.1 * torch.randn(B, L, generator=g)
    ora = oracle.inference(wav, want_margin=True)
    r = model.inference(wav, need_recon=True, need_encoder_out=True)
    assert tuple(r["recon_speech"].shape) == tuple(ora["recon_speech"].shape)
    assert (r["encoder_out"].cpu() - ora["encoder_out"]).abs().max().item() <= EMB_TOL
    res = assert_codes_parity(r["code_indices"][0].cpu().numpy(), ora["code_indices"][0].numpy(), ora["margins"].numpy(),
                              MARGIN, min_exact_rate=0.9, what="freq seeded")
    ok_clip = ~(res["first_stage"] >= 0).any(axis=1)
    for b in np.nonzero(ok_clip)[0]:
        assert (r["recon_speech"][b].cpu() - ora["recon_speech"][b]).abs().max().item() <= WAV_TOL


def test_freq_full_config_shapes_and_oracle_spot_check():
    """BASELINE config 4 architecture (repo YAML, groups = 1) on a short clip against the oracle."""
    from funcodec_b200.encodec import B200Encodec
    cfg = get_config("freqcodec_magphase_16k_n32_ds320")
    sd = init_state_dict(cfg, 0)
    model = B200Encodec(cfg, sd, "cuda:0")
    oracle = OracleFreqCodec(sd, list(zip(cfg.ratios_f, cfg.ratios)))
    g = torch.Generator().manual_seed(4)
    wav = 0.1 * torch.randn(2, 8000, generator=g)
    ora = oracle.inference(wav, want_margin=True)
    r = model.inference(wav, need_recon=True, need_encoder_out=True, need_sub_quants=False)
    assert tuple(r["code_indices"][0].shape) == (32, 2, 26)
    assert (r["encoder_out"].cpu() - ora["encoder_out"]).abs().max().item() <= EMB_TOL
    res = assert_codes_parity(r["code_indices"][0].cpu().numpy(), ora["code_indices"][0].numpy(), ora["margins"].numpy(),
                              MARGIN, min_exact_rate=0.9, what="freq full")
    ok_clip = ~(res["first_stage"] >= 0).any(axis=1)
    for b in np.nonzero(ok_clip)[0]:
        assert (r["recon_speech"][b].cpu() - ora["recon_speech"][b]).abs().max().item() <= WAV_TOL
