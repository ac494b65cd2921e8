"""Pins the FreqCodec (mag_phase, BASELINE config 4) oracle against vectors produced by the UNMODIFIED reference
(tools/gen_golden_freq.py).  CPU only.  The CUDA path for this variant is round-2 work; the oracle is ready for it."""
import os

import numpy as np
import torch

from oracle.freqcodec_oracle import OracleFreqCodec


def test_freqcodec_magphase_oracle_vs_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "freq_magphase_small.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    o = OracleFreqCodec(sd, [tuple(r) for r in z["ratios"]])
    wav = torch.from_numpy(z["wav"])
    r = o.inference(wav, want_margin=True)
    assert r["features"].shape[1:3] == (3, 257)
    assert np.abs(r["encoder_out"].numpy() - z["encoder_out"]).max() <= 5e-6
    assert np.array_equal(r["code_indices"][0].numpy(), z["codes"].astype(np.int64))
    assert np.abs(r["code_embeddings"][0][0].numpy() - z["quant"]).max() <= 5e-6
    assert np.abs(r["code_embeddings"][0][1].numpy() - z["scale"]).max() <= 1e-7
    assert r["recon_speech"].shape == z["recon"].shape
    assert np.abs(r["recon_speech"].numpy() - z["recon"]).max() <= 5e-6


def test_freqcodec_config4_architecture_oracle_vs_reference(golden_dir):
    """BASELINE config 4's architecture (n_filters 32, D 128, K 1024, n_q 32, conv groups = 1) on a 0.5 s clip: the oracle
    against the unmodified reference FreqCodec (tools/gen_golden_freq.py); weights = init_state_dict(cfg, 0)."""
    from funcodec_b200 import get_config, init_state_dict
    z = np.load(os.path.join(golden_dir, "freq_magphase_config4_arch.npz"))
    cfg = get_config(str(z["cfg_name"]))
    sd = init_state_dict(cfg, int(z["seed"]))
    assert abs(float(sum(v.double().abs().sum().item() for v in sd.values())) - float(z["sd_checksum"])) <= 1e-6 * float(z["sd_checksum"])
    o = OracleFreqCodec(sd, list(zip(cfg.ratios_f, cfg.ratios)))
    r = o.inference(torch.from_numpy(z["wav"]), want_margin=True)
    assert np.abs(r["encoder_out"].numpy() - z["encoder_out"]).max() <= 1e-5
    assert np.array_equal(r["code_indices"][0].numpy(), z["codes"].astype(np.int64))
    assert np.abs(r["code_embeddings"][0][0].numpy() - z["quant"]).max() <= 1e-5
    assert r["recon_speech"].shape == z["recon"].shape
    assert np.abs(r["recon_speech"].numpy() - z["recon"]).max() <= 1e-5


def test_freq_layers_oracle_vs_reference_modules(golden_dir):
    """sconv2d / sconvtr2d / resblock2d against the reference MODULES SConv2d / SConvTranspose2d / SEANetResnetBlock2d
    (tools/gen_golden_freq_layers.py): every kernel / stride family, odd lengths, out_padding."""
    from oracle import freqcodec_oracle as FO
    z = np.load(os.path.join(golden_dir, "freq_layers.npz"))

    def sd_of(prefix):
        return {"L." + k[len(prefix) + 4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix + ".sd.")}

    n = 0
    while f"conv{n}.x" in z.files:
        cin, cout, kf, kt, sf, st = [int(v) for v in z[f"conv{n}.meta"]]
        y = FO.sconv2d(torch.from_numpy(z[f"conv{n}.x"]), sd_of(f"conv{n}"), "L", stride=(sf, st))
        assert y.shape == z[f"conv{n}.y"].shape, (n, y.shape)
        assert np.abs(y.numpy() - z[f"conv{n}.y"]).max() <= 2e-6, n
        n += 1
    assert n == 6
    n = 0
    while f"convtr{n}.x" in z.files:
        cin, cout, sf, st, fl, fr, tl, tr = [int(v) for v in z[f"convtr{n}.meta"]]
        y = FO.sconvtr2d(torch.from_numpy(z[f"convtr{n}.x"]), sd_of(f"convtr{n}"), "L", (sf, st), ((fl, fr), (tl, tr)))
        assert y.shape == z[f"convtr{n}.y"].shape, (n, y.shape, z[f"convtr{n}.y"].shape)
        assert np.abs(y.numpy() - z[f"convtr{n}.y"]).max() <= 2e-6, n
        n += 1
    assert n == 3
    y = FO.resblock2d(torch.from_numpy(z["rb.x"]), sd_of("rb"), "L")
    assert np.abs(y.numpy() - z["rb.y"]).max() <= 2e-6
