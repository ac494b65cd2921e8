"""Opt-in GPU tests for code paths written ahead of their hardware validation (run with FCB_EXPERIMENTAL=1).

`tc_stage`: the cp.async-staged producer mode of conv_tc.cu.  These tests replay the layer-level
and model-level parity checks with the option switched on; they are skipped in the default suite until the mode has been
validated and timed on a B200 (the shipped default is tc_stage = 0).
"""
import os

import numpy as np
import pytest
import torch

from funcodec_b200 import get_config, init_state_dict

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FCB_EXPERIMENTAL") != "1", reason="experimental paths: set FCB_EXPERIMENTAL=1")]


def test_stage_mode_conv_layers_1d():
    import test_gpu_layers as TL
    m = TL._models()
    model, sd = m["tc"], m["sd"]
    model.set_option("tc_stage", 1)
    try:
        for layer, cin, T, elu in TL.CASES:
            if cin % 32 != 0 and cin != 16:
                continue
            g = torch.Generator().manual_seed(1000 + T)
            x_btc = torch.randn(2, T, cin, generator=g)
            y, stats, row_off = model.debug_conv(layer, x_btc, elu=elu, want_stats=".lstm." not in layer)
            ref, ref_off, kept = TL._truth(sd, layer, x_btc.permute(0, 2, 1), elu)
            ref_btc = ref.permute(0, 2, 1)
            rms = ref_btc.pow(2).mean().sqrt().item()
            err = (y.cpu().double() - ref_btc).abs().max().item()
            print(f"stage {layer:28s} T={T:5d} rel err {err / rms:.3e}")
            assert err <= 2e-5 * rms, (layer, err, rms)
    finally:
        model.set_option("tc_stage", 0)


def test_stage_mode_conv_layers_2d(golden_dir):
    import test_gpu_freq as TF
    cfg, sd, model, _ = TF._full()
    model.set_option("tc_stage", 1)
    try:
        for layer, cin, F, T, elu in TF.CASES2:
            TF._check_conv2d_layer(model, sd, layer, cin, F, T, elu, 7)
    finally:
        model.set_option("tc_stage", 0)


def test_stage_mode_model_matches_default():
    """Model level: identical codes and <= 1e-5 waveform difference against the default producer mode (same math, only
    the way the rows reach the transform differs)."""
    from funcodec_b200.encodec import B200Encodec
    cfg = get_config("encodec_16k_n32_ds640")
    sd = init_state_dict(cfg, 0)
    a = B200Encodec(cfg, sd, "cuda:0")
    b = B200Encodec(cfg, sd, "cuda:0", options={"tc_stage": 1})
    g = torch.Generator().manual_seed(3)
    wav = 0.1 * torch.randn(3, 16000 * 2 + 77, generator=g)
    ra, rb = a.inference(wav), b.inference(wav)
    same = (ra["code_indices"][0] == rb["code_indices"][0]).all(dim=0).float().mean().item()
    assert same >= 0.999, same
    assert (ra["recon_speech"] - rb["recon_speech"]).abs().max().item() <= 1e-5 or same < 1.0


def test_small_cout_conv2d(golden_dir):
    """`conv2d_small_cout`: the halo-tile SIMT kernel for the 32 -> 3 (and the small model's 4 -> 3) output conv against the
    float64 truth, plus the config-4 architecture end to end with it."""
    import test_gpu_freq as TF
    cfg, sd, model, oracle = TF._full()
    zs, cfgs, sds, models, _ = TF._small(golden_dir)
    for m, s_, cases in ((model, sd, [("decoder.model.16", 32, 40, 150, True), ("decoder.model.16", 32, 257, 11, True),
                                      ("decoder.model.16", 32, 9, 333, False)]),):
        m.set_option("conv2d_small_cout", 1)
        try:
            for layer, cin, F, T, elu in cases:
                TF._check_conv2d_layer(m, s_, layer, cin, F, T, elu, 7)
        finally:
            m.set_option("conv2d_small_cout", 0)
    model.set_option("conv2d_small_cout", 1)
    try:
        g = torch.Generator().manual_seed(4)
        wav = 0.1 * torch.randn(2, 8000, generator=g)
        ora = oracle.inference(wav, want_margin=True)
        quant = ora["code_embeddings"][0][0]
        d = model.inference_decoding_emb(quant)
        dref = oracle.decode_frame(quant, None)
        assert (d["recon_speech"].cpu() - dref).abs().max().item() <= 1e-3
    finally:
        model.set_option("conv2d_small_cout", 0)


def test_stft_as_gemm(golden_dir):
    """`stft_tc`: STFT / iSTFT as tensor-core GEMMs against the oracle (torch.stft / istft), config-4 architecture and the
    small golden model, encoder output and decode-only waveform."""
    import test_gpu_freq as TF
    for getter in (TF._full, lambda: TF._small(golden_dir)[1:]):
        cfg, sd, model, oracle = getter()
        model.set_option("stft_tc", 1)
        try:
            g = torch.Generator().manual_seed(9)
            wav = 0.1 * torch.randn(2, 160 * 33 + 17, generator=g)
            ora = oracle.inference(wav, want_margin=True)
            r = model.inference(wav, need_recon=True, need_encoder_out=True, need_sub_quants=False)
            err = (r["encoder_out"].cpu() - ora["encoder_out"]).abs().max().item()
            print(f"stft_tc {cfg.name}: encoder_out max-abs err {err:.3e}")
            assert err <= 5e-5
            quant = ora["code_embeddings"][0][0]
            d = model.inference_decoding_emb(quant)
            dref = oracle.decode_frame(quant, None)
            derr = (d["recon_speech"].cpu() - dref).abs().max().item()
            print(f"stft_tc {cfg.name}: decode-only max-abs err {derr:.3e}")
            assert derr <= 1e-3
        finally:
            model.set_option("stft_tc", 0)


def test_lstm_prefetch_poll_is_bit_identical():
    """`lstm_prefetch_poll` only changes WHEN the loader warp reads a group's counter: outputs must be bit-identical."""
    from funcodec_b200.encodec import B200Encodec
    cfg = get_config("encodec_16k_n32_ds640")
    sd = init_state_dict(cfg, 0)
    a = B200Encodec(cfg, sd, "cuda:0")
    b = B200Encodec(cfg, sd, "cuda:0", options={"lstm_prefetch_poll": 1})
    g = torch.Generator().manual_seed(13)
    wav = 0.1 * torch.randn(19, 16000 + 321, generator=g)          # 3 clip groups, the last one partial
    ra, rb = a.inference(wav), b.inference(wav)
    assert torch.equal(ra["code_indices"][0], rb["code_indices"][0])
    assert torch.equal(ra["recon_speech"], rb["recon_speech"])


def test_m256_deep_layers():
    """`tc_m256`: the deep layers (C_in >= 256) on conv_tc_m256.cu (M = 256 rows per CTA, N = 64) against float64, and the
    model against the default kernels."""
    import test_gpu_layers as TL
    from funcodec_b200.encodec import B200Encodec
    m = TL._models()
    sd, cfg = m["sd"], m["cfg"]
    model = B200Encodec(cfg, sd, "cuda:0", options={"tc_m256": 1})
    for layer, cin, T, elu in TL.CASES + [("encoder.model.12", 256, 5 * 300 + 1, True), ("decoder.model.6", 512, 700, True)]:
        if cin < 256:
            continue
        g = torch.Generator().manual_seed(2000 + T)
        x_btc = torch.randn(2, T, cin, generator=g)
        y, stats, row_off = model.debug_conv(layer, x_btc, elu=elu, want_stats=".lstm." not in layer)
        ref, ref_off, kept = TL._truth(sd, layer, x_btc.permute(0, 2, 1), elu)
        ref_btc = ref.permute(0, 2, 1)
        rms = ref_btc.pow(2).mean().sqrt().item()
        err = (y.cpu().double() - ref_btc).abs().max().item()
        print(f"m256 {layer:28s} T={T:5d} rel err {err / rms:.3e}")
        assert err <= 2e-5 * rms, (layer, err, rms)
        if stats is not None:
            mean = ref.mean(dim=(1, 2))
            assert (stats.cpu().double()[:, 0] - mean).abs().max().item() <= 1e-5 * rms
    g = torch.Generator().manual_seed(3)
    wav = 0.1 * torch.randn(3, 16000 * 2 + 77, generator=g)
    ra, rb = m["tc"].inference(wav), model.inference(wav)
    same = (ra["code_indices"][0] == rb["code_indices"][0]).all(dim=0).float().mean().item()
    assert same >= 0.99, same


def test_grouped_freq_model(golden_dir):
    """conv_group_ratio / tr_conv_group_ratio > 0: the engine expands the grouped weights to dense block-diagonal matrices at
    pack time (engine.cu pack_conv2d / pack_convtr2d); against the vectors of the unmodified grouped reference model.
    (Opt-in until it has run on hardware; the expansion itself is checked on CPU in test_oracle_golden_freq.py.)"""
    from funcodec_b200.encodec import B200Encodec
    from oracle.freqcodec_oracle import OracleFreqCodec
    from parity_utils import assert_codes_parity
    z = np.load(os.path.join(golden_dir, "freq_magphase_small_grouped.npz"))
    cfg = get_config(str(z["cfg_name"]))
    sd = init_state_dict(cfg, int(z["seed"]))
    model = B200Encodec(cfg, sd, "cuda:0")
    wav = torch.from_numpy(z["wav"])
    ora = OracleFreqCodec(sd, list(zip(cfg.ratios_f, cfg.ratios))).inference(wav, want_margin=True)
    r = model.inference(wav, need_recon=True, need_encoder_out=True)
    assert np.abs(r["encoder_out"].cpu().numpy() - z["encoder_out"]).max() <= 5e-5
    res = assert_codes_parity(r["code_indices"][0].cpu().numpy(), z["codes"], ora["margins"].numpy(), 2e-3, min_exact_rate=0.9,
                              what="grouped")
    if not (res["first_stage"] >= 0).any():
        assert np.abs(r["recon_speech"].cpu().numpy() - z["recon"]).max() <= 1e-4


def test_freq_ds640_ratio_set(golden_dir):
    """conf/freqcodec_mag_phase_16k_n32_600k_step_ds640.yaml's ratio set (time strides 2, 1, 2, 1) against vectors of the
    unmodified reference (opt-in until it has run on hardware; same kernels as the ds320 set, different strides)."""
    from funcodec_b200.encodec import B200Encodec
    from oracle.freqcodec_oracle import OracleFreqCodec
    from parity_utils import assert_codes_parity
    z = np.load(os.path.join(golden_dir, "freq_magphase_small_ds640.npz"))
    cfg = get_config(str(z["cfg_name"]))
    sd = init_state_dict(cfg, int(z["seed"]))
    model = B200Encodec(cfg, sd, "cuda:0")
    wav = torch.from_numpy(z["wav"])
    ora = OracleFreqCodec(sd, list(zip(cfg.ratios_f, cfg.ratios))).inference(wav, want_margin=True)
    r = model.inference(wav, need_recon=True, need_encoder_out=True)
    assert np.abs(r["encoder_out"].cpu().numpy() - z["encoder_out"]).max() <= 5e-5
    res = assert_codes_parity(r["code_indices"][0].cpu().numpy(), z["codes"], ora["margins"].numpy(), 2e-3, min_exact_rate=0.9,
                              what="ds640 ratios")
    assert tuple(r["recon_speech"].shape) == tuple(z["recon"].shape)
    if not (res["first_stage"] >= 0).any():
        assert np.abs(r["recon_speech"].cpu().numpy() - z["recon"]).max() <= 1e-4
