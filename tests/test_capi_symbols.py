"""The C-ABI library loads on a CPU-only box and exports every symbol include/funcodec_b200.h declares
(no compute calls without a GPU); host-side config logic matches the reference formulas."""
import ctypes
import os
import re

import pytest

from funcodec_b200 import _capi, get_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "funcodec_b200.h")).read()
    declared = set(re.findall(r"FCB_API\s+[\w\s\*]+?\b(fcb_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_capi.SYMBOLS), (declared ^ set(_capi.SYMBOLS))
    lib = _capi.load_library()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.fcb_version().startswith(b"funcodec_b200")


def test_create_without_gpu_fails_loudly_not_silently():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from funcodec_b200.encodec import B200Encodec
    from funcodec_b200 import init_state_dict
    cfg = get_config("tiny_ds40")
    with pytest.raises(Exception):
        B200Encodec(cfg, init_state_dict(cfg, 0), "cuda:0")
    with pytest.raises(_capi.FcbError):
        B200Encodec(cfg, init_state_dict(cfg, 0), "cpu")


def test_bandwidth_to_quantizers():
    """vq.py:105-117 and codec_inference.py:121-125."""
    cfg = get_config("encodec_16k_n32_ds640")
    assert cfg.bandwidth_per_quantizer() == 250.0
    assert [cfg.num_quantizers_for_bandwidth(b) for b in (None, 0, 250, 499, 4000, 8000, 99999)] == [32, 32, 1, 1, 16, 32, 399]
    cfg = get_config("encodec_16k_n32_ds320")
    assert cfg.bandwidth_per_quantizer() == 500.0 and cfg.hop_length == 320 and cfg.frames(480000) == 1500


def test_config_from_reference_like_model():
    """integration.config_from_reference_model on a duck-typed stand-in for the reference Encodec module."""
    import types
    import torch
    from funcodec_b200 import init_state_dict
    from funcodec_b200.integration import config_from_reference_model, UnsupportedReferenceModel
    cfg = get_config("encodec_16k_n32_ds320")
    sd = init_state_dict(cfg, 0)
    m = types.SimpleNamespace(
        encoder=types.SimpleNamespace(ratios=list(reversed(cfg.ratios))), decoder=types.SimpleNamespace(ratios=list(cfg.ratios)),
        quantizer=types.SimpleNamespace(sampling_rate=16000, encoder_hop_length=320, codebook_size=1024, input_proj=None, input_act=None),
        audio_normalize=True, segment_dur=None, codec_domain="time", state_dict=lambda: sd)
    got = config_from_reference_model(m)
    for f in ("ratios", "n_filters", "dimension", "kernel_size", "last_kernel_size", "residual_kernel_size", "lstm_layers",
              "codebook_size", "num_quantizers", "sample_rate", "audio_normalize"):
        assert getattr(got, f) == getattr(cfg, f), f
    m.segment_dur = 1.0                     # segmenting is handled by the wrapper (fcb_roundtrip_segmented), not the config
    assert config_from_reference_model(m).ratios == cfg.ratios
    m.quantizer.input_proj = object()
    with pytest.raises(UnsupportedReferenceModel):
        config_from_reference_model(m)


def test_config_from_reference_like_freqcodec():
    """Same, for a mag_phase FreqCodec (codec_freq.py:118-215): [freq, time] ratio pairs, domain_conf."""
    import types
    from funcodec_b200 import init_state_dict
    from funcodec_b200.integration import config_from_reference_model, UnsupportedReferenceModel
    cfg = get_config("freq_small")
    sd = init_state_dict(cfg, 0)
    pairs = [[f, t] for f, t in zip(cfg.ratios_f, cfg.ratios)]
    m = types.SimpleNamespace(
        encoder=types.SimpleNamespace(ratios=list(reversed(pairs))), decoder=types.SimpleNamespace(ratios=pairs),
        quantizer=types.SimpleNamespace(sampling_rate=cfg.sample_rate, encoder_hop_length=cfg.hop_length,
                                        codebook_size=cfg.codebook_size, input_proj=None, input_act=None),
        audio_normalize=True, segment_dur=None, codec_domain=["mag_phase", "mag_phase"],
        domain_conf={"n_fft": cfg.n_fft, "hop_length": cfg.stft_hop}, state_dict=lambda: sd)
    got = config_from_reference_model(m)
    for f in ("arch", "ratios", "ratios_f", "n_fft", "stft_hop", "n_filters", "dimension", "kernel_size", "last_kernel_size",
              "residual_kernel_size", "lstm_layers", "codebook_size", "num_quantizers"):
        assert getattr(got, f) == getattr(cfg, f), f
    assert (got.conv_group_ratio, got.tr_conv_group_ratio) == (-1, -1)
    # grouped 2-D convs: the ratios are recovered from the weight shapes
    gcfg = get_config("freq_small_grouped")
    gsd = init_state_dict(gcfg, 0)
    m.state_dict = lambda: gsd
    got = config_from_reference_model(m)
    assert (got.conv_group_ratio, got.tr_conv_group_ratio, got.n_filters) == (gcfg.conv_group_ratio, gcfg.tr_conv_group_ratio, gcfg.n_filters)
    m.state_dict = lambda: sd
    m.segment_dur = 1.0
    with pytest.raises(UnsupportedReferenceModel):
        config_from_reference_model(m)
    m.segment_dur = None
    m.codec_domain = ["stft", "stft"]
    with pytest.raises(UnsupportedReferenceModel):
        config_from_reference_model(m)
