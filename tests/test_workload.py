"""The roofline numerators: funcodec_b200.workload re-derives SURVEY.md §8(d)'s per-clip figures (probed there on the reference
modules with forward hooks) from the hyper-parameters alone, and bench.py's per-config constants equal them."""
import importlib.util
import os

import pytest

from funcodec_b200 import get_config
from funcodec_b200.workload import workload_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (preset, samples) -> SURVEY §8(d): conv MB, conv GMAC, LSTM GMAC, RVQ GFLOP @ n_q = 32, weight MB, whole path GMAC
SURVEY_8D = {
    ("encodec_16k_n32_ds640", 160000): (1066.6, 33.10, 8.39, 2.10, 230.2, 42.54),
    ("encodec_16k_n32_ds320", 160000): (780.1, 15.67, 4.19, 4.19, 59.4, None),
    ("encodec_16k_n32_ds320", 480000): (2340.1, 47.01, 12.58, 12.58, 59.4, 65.88),
    ("freqcodec_magphase_16k_n32_ds320", 160000): (803.2, 24.95, 4.20, 4.20, 64.9, 31.26),
    ("freqcodec_magphase_16k_n32_ds320_gr8", 160000): (803.2, 10.42, 4.20, 4.20, None, None),
}


@pytest.mark.parametrize("key", list(SURVEY_8D), ids=lambda k: f"{k[0]}-{k[1]}")
def test_model_reproduces_the_survey_figures(key):
    w = workload_model(get_config(key[0]), key[1])
    mb, gmac, lstm, rvq, wmb, total = SURVEY_8D[key]
    assert round(w["conv_bytes"] / 1e6, 1) == mb
    assert round(w["conv_macs"] / 1e9, 2) == gmac
    assert round(w["lstm_macs"] / 1e9, 2) == lstm
    assert round(w["rvq_flops"] / 1e9, 2) == rvq
    if wmb is not None:
        assert round(w["weight_bytes"] / 1e6, 1) == wmb
    if total is not None:
        assert round(w["total_macs"] / 1e9, 2) == total


def test_bench_constants_equal_the_model():
    spec = importlib.util.spec_from_file_location("bench_for_constants", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for name, c in bench.ALGO.items():
        w = workload_model(get_config(name), 160000)
        assert round(w["conv_bytes"] / 1e6, 1) == round(c["conv_bytes_per_10s"] / 1e6, 1), name
        assert round(w["conv_macs"] / 1e9, 2) == c["conv_gmac_per_10s"], name
        assert round(w["lstm_macs"] / 1e9, 2) == c["lstm_gmac_per_10s"], name
        assert round(w["rvq_flops"] / 1e9, 2) == c["rvq_gflop_per_10s_nq32"], name
        if not name.endswith("_gr8"):       # gr8: the engine reads zero-expanded dense weights (bench.py's comment)
            assert round(w["weight_bytes"] / 1e6, 1) == round(c["weight_bytes"] / 1e6, 1), name


def test_scaling_laws():
    """bytes and MACs are linear in the clip length (up to padding rows), the RVQ term in n_q; stacked residual blocks and the
    weight_norm parameter count follow the topology."""
    cfg = get_config("encodec_16k_n32_ds640")
    a, b = workload_model(cfg, 160000), workload_model(cfg, 320000)
    assert abs(b["conv_macs"] / a["conv_macs"] - 2) < 1e-9 and abs(b["conv_bytes"] / a["conv_bytes"] - 2) < 1e-4
    assert workload_model(cfg, 160000, n_q=8)["rvq_flops"] * 4 == a["rvq_flops"]
    ss = workload_model(get_config("soundstream_noncausal_16k_n32_ds320"), 160000)
    e3 = workload_model(get_config("encodec_16k_n32_ds320"), 160000)
    assert ss["lstm_macs"] == 0 and ss["conv_macs"] > 2 * e3["conv_macs"] - 16 * 1e9 and ss["frames"] == e3["frames"] == 500
    wn = workload_model(get_config("soundstream_16k_n32_ds320"), 160000)
    assert wn["conv_macs"] == ss["conv_macs"] and wn["conv_params"] < ss["conv_params"]      # g per channel vs gamma + beta


def test_stat_flops_line_quotes_the_per_second_totals():
    """4.25 GMAC (ds640) / 2.20 GMAC (ds320) per second of audio at n_q = 32."""
    from funcodec_b200.bin.codec_inference import stat_flops_line
    assert "model flops: 4.25G MACs per second" in stat_flops_line(get_config("encodec_16k_n32_ds640"), 16000)
    assert "model flops: 2.20G MACs per second" in stat_flops_line(get_config("encodec_16k_n32_ds320"))
    assert "RVQ@2 " in stat_flops_line(get_config("encodec_16k_n32_ds640"), 500)


def test_conv_launch_list_matches_the_engine_order_and_totals():
    """48 conv-kernel launches per ds640 round trip (24 per side, the four LSTM input GEMMs among them); without those four the
    per-launch bytes and MACs add up to the stack totals."""
    from funcodec_b200.workload import conv_launches
    cfg = get_config("encodec_16k_n32_ds640")
    ls = conv_launches(cfg, 160000)
    assert len(ls) == 48 and [l["name"] for l in ls[:5]] == ["enc.conv0", "enc.0.rb0.k3", "enc.0.rb0.1x1", "enc.0.rb0.shortcut", "enc.0.down"]
    assert ls[23]["name"] == "enc.final" and ls[24]["name"] == "dec.conv0" and ls[-1]["name"] == "dec.final"
    w = workload_model(cfg, 160000)
    core = [l for l in ls if ".lstm." not in l["name"]]
    assert len(core) == 44 and sum(l["macs"] for l in core) == w["conv_macs"] and sum(l["bytes"] for l in core) == w["conv_bytes"]
    assert sum(l["macs"] for l in ls if ".lstm." in l["name"]) * 2 == w["lstm_macs"]      # W_ih half of the LSTM work
