"""Host plumbing (funcodec_b200/pipeline.py) against the reference's formats -- CPU only."""
import json
import os

import numpy as np
import torch

from funcodec_b200 import pipeline as P


def test_wrap_pad_matches_numpy_wrap():
    clips = [np.arange(5, dtype=np.float32), np.arange(12, dtype=np.float32), np.arange(1, dtype=np.float32) + 7]
    x, lens = P.wrap_pad_batch(clips)
    assert x.shape == (3, 12) and lens.tolist() == [5, 12, 1]
    assert x[0].tolist() == [0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 0, 1]       # nets_utils.pad_list_with_mod(mode="wrap")
    assert x[2].tolist() == [7.0] * 12


def test_codecs_txt_roundtrip_matches_reference_loader():
    g = torch.Generator().manual_seed(0)
    codes = torch.randint(0, 1024, (4, 3, 9), generator=g)             # [n_q, B, T']
    line = P.format_indices_line("utt_01", [codes], batch_id=1, length=7)
    key, js = line.strip().split(" ", 1)
    assert key == "utt_01" and json.loads(js) == [codes[:, 1, :7].tolist()]
    k2, arr = P.parse_indices_line(line)                                # == load_codec_json: [T', n_q]
    assert k2 == "utt_01" and arr.shape == (7, 4) and np.array_equal(arr, codes[:, 1, :7].numpy().T)


def test_peak_limit_and_wav_io(tmp_path):
    w = torch.tensor([[0.5, -2.0, 1.0]])
    assert torch.allclose(P.peak_limit(w, rescale=True), w * (0.99 / 2.0))
    assert P.peak_limit(w, rescale=False).abs().max().item() <= 0.99 + 1e-6
    path = os.path.join(tmp_path, "a.wav")
    x = torch.sin(torch.arange(1600) / 10.0) * 0.5
    P.save_wav_pcm16(path, x.view(1, -1), 16000, rescale=True)
    y, sr = P.load_wav(path)
    assert sr == 16000 and np.abs(y - x.numpy()).max() <= 1.0 / 32768 + 1e-6


def test_segment_plan_matches_the_reference_arithmetic():
    """fcb_plan_segments_for_hop (pure host arithmetic of the C ABI) against Encodec._encode's offsets / lengths
    (codec_basic.py:346-358) and the condition under which the reference's _linear_overlap_add raises (:101-112)."""
    import ctypes
    import random
    import torch
    from funcodec_b200 import _capi
    from oracle import encodec_oracle as O
    lib = _capi.load_library()
    rnd = random.Random(7)
    seen_ok = seen_bad = 0
    for _ in range(400):
        hop = rnd.choice([40, 320, 640])
        seg = rnd.randint(hop // 2, 12 * hop)
        stride = max(1, int((1 - rnd.choice([0.0, 0.01, 0.1, 0.25, 0.5, 0.75])) * seg))
        L = rnd.randint(1, 40 * hop)
        offsets = list(range(0, L, stride))
        lens = [min(seg, L - o) for o in offsets]
        dec = [-(-n // hop) * hop for n in lens]
        n_tail = sum(1 for n in lens if n < seg)
        try:
            O.linear_overlap_add([torch.zeros(1, 1, d) + 1 for d in dec], stride)
            ref_ok = True
        except RuntimeError:
            ref_ok = False
        plan = _capi.FcbSegmentPlan()
        rc = lib.fcb_plan_segments_for_hop(hop, L, seg, stride, ctypes.byref(plan))
        if n_tail > _capi.FCB_MAX_TAIL_SEGMENTS:
            assert rc != 0
            continue
        assert (rc == 0) == ref_ok, (hop, seg, stride, L, rc, ref_ok)
        if rc != 0:
            seen_bad += 1
            continue
        seen_ok += 1
        assert plan.n_seg == len(offsets) and plan.n_full == len(offsets) - n_tail and plan.n_tail == n_tail
        assert plan.frames_full * hop == -(-seg // hop) * hop == plan.decoded_full
        assert [plan.tail_len[i] for i in range(n_tail)] == lens[plan.n_full:]
        assert [plan.tail_frames[i] * hop for i in range(n_tail)] == dec[plan.n_full:]
        assert plan.total_frames == sum(d // hop for d in dec)
    assert seen_ok > 50 and seen_bad > 10, (seen_ok, seen_bad)


def test_packed_codes_format_roundtrip(tmp_path):
    """funcodec_b200/codes_format.py: 10-bit packed container <-> codecs.txt, bit-exact both ways, ragged
    lengths, several codebook sizes, LSB-first packing order, corruption is detected."""
    import io
    from funcodec_b200 import codes_format as CF
    rng = np.random.default_rng(0)
    assert CF.bits_for_codebook(1024) == 10 and CF.bits_for_codebook(1000) == 10 and CF.bits_for_codebook(2) == 1
    # known answer: indices 1 and 2 at 10 bits, LSB first -> 0x01, 0x08 (bit 11), 0x00
    assert CF.pack_indices(np.array([[1, 2]]), 10) == bytes([0x01, 0x08, 0x00])
    for K in (1024, 64, 2, 4096, 1000):
        buf = io.BytesIO()
        items = []
        for i, (nq, T) in enumerate([(32, 250), (8, 1), (1, 17), (4, 0)]):
            c = rng.integers(0, K, size=(nq, T))
            items.append((f"utt-{K}-{i}", c))
            CF.write_record(buf, items[-1][0], c, K)
        buf.seek(0)
        got = list(CF.read_records(buf))
        assert [k for k, _ in got] == [k for k, _ in items]
        for (_, a), (_, b) in zip(got, items):
            assert a.shape == b.shape and np.array_equal(a, b)
    # codecs.txt -> packed -> codecs.txt is the identity on the parsed arrays; 10 bits/index vs JSON
    txt = os.path.join(tmp_path, "codecs.txt")
    codes = torch.from_numpy(rng.integers(0, 1024, size=(32, 3, 250)))
    with open(txt, "wt") as f:
        for b in range(3):
            f.write(P.format_indices_line(f"u{b}", [codes], b, 250 - 10 * b))
    n, tbytes, pbytes = CF.codecs_txt_to_packed(txt, os.path.join(tmp_path, "codes.fcb"))
    assert n == 3 and pbytes < tbytes / 3
    assert CF.packed_to_codecs_txt(os.path.join(tmp_path, "codes.fcb"), os.path.join(tmp_path, "back.txt")) == 3
    a = [P.parse_indices_line(l) for l in open(txt)]
    b = [P.parse_indices_line(l) for l in open(os.path.join(tmp_path, "back.txt"))]
    assert all(ka == kb and np.array_equal(xa, xb) for (ka, xa), (kb, xb) in zip(a, b))
    import pytest
    with pytest.raises(ValueError):
        list(CF.read_records(io.BytesIO(b"XXXX" + bytes(8))))
    with pytest.raises(ValueError):
        CF.pack_indices(np.array([[1024]]), 10)
    raw = open(os.path.join(tmp_path, "codes.fcb"), "rb").read()
    with pytest.raises(ValueError):
        list(CF.read_records(io.BytesIO(raw[:-3])))
