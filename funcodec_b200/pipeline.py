"""Host-side plumbing around the hot path: what `funcodec/bin/codec_inference.py` does between
the scp files and the model -- wrap-padded variable-length batches, the `codecs.txt` JSONL format, peak-limited PCM16
output -- so that `encoding_decoding.sh`-style runs work end to end on top of B200Encodec.  Pure host logic (numpy /
stdlib); no CUDA here.

References: collate with pad_mode="wrap" `funcodec/bin/codec_inference.py:257-261` -> `funcodec/modules/nets_utils.py:65-98`;
`write_indices` `codec_inference.py:288-299`; `load_codec_json` `funcodec/datasets/iterable_dataset.py:54-58`;
per-utterance trimming `codec_inference.py:358-367`; `save_audio` `:153-161`.
"""
import json
import os
import wave
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch


def wrap_pad_batch(clips: Sequence[np.ndarray]) -> Tuple[torch.Tensor, torch.Tensor]:
    """pad_list_with_mod(mode="wrap"): short clips are extended by repeating themselves from the start; RMS and
    GroupNorm then run over the padded length (as in the reference) and outputs are trimmed by `lengths`."""
    lengths = torch.tensor([int(c.shape[0]) for c in clips], dtype=torch.int64)
    max_len = int(lengths.max())
    out = [np.pad(np.asarray(c, dtype=np.float32), (0, max_len - c.shape[0]), mode="wrap") for c in clips]
    return torch.from_numpy(np.stack(out, axis=0)), lengths


def format_indices_line(key: str, code_indices: List[torch.Tensor], batch_id: int, length: int) -> str:
    """`write_indices`: one line `uttid [[[q0 codes...],[q1 codes...],...]]` (n_frame x n_q x T', n_frame == 1)."""
    to_write = [x[:, batch_id, :length].cpu().numpy().tolist() for x in code_indices]
    return key + " " + json.dumps(to_write) + "\n"


def parse_indices_line(line: str) -> Tuple[str, np.ndarray]:
    """Inverse of format_indices_line via the reference's load_codec_json: returns (key, codes [T', n_q])."""
    key, json_str = line.strip().split(" ", 1)
    array = np.array(json.loads(json_str))
    if array.ndim == 3:
        array = array[0]
    return key, array.T


def peak_limit(wav: torch.Tensor, rescale: bool = True, limit: float = 0.99) -> torch.Tensor:
    """`save_audio`'s amplitude handling: rescale so that the peak is <= 0.99, or clamp."""
    mx = wav.abs().max()
    if rescale:
        return wav * min(limit / float(mx), 1.0) if float(mx) > 0 else wav
    return wav.clamp(-limit, limit)


def save_wav_pcm16(path: str, wav: torch.Tensor, sample_rate: int, rescale: bool = True) -> None:
    """16-bit PCM mono wav (the reference calls torchaudio.save(encoding='PCM_S', bits_per_sample=16))."""
    x = peak_limit(wav.detach().cpu().float().reshape(-1), rescale).numpy()
    pcm = np.clip(np.round(x * 32768.0), -32768, 32767).astype("<i2")
    with wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(sample_rate)
        f.writeframes(pcm.tobytes())


def load_wav(path: str) -> Tuple[np.ndarray, int]:
    with wave.open(path, "rb") as f:
        assert f.getsampwidth() == 2, "only 16-bit PCM is supported by this loader"
        sr, n, ch = f.getframerate(), f.getnframes(), f.getnchannels()
        data = np.frombuffer(f.readframes(n), dtype="<i2").astype(np.float32) / 32768.0
    if ch > 1:
        data = data.reshape(-1, ch).mean(axis=1)
    return data, sr


def read_scp(path: str) -> List[Tuple[str, str]]:
    out = []
    with open(path, "rt", encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            if line:
                k, v = line.split(maxsplit=1)
                out.append((k, v))
    return out


def batches(items: Sequence, batch_size: int) -> Iterator[Sequence]:
    for i in range(0, len(items), batch_size):
        yield items[i: i + batch_size]


def select_keys(items: Sequence[Tuple[str, object]], key_file: Optional[str]) -> List[Tuple[str, object]]:
    """`--key_file` (codec_inference.py:479 -> IterableESPnetDataset): iterate over the keys of `key_file` (first column),
    each looked up in the data scp -- this is how `encoding_decoding.sh:79-86` shards one wav.scp over JOB processes."""
    if not key_file:
        return list(items)
    table = dict(items)
    out = []
    with open(key_file, "rt", encoding="utf-8") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            k = line.split(maxsplit=1)[0]
            if k not in table:
                raise KeyError(f"key_file entry {k!r} is not in the data file")
            out.append((k, table[k]))
    return out


class IndicesWriter:
    """`write_indices` (codec_inference.py:277-299): `codecs.txt` JSONL (`--indices_save_type text`) or a Kaldi
    `indices.ark/.scp` float matrix [T', n_q] per utterance (`ark`).  Disabled when `--need_indices` is false."""

    def __init__(self, output_dir: str, enabled: bool, save_type: str = "text"):
        self.fout, self.ark = None, None
        if not enabled:
            return
        if save_type == "ark":
            from .kaldi_io import ArkScpWriter
            self.ark = ArkScpWriter(os.path.join(output_dir, "indices"))
        else:
            self.fout = open(os.path.join(output_dir, "codecs.txt"), "wt")

    def write(self, key: str, code_indices: List[torch.Tensor], batch_id: int, length: int) -> None:
        if self.ark is not None:
            mats = [x[:, batch_id, :length].cpu().float().numpy().T for x in code_indices]
            self.ark(key, np.concatenate(mats, axis=0))
        elif self.fout is not None:
            self.fout.write(format_indices_line(key, code_indices, batch_id, length))

    def close(self) -> None:
        if self.ark is not None:
            self.ark.close()
        if self.fout is not None:
            self.fout.close()


def sub_quants_matrix(sub_quants: List[torch.Tensor], batch_id: int, length: int) -> np.ndarray:
    """`write_sub_quants` (codec_inference.py:301-311): list of [n_q, B, D, T'] -> [T', n_q * D]."""
    x = torch.cat(sub_quants, dim=-1).permute(1, 3, 0, 2)[batch_id][:length]
    return x.reshape(x.shape[0], -1).cpu().numpy()


def _wav_name(output_dir: str, key: str) -> str:
    return os.path.join(output_dir, key if key.endswith(".wav") else key + ".wav")


def run_encode(s2t, wav_scp: str, output_dir: str, batch_size: int = 16, bit_width: Optional[int] = None,
               run_mod: str = "encode", use_scale: bool = True, save_recon: Optional[bool] = None,
               key_file: Optional[str] = None, need_indices: bool = True, indices_save_type: str = "text",
               need_sub_quants: bool = False) -> int:
    """`--run_mod encode|inference`: wav.scp -> codecs.txt / indices.ark (+ codec_emb.ark, + reconstructed wavs in
    `inference` mode).  Returns the number of utterances."""
    os.makedirs(output_dir, exist_ok=True)
    hop = s2t.model.quantizer.encoder_hop_length
    sr_model = s2t.model.quantizer.sampling_rate
    if save_recon is None:
        save_recon = run_mod == "inference"
    writer = IndicesWriter(output_dir, need_indices, indices_save_type)
    sq_writer = None
    if need_sub_quants:
        from .kaldi_io import ArkScpWriter
        sq_writer = ArkScpWriter(os.path.join(output_dir, "codec_emb"))
    n = 0
    try:
        for group in batches(select_keys(read_scp(wav_scp), key_file), batch_size):
            clips = []
            for key, path in group:
                x, sr = load_wav(path)
                if sr != sr_model:
                    raise ValueError(f"{key}: sample rate {sr} != model rate {sr_model} (resampling is out of scope)")
                clips.append(x)
            speech, lengths = wrap_pad_batch(clips)
            codes, _, recon, sub = s2t(speech, need_recon=True, bit_width=bit_width, use_scale=use_scale, run_mod=run_mod)
            for i, (key, _) in enumerate(group):
                ilen = int(lengths[i])
                codec_len = -(-ilen // hop)
                if save_recon and recon is not None:
                    save_wav_pcm16(_wav_name(output_dir, key), recon[i].cpu()[:, :ilen], sr_model, rescale=True)
                if codes is not None:
                    writer.write(key, codes, i, codec_len)
                if sq_writer is not None and sub is not None and sub[0] is not None:
                    sq_writer(key, sub_quants_matrix(sub, i, codec_len))
                n += 1
    finally:
        writer.close()
        if sq_writer is not None:
            sq_writer.close()
    return n


def _decode_groups(s2t, items, output_dir, batch_size, bit_width, run_mod, to_tensor) -> int:
    hop = s2t.model.quantizer.encoder_hop_length
    sr_model = s2t.model.quantizer.sampling_rate
    n = 0
    for group in batches(items, batch_size):
        lens = [c.shape[0] for _, c in group]
        tmax = max(lens)
        toks = np.stack([np.pad(c, ((0, tmax - c.shape[0]), (0, 0)), mode="wrap") for _, c in group], axis=0)
        _, _, recon, _ = s2t(to_tensor(toks), bit_width=bit_width, run_mod=run_mod)
        for i, (key, _) in enumerate(group):
            save_wav_pcm16(_wav_name(output_dir, key), recon[i].cpu()[:, : lens[i] * hop], sr_model, rescale=True)
            n += 1
    return n


def run_decode(s2t, codecs_txt: str, output_dir: str, batch_size: int = 16, bit_width: Optional[int] = None,
               key_file: Optional[str] = None) -> int:
    """`--run_mod decode`: codecs.txt (`codec_json`) -> wavs.  Code sequences of different lengths are wrap-padded like the
    reference's collate (int arrays) and the output is trimmed to codec_len * hop (codec_inference.py:358-361)."""
    os.makedirs(output_dir, exist_ok=True)
    with open(codecs_txt, "rt") as f:
        items = [parse_indices_line(line) for line in f if line.strip()]
    items = select_keys(items, key_file)
    return _decode_groups(s2t, items, output_dir, batch_size, bit_width, "decode",
                          lambda a: torch.from_numpy(a.astype(np.int64)))


def run_decode_emb(s2t, emb_scp: str, output_dir: str, batch_size: int = 16, key_file: Optional[str] = None) -> int:
    """`--run_mod decode_emb` (codec_inference.py:118-119; encoding_decoding.sh stage 3): a Kaldi scp of [T', D] embedding
    matrices (`kaldi_ark`) -> wavs through `inference_decoding_emb`."""
    from .kaldi_io import read_mat
    os.makedirs(output_dir, exist_ok=True)
    items = [(k, read_mat(spec)) for k, spec in select_keys(read_scp(emb_scp), key_file)]
    return _decode_groups(s2t, items, output_dir, batch_size, None, "decode_emb",
                          lambda a: torch.from_numpy(a.astype(np.float32)))


def load_items(path: str, dtype: str, key_file: Optional[str] = None) -> List[Tuple[str, np.ndarray]]:
    """The part of `build_streaming_iterator` (codec_inference.py:249-264) this path uses: one data file of type `sound`
    (wav.scp -> float32 samples), `codec_json` / `text` (codecs.txt -> int codes [T', n_q]) or `kaldi_ark` (scp -> float
    matrices [T', D]), restricted and ordered by `key_file`."""
    if dtype == "sound":
        return [(k, load_wav(pth)) for k, pth in select_keys(read_scp(path), key_file)]
    if dtype in ("codec_json", "text"):
        with open(path, "rt") as f:
            return select_keys([parse_indices_line(line) for line in f if line.strip()], key_file)
    if dtype == "kaldi_ark":
        from .kaldi_io import read_mat
        return [(k, read_mat(spec)) for k, spec in select_keys(read_scp(path), key_file)]
    raise ValueError(f"data type {dtype!r} is not read by this path (sound, codec_json, kaldi_ark)")


def forward_items(s2t, items: Sequence[Tuple[str, np.ndarray]], output_path: Optional[str], batch_size: int = 1,
                  bit_width: Optional[int] = None, use_scale: bool = True, run_mod: str = "inference",
                  need_indices: bool = False, indices_save_type: str = "text", need_sub_quants: bool = False,
                  sample_rate: Optional[int] = None, file_sample_rate: Optional[int] = None) -> List[Dict]:
    """The batch loop of `inference_modelscope._forward` (codec_inference.py:313-381) for any run_mod: wrap-padded batches through
    `s2t`, per-utterance trimming (decode modes: codec_len * hop samples; else the input length and ceil(len / hop) frames), then
    either files under `output_path` (wav + codecs.txt / indices.ark + codec_emb.ark as requested) or -- with no output path -- the
    reference's in-memory result list [{"key": uttid, "value": recon_wav [1, L] or None}].

    `file_sample_rate` != the model rate (`--file_sampling_rate`, codec_inference.py:271-274,319-323,353-357; inference / encode
    only): the wrap-padded batch is resampled to the model rate with torchaudio.functional.resample exactly where the reference
    does it, the reconstruction is resampled back and trimmed to the input length, files are written at the file rate -- and, as
    in the reference, the per-utterance frame count is ceil(input length AT THE FILE RATE / hop), which for a higher file rate
    simply keeps every frame of the (wrap-padded) batch row."""
    hop = s2t.model.quantizer.encoder_hop_length
    sr_model = s2t.model.quantizer.sampling_rate
    resample = file_sample_rate is not None and int(file_sample_rate) != int(sr_model)
    if resample and run_mod in ("decode", "decode_emb"):
        raise NotImplementedError("file_sampling_rate != sampling_rate is only defined for the inference / encode modes")
    sr_in = int(file_sample_rate) if resample else sr_model
    writer, sq_writer = None, None
    if output_path is not None:
        os.makedirs(output_path, exist_ok=True)
        writer = IndicesWriter(output_path, need_indices, indices_save_type)
        if need_sub_quants:
            from .kaldi_io import ArkScpWriter
            sq_writer = ArkScpWriter(os.path.join(output_path, "codec_emb"))
    results: List[Dict] = []
    try:
        for group in batches(list(items), batch_size):
            arrs = []
            for key, a in group:
                if isinstance(a, tuple):                        # (samples, rate) from load_wav
                    a, sr = a
                    if sr != sr_in:
                        raise ValueError(f"{key}: sample rate {sr} != {sr_in} (pass its rate as file_sampling_rate)")
                arrs.append(np.asarray(a))
            lens = [int(a.shape[0]) for a in arrs]
            tmax = max(lens)
            pad = [np.pad(a, ((0, tmax - a.shape[0]),) + ((0, 0),) * (a.ndim - 1), mode="wrap") for a in arrs]
            batch = np.stack(pad, axis=0)
            if run_mod == "decode":
                speech = torch.from_numpy(batch.astype(np.int64))
            else:
                speech = torch.from_numpy(batch.astype(np.float32))
            if resample:
                import torchaudio
                speech = torchaudio.functional.resample(speech, orig_freq=sr_in, new_freq=sr_model)
            codes, _, recon, sub = s2t(speech, need_recon=True, bit_width=bit_width, use_scale=use_scale, run_mod=run_mod)
            if resample and recon is not None:
                recon = torchaudio.functional.resample(recon.cpu(), orig_freq=sr_model, new_freq=sr_in)
            for i, (key, _) in enumerate(group):
                if run_mod in ("decode", "decode_emb"):
                    codec_len = lens[i]
                    ilen = codec_len * hop
                else:
                    ilen = lens[i]
                    codec_len = -(-ilen // hop)
                recon_wav = recon[i].cpu()[:, :ilen] if recon is not None else None
                if output_path is None:
                    results.append({"key": key, "value": recon_wav})
                    continue
                if recon_wav is not None:
                    save_wav_pcm16(_wav_name(output_path, key), recon_wav, sr_in if resample else (sample_rate or sr_model), rescale=True)
                if codes is not None:
                    writer.write(key, codes, i, codec_len)
                if sq_writer is not None and sub is not None and sub[0] is not None:
                    sq_writer(key, sub_quants_matrix(sub, i, codec_len))
    finally:
        if writer is not None:
            writer.close()
        if sq_writer is not None:
            sq_writer.close()
    return results
