"""The three ways LauraTTS (the largest in-repo consumer of the codec hot path) calls the codec, replayed
verbatim against any `Speech2Token`-shaped callable:

  * prompt encoding      funcodec/bin/text2audio_inference.py:157-160
        codec = codec_model(prompt_audio, run_mod="encode")[0][0].squeeze(1).transpose(0, 1);  continual = codec[:, :predict_nq]
  * LM-only vocoding     text2audio_inference.py:180-184
        _, _, wav, _ = codec_model(decoded_codec[:, continual_length:], bit_width=None, run_mod="decode")
  * embedding vocoding   funcodec/models/audio_generation/laura_model.py:565
        _, _, wav, _ = codec_model(codec_emb[:, continual_length:], run_mod="decode_emb")

The language model between the calls is out of scope; its outputs are stood in for by deterministic functions of the prompt
codes (the prompt followed by a shifted copy; the sum of the first predict_nq codewords)."""
import torch


def laura_codec_calls(codec_model, prompt_audio, embed, predict_nq=2, exclude_prompt=True):
    codec = codec_model(prompt_audio, run_mod="encode")[0][0].squeeze(1).transpose(0, 1)          # [T', n_q]
    continual = codec[:, :predict_nq].tolist()
    continual_length = len(continual) if exclude_prompt else 0
    prompt = torch.tensor(continual, dtype=torch.int64, device=codec.device)                       # [T', predict_nq]
    generated = torch.roll(prompt, shifts=3, dims=0)                                               # stand-in for the LM
    decoded_codec = torch.cat([prompt, generated], dim=0).unsqueeze(0)                             # [1, 2T', predict_nq]
    _, _, gen_only_lm, _ = codec_model(decoded_codec[:, continual_length:], bit_width=None, run_mod="decode")
    # stand-in for cal_codec_emb: dense embedding = sum of the predicted groups' codewords (QuantizerCodebook.forward)
    emb = torch.zeros(1, decoded_codec.shape[1], embed.shape[-1], device=codec.device)
    for q in range(predict_nq):
        emb = emb + embed[q].to(codec.device)[decoded_codec[0, :, q]].unsqueeze(0)
    _, _, gen, _ = codec_model(emb[:, continual_length:], run_mod="decode_emb")                    # a non-contiguous slice
    return dict(codec=codec, continual=continual, gen_only_lm=gen_only_lm, gen=gen)


class OracleSpeech2Token:
    """The CPU oracle behind the Speech2Token call signature (codec_inference.py:86-134), for the parity tests only."""

    def __init__(self, oracle):
        self.oracle = oracle

    def __call__(self, speech, ppg=None, need_recon=True, bit_width=None, use_scale=True, run_mod="inference"):
        o = self.oracle
        if run_mod == "inference":
            r = o.inference(speech, need_recon=need_recon, bit_width=bit_width, use_scale=use_scale)
        elif run_mod == "encode":
            r = o.inference(speech, need_recon=False, bit_width=bit_width)
        elif run_mod == "decode_emb":
            r = o.inference_decoding_emb(speech)
        else:
            r = o.inference_decoding(speech)
        return r["code_indices"], r["code_embeddings"], r["recon_speech"], r["sub_quants"]
