"""CLI with the flag names of `python -m funcodec.bin.codec_inference` that `egs/LibriTTS/codec/encoding_decoding.sh`
passes (funcodec/bin/codec_inference.py:428-558), backed by the CUDA library.

    python -m funcodec_b200.bin.codec_inference --run_mod encode --config_file config.yaml --model_file model.pth \
        --data_path_and_name_and_type wav.scp,speech,sound --output_dir out --batch_size 16 --bit_width 8000
    python -m funcodec_b200.bin.codec_inference --run_mod decode --data_path_and_name_and_type out/codecs.txt,speech,codec_json ...

`config.yaml` is the training config the reference saves (encoder_conf / quantizer_conf / decoder_conf / model_conf);
`model.pth` is the plain state_dict.  Only the configurations listed in DESIGN.md §9 are accepted.
"""
import argparse
import math
import sys

import torch
import yaml

from funcodec_b200.config import CodecConfig
from funcodec_b200.encodec import B200Encodec
from funcodec_b200.pipeline import run_decode, run_encode
from funcodec_b200.speech2token import Speech2Token


def config_from_yaml(path: str) -> CodecConfig:
    with open(path, "rt", encoding="utf-8") as f:
        a = yaml.safe_load(f)
    enc, dec, q, m = a.get("encoder_conf", {}), a.get("decoder_conf", {}), a.get("quantizer_conf", {}), a.get("model_conf", {})
    if enc.get("norm") != "time_group_norm" or enc.get("causal", False) or m.get("segment_dur") is not None:
        raise SystemExit("unsupported configuration (needs norm: time_group_norm, causal: false, segment_dur: null)")
    ratios = tuple(dec.get("ratios", [8, 5, 4, 2]))
    if tuple(enc.get("ratios", [8, 5, 4, 2])) != ratios:
        raise SystemExit("encoder and decoder ratios differ")
    if int(q.get("encoder_hop_length", 320)) != math.prod(ratios):
        raise SystemExit("quantizer_conf.encoder_hop_length != prod(ratios)")
    return CodecConfig(name="from_yaml", ratios=ratios, n_filters=int(enc.get("n_filters", 32)),
                       dimension=int(m.get("odim", 128)), codebook_size=int(q.get("codebook_size", 1024)),
                       num_quantizers=int(q.get("num_quantizers", 32)), sample_rate=int(q.get("sampling_rate", 16000)),
                       audio_normalize=bool(m.get("audio_normalize", True)),
                       lstm_layers=int(enc.get("seq_layer_num", 2)))


def main(argv=None):
    p = argparse.ArgumentParser(description="Speech Tokenizer (B200)")
    p.add_argument("--output_dir", required=True)
    p.add_argument("--config_file", required=True)
    p.add_argument("--model_file", required=True)
    p.add_argument("--data_path_and_name_and_type", required=True, help="path,name,type (sound | codec_json)")
    p.add_argument("--run_mod", default="inference", choices=["inference", "encode", "decode"])
    p.add_argument("--batch_size", type=int, default=16)
    p.add_argument("--bit_width", type=int, default=None)
    p.add_argument("--use_scale", type=lambda s: s.lower() in ("1", "true"), default=False)
    p.add_argument("--gpuid_list", default="0")
    args = p.parse_args(argv)
    path = args.data_path_and_name_and_type.split(",")[0]
    device = f"cuda:{args.gpuid_list.split(',')[0] or 0}"
    cfg = config_from_yaml(args.config_file)
    sd = torch.load(args.model_file, map_location="cpu")
    s2t = Speech2Token(B200Encodec(cfg, sd, device), device)
    if args.run_mod == "decode":
        n = run_decode(s2t, path, args.output_dir, args.batch_size, args.bit_width)
    else:
        n = run_encode(s2t, path, args.output_dir, args.batch_size, args.bit_width, args.run_mod, args.use_scale,
                       save_recon=(args.run_mod == "inference"))
    print(f"processed {n} utterances -> {args.output_dir}")


if __name__ == "__main__":
    main(sys.argv[1:])
