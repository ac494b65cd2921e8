"""CPU ORACLE for the FunCodec encode -> RVQ -> decode hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (CPU, torch tensor ops == the very ATen calls the reference
itself makes; the reference is 100 % Python over torch) of the reference algorithm.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
it.  The product (funcodec_b200/) never does; it fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests / golden vectors for this path, so the
oracle is pinned against outputs of the UNMODIFIED reference modules imported from /root/reference
in the build container: tools/gen_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py.

Every function cites the reference file:line it follows (paths relative to /root/reference/).
All tensors use the reference's own layouts ([B, C, T] activations, [B, T', D] embeddings,
[n_q, B, T'] int64 codes) so the parity tests read like calls on the reference modules.
The state_dict uses the reference's parameter names.
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

EPS_GN = 1e-5  # nn.GroupNorm default eps (funcodec/modules/normed_modules/conv.py:52)


# --------------------------------------------------------------------------- padding arithmetic
def extra_padding_for_conv1d(length: int, kernel_size: int, stride: int, padding_total: int) -> int:
    """funcodec/modules/normed_modules/conv.py:57-64 (get_extra_padding_for_conv1d)."""
    n_frames = (length - kernel_size + padding_total) / stride + 1
    ideal_length = (math.ceil(n_frames) - 1) * stride + (kernel_size - padding_total)
    return ideal_length - length


def pad1d_reflect(x: torch.Tensor, paddings: Tuple[int, int]) -> torch.Tensor:
    """conv.py:82-99 (pad1d, mode='reflect'): short inputs are zero-extended before reflecting."""
    length = x.shape[-1]
    pl, pr = paddings
    assert pl >= 0 and pr >= 0
    max_pad = max(pl, pr)
    extra = 0
    if length <= max_pad:
        extra = max_pad - length + 1
        x = F.pad(x, (0, extra))
    padded = F.pad(x, (pl, pr), mode="reflect")
    end = padded.shape[-1] - extra
    return padded[..., :end]


def conv_paddings(length: int, k: int, s: int, d: int, causal: bool = False) -> Tuple[int, int]:
    """SConv1d padding (conv.py:245-258): returns (pad_left, pad_right incl. extra); causal pads (padding_total, extra)."""
    padding_total = (k - 1) * d - (s - 1)
    extra = extra_padding_for_conv1d(length, k, s, padding_total)
    if causal:
        return padding_total, extra
    pr = padding_total // 2
    pl = padding_total - pr
    return pl, pr + extra


def normed_weight(p: Dict[str, torch.Tensor], base: str) -> torch.Tensor:
    """apply_parametrization_norm (conv.py:25-35): `norm: weight_norm` wraps the conv in torch.nn.utils.weight_norm, whose
    forward pre-hook computes weight = _weight_norm(weight_v, weight_g, dim=0); the other norms keep the plain `weight`."""
    if base + ".weight_g" in p:
        return torch._weight_norm(p[base + ".weight_v"], p[base + ".weight_g"], 0)
    return p[base + ".weight"]


def maybe_group_norm(y, p: Dict[str, torch.Tensor], base: str):
    """get_norm_module (conv.py:37-55): GroupNorm(1, C) for time_group_norm, nn.Identity for weight_norm / none."""
    if base + ".weight" in p:
        return F.group_norm(y, 1, p[base + ".weight"], p[base + ".bias"], EPS_GN)
    return y


# --------------------------------------------------------------------------- L1 modules
def sconv1d(x, p: Dict[str, torch.Tensor], prefix: str, stride: int = 1, dilation: int = 1, causal: bool = False):
    """SConv1d.forward (conv.py:243-261) + NormConv1d.forward (conv.py:155-164):
    reflect pad -> Conv1d(bias) -> GroupNorm(1, C_out) (time_group_norm) / nothing (weight_norm, none)."""
    w = normed_weight(p, prefix + ".conv.conv")
    b = p[prefix + ".conv.conv.bias"]
    k = w.shape[-1]
    pl, pr = conv_paddings(x.shape[-1], k, stride, dilation, causal)
    x = pad1d_reflect(x, (pl, pr))
    y = F.conv1d(x, w, b, stride=stride, dilation=dilation)
    return maybe_group_norm(y, p, prefix + ".conv.norm")


def sconvtr1d(x, p: Dict[str, torch.Tensor], prefix: str, stride: int, causal: bool = False):
    """SConvTranspose1d.forward (conv.py:281-305) + NormConvTranspose1d (conv.py:198-202):
    ConvTranspose1d -> GroupNorm(1, C_out) over the UNtrimmed output (time_group_norm only) -> trim (pl, pr); causal with
    trim_right_ratio = 1 trims everything on the right (conv.py:293-297)."""
    w = normed_weight(p, prefix + ".convtr.convtr")  # [Cin, Cout, k]
    b = p[prefix + ".convtr.convtr.bias"]
    k = w.shape[-1]
    y = F.conv_transpose1d(x, w, b, stride=stride)
    y = maybe_group_norm(y, p, prefix + ".convtr.norm")
    padding_total = k - stride
    pr = padding_total if causal else padding_total // 2
    pl = padding_total - pr
    return y[..., pl: y.shape[-1] - pr]


def elu(x):
    """get_activation('ELU', alpha=1.0) (funcodec/modules/activations.py:24-30)."""
    return F.elu(x, alpha=1.0)


def resblock(x, p, prefix: str, res_kernel: int = 3, dilation: int = 1, causal: bool = False):
    """SEANetResnetBlock.forward (seanet_encoder.py:16-61): shortcut(x) + block(x), true_skip=False; the first block conv
    carries the dilation (dilations=[dilation_base ** j, 1], seanet_encoder.py:125)."""
    h = sconv1d(elu(x), p, prefix + ".block.1", dilation=dilation, causal=causal)
    h = sconv1d(elu(h), p, prefix + ".block.3", causal=causal)
    return sconv1d(x, p, prefix + ".shortcut", causal=causal) + h


def lstm_manual(x_tbc, p, prefix: str, num_layers: int):
    """nn.LSTM semantics restated as an explicit loop (gate order i, f, g, o; zero initial state;
    two biases) -- used for the float64 'truth' runs and to cross-check lstm_aten."""
    T, B, H = x_tbc.shape
    inp = x_tbc
    for l in range(num_layers):
        w_ih = p[f"{prefix}.lstm.weight_ih_l{l}"]
        w_hh = p[f"{prefix}.lstm.weight_hh_l{l}"]
        b_ih = p[f"{prefix}.lstm.bias_ih_l{l}"]
        b_hh = p[f"{prefix}.lstm.bias_hh_l{l}"]
        h = x_tbc.new_zeros(B, H)
        c = x_tbc.new_zeros(B, H)
        gx = inp @ w_ih.t() + b_ih  # [T, B, 4H]
        outs = []
        for t in range(T):
            g = gx[t] + h @ w_hh.t() + b_hh
            i, f, gg, o = g.chunk(4, dim=-1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs)
    return inp


def lstm_aten(x_tbc, p, prefix: str, num_layers: int):
    """The reference's own call: nn.LSTM(dimension, dimension, num_layers) (lstm.py:20,24)."""
    H = x_tbc.shape[-1]
    m = torch.nn.LSTM(H, H, num_layers).to(device=x_tbc.device, dtype=x_tbc.dtype)
    with torch.no_grad():
        for name, par in m.named_parameters():
            par.copy_(p[f"{prefix}.lstm.{name}"])
        y, _ = m(x_tbc)
    return y


def slstm(x, p, prefix: str, num_layers: int = 2, manual: bool = False):
    """SLSTM.forward (funcodec/modules/normed_modules/lstm.py:22-28): [B,C,T]->[T,B,C], LSTM, +skip."""
    xt = x.permute(2, 0, 1)
    y = (lstm_manual if manual else lstm_aten)(xt, p, prefix, num_layers)
    y = y + xt
    return y.permute(1, 2, 0)


# --------------------------------------------------------------------------- L2 sub-models
def sub_dict(p: Dict[str, torch.Tensor], head: str) -> Dict[str, torch.Tensor]:
    n = len(head)
    return {k[n:]: v for k, v in p.items() if k.startswith(head)}


def seanet_encoder(x, p, ratios: Sequence[int], lstm_layers: int = 2, manual_lstm: bool = False,
                   n_residual_layers: int = 1, dilation_base: int = 2, causal: bool = False):
    """SEANetEncoder.forward (seanet_encoder.py:108-162,171-185).  x: [B,1,L] -> [B,T',D].
    `ratios` as given in the YAML; the encoder applies them reversed (seanet_encoder.py:102).  n_residual_layers > 1 (the
    soundstream_* YAMLs) stacks residual blocks with dilations dilation_base ** j (:122-128); lstm_layers = 0 is
    `seq_model: none`."""
    h = sconv1d(x, p, "model.0", causal=causal)
    n = 1
    for r in reversed(list(ratios)):
        for j in range(n_residual_layers):
            h = resblock(h, p, f"model.{n + j}", dilation=dilation_base ** j, causal=causal)
        h = sconv1d(elu(h), p, f"model.{n + n_residual_layers + 1}", stride=r, causal=causal)
        n += n_residual_layers + 2
    if lstm_layers > 0:
        h = slstm(h, p, f"model.{n}", lstm_layers, manual_lstm)
        n += 1
    h = sconv1d(elu(h), p, f"model.{n + 1}", causal=causal)
    return h.permute(0, 2, 1)


def seanet_decoder(z, p, ratios: Sequence[int], lstm_layers: int = 2, manual_lstm: bool = False,
                   n_residual_layers: int = 1, dilation_base: int = 2, causal: bool = False):
    """SEANetDecoder.forward (seanet_decoder.py:107-172,177-180).  z: [B,T',D] -> [B,1,T'*hop]."""
    h = sconv1d(z.permute(0, 2, 1), p, "model.0", causal=causal)
    n = 1
    if lstm_layers > 0:
        h = slstm(h, p, "model.1", lstm_layers, manual_lstm)
        n = 2
    for r in ratios:
        h = sconvtr1d(elu(h), p, f"model.{n + 1}", stride=r, causal=causal)
        for j in range(n_residual_layers):
            h = resblock(h, p, f"model.{n + 2 + j}", dilation=dilation_base ** j, causal=causal)
        n += n_residual_layers + 2
    return sconv1d(elu(h), p, f"model.{n + 1}", causal=causal)


# --------------------------------------------------------------------------- RVQ
def num_quantizers_for_bandwidth(n_q_max: int, bins: int, sample_rate: int, hop: int,
                                 bandwidth: Optional[float]) -> int:
    """ResidualVectorQuantizer.get_num_quantizers_for_bandwidth (vq.py:105-117)."""
    bw_per_q = math.log2(bins) * sample_rate / hop
    n_q = n_q_max
    if bandwidth and bandwidth > 0.0:
        n_q = int(max(1, math.floor(bandwidth / bw_per_q)))
    return n_q


def codebook_quantize(x2d, embed, want_margin: bool = False):
    """EuclideanCodebook.quantize (ddp_core_vq.py:180-188): argmax of -(|x|^2 - 2 x.C^T + |c|^2),
    first maximal index wins.  x2d: [M, D]; embed: [K, D]."""
    e = embed.t()
    dist = -(x2d.pow(2).sum(1, keepdim=True) - 2 * x2d @ e + e.pow(2).sum(0, keepdim=True))
    ind = dist.max(dim=-1).indices
    if not want_margin:
        return ind, None
    top2 = dist.topk(2, dim=-1).values
    return ind, (top2[:, 0] - top2[:, 1])


def rvq_forward(x_bdt, embed, n_q: int, want_margin: bool = False):
    """DistributedResidualVectorQuantization.forward, eval path (ddp_core_vq.py:367-418) with
    VectorQuantization.forward (:305-324) and EuclideanCodebook.forward (:212-241) inlined.
    x_bdt: [B, D, T'] -> (quantized [B,D,T'], codes [n_q,B,T'] i64, sub_quants [n_q,B,D,T'],
    margins [n_q,B,T'] | None)."""
    B, D, T = x_bdt.shape
    quantized_out = torch.zeros_like(x_bdt)
    residual = x_bdt
    all_idx, all_sub, all_margin = [], [], []
    for q in range(n_q):
        xin = residual.permute(0, 2, 1)  # rearrange b d n -> b n d (:306)
        flat = xin.reshape(-1, D)
        ind, margin = codebook_quantize(flat, embed[q], want_margin)
        quant = F.embedding(ind.view(B, T), embed[q]).permute(0, 2, 1)  # dequantize (:190-192)
        residual = residual - quant
        quantized_out = quantized_out + quant
        all_idx.append(ind.view(B, T))
        all_sub.append(quant)
        if want_margin:
            all_margin.append(margin.view(B, T))
    return (quantized_out, torch.stack(all_idx), torch.stack(all_sub),
            torch.stack(all_margin) if want_margin else None)


def rvq_decode(codes_qbt, embed):
    """DistributedResidualVectorQuantization.decode (ddp_core_vq.py:442-453): sum_q embedding.
    codes: [n_q, B, T'] -> [B, D, T'] (after VectorQuantization.decode's b n d -> b d n)."""
    out = torch.tensor(0.0, dtype=embed.dtype)
    for q, ind in enumerate(codes_qbt):
        out = out + F.embedding(ind, embed[q]).permute(0, 2, 1)
    return out


# --------------------------------------------------------------------------- L3 model (Encodec)
def linear_overlap_add(frames, stride: int):
    """_linear_overlap_add (codec_basic.py:77-116): triangular weights taken from the FIRST frame's length (a shorter
    later frame uses the leading part of the same triangle), weighted sum / sum of weights."""
    total_size = stride * (len(frames) - 1) + frames[-1].shape[-1]
    frame_length = frames[0].shape[-1]
    t = torch.linspace(0, 1, frame_length + 2, dtype=frames[0].dtype)[1:-1]
    weight = 0.5 - (t - 0.5).abs()
    sum_weight = torch.zeros(total_size, dtype=frames[0].dtype)
    out = torch.zeros(*frames[0].shape[:-1], total_size, dtype=frames[0].dtype)
    offset = 0
    for frame in frames:
        n = frame.shape[-1]
        out[..., offset:offset + n] += weight[:n] * frame
        sum_weight[offset:offset + n] += weight[:n]
        offset += stride
    return out / sum_weight


def segment_plan(length: int, sample_rate: int, segment_dur, overlap_ratio: float):
    """(segment_length, stride, offsets) of Encodec._encode (codec_basic.py:287-298,346-358)."""
    if segment_dur is None:
        return length, length, [0]
    seg = int(segment_dur * sample_rate)
    stride = max(1, int((1 - overlap_ratio) * seg))
    return seg, stride, list(range(0, length, stride))


class OracleEncodec:
    """Restatement of Encodec's inference methods (funcodec/models/codec_basic.py:670-836) for the
    shipped time-domain configs: norm time_group_norm / weight_norm / none (read off the state_dict keys), causal or not,
    stacked dilated residual blocks, LSTM or no sequence model, audio_normalize, RVQ without projections
    (CostumeQuantizer, costume_quantizer.py:77-119); segment_dur=None (one frame) or a segment length in seconds
    (per-segment normalisation / encode / decode + linear overlap-add, codec_basic.py:334-359,382-396)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], ratios: Sequence[int], sample_rate: int = 16000,
                 lstm_layers: int = 2, audio_normalize: bool = True, dtype=torch.float32,
                 manual_lstm: bool = False, segment_dur=None, overlap_ratio: float = 0.01, device="cpu",
                 n_residual_layers: int = 1, dilation_base: int = 2, causal: bool = False):
        self.segment_dur = segment_dur
        self.causal = causal                            # encoder_conf / decoder_conf causal (soundstream_16k YAML); the norm
        #                                                 (time_group_norm / weight_norm / none) follows from the state_dict keys
        self.overlap_ratio = overlap_ratio
        # device != "cpu" is only used by bench.py's labelled CUDA-eager context leg (same ATen ops on the GPU)
        sd = {k: v.detach().to(device, dtype) if v.is_floating_point() else v.detach().to(device)
              for k, v in state_dict.items()}
        self.enc = sub_dict(sd, "encoder.")
        self.dec = sub_dict(sd, "decoder.")
        self.embed = sd["quantizer.rq.model.embed"]
        self.ratios = list(ratios)
        self.hop = int(math.prod(self.ratios))
        self.sample_rate = sample_rate
        self.lstm_layers = lstm_layers
        self.n_residual_layers = n_residual_layers      # stacked dilated residual blocks (soundstream_noncausal YAMLs)
        self.dilation_base = dilation_base
        self.audio_normalize = audio_normalize
        self.dtype = dtype
        self.manual_lstm = manual_lstm
        self.n_q_max, self.bins, self.dim = self.embed.shape

    @classmethod
    def from_config(cls, state_dict, cfg, **kw):
        """Oracle for a hyper-parameter record with the YAML-derived fields (ratios, sample_rate, lstm_layers, audio_normalize,
        n_residual_layers, dilation_base, causal); the norm follows from the state_dict keys."""
        return cls(state_dict, cfg.ratios, cfg.sample_rate, cfg.lstm_layers, audio_normalize=cfg.audio_normalize,
                   n_residual_layers=getattr(cfg, "n_residual_layers", 1), dilation_base=getattr(cfg, "dilation_base", 2),
                   causal=bool(getattr(cfg, "causal", False)), **kw)

    # codec_basic.py:361-380
    def encode_frame(self, x_b1l):
        scale = None
        if self.audio_normalize:
            mono = x_b1l.mean(dim=1, keepdim=True)
            volume = mono.pow(2).mean(dim=2, keepdim=True).sqrt()
            scale = 1e-8 + volume
            x_b1l = x_b1l / scale
            scale = scale.view(-1, 1)
        emb = seanet_encoder(x_b1l, self.enc, self.ratios, self.lstm_layers, self.manual_lstm, self.n_residual_layers, self.dilation_base,
                             self.causal)
        return emb, scale

    # codec_basic.py:398-408
    def decode_frame(self, emb_btd, scale):
        out = seanet_decoder(emb_btd, self.dec, self.ratios, self.lstm_layers, self.manual_lstm, self.n_residual_layers, self.dilation_base,
                             self.causal)
        if scale is not None:
            out = out * scale.view(-1, 1, 1)
        return out

    def n_q_for(self, bit_width):
        return num_quantizers_for_bandwidth(self.n_q_max, self.bins, self.sample_rate, self.hop, bit_width)

    @torch.no_grad()
    def inference(self, speech, need_recon=True, bit_width=None, use_scale=True, want_margin=False):
        """Encodec.inference / inference_encoding (codec_basic.py:670-764)."""
        speech = speech.to(self.dtype)
        if speech.dim() == 2:
            speech = speech.unsqueeze(1)
        if self.segment_dur is not None:
            return self._inference_segmented(speech, need_recon, bit_width, use_scale, want_margin)
        emb, scale = self.encode_frame(speech)
        n_q = self.n_q_for(bit_width)
        quant, codes, sub, margins = rvq_forward(emb.permute(0, 2, 1), self.embed, n_q, want_margin)
        quant_btd = quant.permute(0, 2, 1)
        recon = None
        if need_recon:
            recon = self.decode_frame(quant_btd, scale if use_scale else None)[:, :, :speech.shape[-1]]
        return dict(recon_speech=recon, code_indices=[codes],
                    code_embeddings=[(quant_btd, scale if use_scale else None)],
                    sub_quants=[sub], encoder_out=emb, margins=margins)

    def _inference_segmented(self, speech, need_recon, bit_width, use_scale, want_margin):
        seg, stride, offsets = segment_plan(speech.shape[-1], self.sample_rate, self.segment_dur, self.overlap_ratio)
        n_q = self.n_q_for(bit_width)
        idx, embs, subs, encs, margins, frames = [], [], [], [], [], []
        for off in offsets:
            emb, scale = self.encode_frame(speech[:, :, off:off + seg])
            quant, codes, sub, mg = rvq_forward(emb.permute(0, 2, 1), self.embed, n_q, want_margin)
            quant_btd = quant.permute(0, 2, 1)
            idx.append(codes); subs.append(sub); encs.append(emb); margins.append(mg)
            embs.append((quant_btd, scale if use_scale else None))
            if need_recon:
                frames.append(self.decode_frame(quant_btd, scale if use_scale else None))
        recon = linear_overlap_add(frames, stride)[:, :, :speech.shape[-1]] if need_recon else None
        return dict(recon_speech=recon, code_indices=idx, code_embeddings=embs, sub_quants=subs, encoder_out=encs,
                    margins=margins)

    @torch.no_grad()
    def inference_decoding(self, token_idx_btq):
        """Encodec.inference_decoding (codec_basic.py:766-802): tokens [B,T',n_q] -> wav."""
        codes = token_idx_btq.permute(2, 0, 1)
        emb = rvq_decode(codes, self.embed).transpose(1, 2)
        return dict(recon_speech=self.decode_frame(emb, None), code_indices=None,
                    code_embeddings=[(emb, None)], sub_quants=None)

    @torch.no_grad()
    def inference_decoding_emb(self, emb_btd):
        """Encodec.inference_decoding_emb (codec_basic.py:804-836)."""
        emb_btd = emb_btd.to(self.dtype)
        return dict(recon_speech=self.decode_frame(emb_btd, None), code_indices=None,
                    code_embeddings=[(emb_btd, None)], sub_quants=None)
