// FreqCodec 2-D path (BASELINE config 4), first correct CUDA version: channels-last [B][F][T][C] SConv2d /
// SConvTranspose2d on the fp32 SIMT pipe, plus the STFT / iSTFT front and back ends with the mag_phase transforms.
//
// Reference: SConv2d / SConvTranspose2d / pad2d / unpad2d funcodec/modules/normed_modules/conv.py:102-141,317-447;
// FreqCodec._encode_frame / _decode_frame (mag_phase) funcodec/models/codec_freq.py:330-342,365-373,406-425,446-448;
// torchaudio Spectrogram / InverseSpectrogram defaults (n_fft 512, hop 160, periodic hann, center, reflect).
//
// The 2-D conv reuses the 1-D design (conv_simt.cu): every (clip, output frequency row) is a pseudo-clip of a 1-D conv
// along time whose input channels are the K_F frequency taps x C_in channels, gathered from K_F input rows while the
// tile is staged (reflect / zero indexing on both axes, deferred GroupNorm + resblock add + ELU on load).  A transposed
// conv (k = 2s per axis) is the 2x2-tap zero-padded conv with C_out' = s_f*s_t*C_out whose epilogue scatters phase
// (p_f, p_t) to row f*s_f + p_f, column t*s_t + p_t.  GroupNorm partials are emitted per pseudo-clip and summed per clip.
// These kernels are correctness-first (HBM- and FMA-bound rooflines as in conv_simt.cu); the tensor-core version is next.
#include "common.cuh"
#include "kernels.h"

namespace fcb {

template <int TX, int TM>
__global__ void __launch_bounds__(256, 2) conv2d_cl_kernel(const Conv2dParams p) {
    constexpr int TN = 8;
    constexpr int TY = 256 / TX;
    constexpr int CO_TILE = TX * TN;
    constexpr int T_TILE = TY * TM;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int b = blockIdx.z / p.F_out, f_out = blockIdx.z - b * p.F_out;
    const int t0 = blockIdx.x * T_TILE;
    const int co0 = blockIdx.y * CO_TILE;
    const int C_in = p.C_in, cic = p.cic, KT = p.KT, ST = p.ST;
    const int R = (T_TILE - 1) * ST + (KT - 1) + 1;
    const int pitch = cic + 1;
    const bool has1 = p.in1.x != nullptr;

    float* Ws = smem;                                 // [KT][cic][CO_TILE]
    float* Xs = Ws + KT * cic * CO_TILE;              // [R][pitch]
    const float* cf0 = p.in0.coef ? p.in0.coef + (long long)b * 2 * C_in : nullptr;
    const float* cf1 = (has1 && p.in1.coef) ? p.in1.coef + (long long)b * 2 * C_in : nullptr;
    const int gt_max = (p.T_out - 1) * ST - p.pad_t + (KT - 1);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int c_st = tid % cic, r_st = tid / cic, r_step = 256 / cic;
    const int CK = p.KF * C_in;                      // gathered input channels

    for (int ci0 = 0; ci0 < CK; ci0 += cic) {
        __syncthreads();
        const int kfi = ci0 / C_in, cbase = ci0 - kfi * C_in;
        // frequency row of this tap
        int f_src = f_out * p.SF + kfi - p.pad_f;
        bool f_ok = true;
        if (p.pad_zero) f_ok = f_src >= 0 && f_src < p.F_in;
        else f_src = reflect_index(f_src, p.F_in);
        {
            const int c = cbase + c_st;
            float a0 = 1.f, b0 = 0.f, a1 = 1.f, b1 = 0.f;
            if (cf0) { a0 = __ldg(cf0 + c); b0 = __ldg(cf0 + C_in + c); }
            if (cf1) { a1 = __ldg(cf1 + c); b1 = __ldg(cf1 + C_in + c); }
            const float* x0 = p.in0.x + (((long long)b * p.in0.F_raw + p.in0.f_off + f_src) * p.in0.T_raw + p.in0.t_off) * C_in + c;
            const float* x1 = has1 ? p.in1.x + (((long long)b * p.in1.F_raw + p.in1.f_off + f_src) * p.in1.T_raw + p.in1.t_off) * C_in + c : nullptr;
            for (int row = r_st; row < R; row += r_step) {
                const int gt = t0 * ST - p.pad_t + row;
                float v = 0.f;
                bool ok = f_ok && gt <= gt_max;
                int src = gt;
                if (p.pad_zero) ok = ok && gt >= 0 && gt < p.T_in;
                else { src = reflect_index(gt, p.T_in); ok = ok && src >= 0 && src < p.T_in; }
                if (ok) {
                    v = fmaf(__ldg(x0 + (long long)src * C_in), a0, b0);
                    if (has1) v = v + fmaf(__ldg(x1 + (long long)src * C_in), a1, b1);
                    if (p.elu) v = elu1(v);
                }
                Xs[row * pitch + c_st] = v;
            }
        }
        for (int e = tid; e < KT * cic * CO_TILE; e += 256) {
            const int j = e % CO_TILE;
            const int kc = e / CO_TILE;
            const int k = kc / cic, c = kc - k * cic;
            const int co = co0 + j;
            Ws[e] = co < p.C_out_eff ? __ldg(p.w + ((long long)k * CK + ci0 + c) * p.C_out_eff + co) : 0.f;
        }
        __syncthreads();
        for (int c = 0; c < cic; ++c) {
            for (int k = 0; k < KT; ++k) {
                const float* xr = Xs + (ty * ST + k) * pitch + c;
                float a[TM];
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = xr[i * TY * ST * pitch];
                const float* wrow = Ws + (k * cic + c) * CO_TILE + tx * 4;
                const float4 w0 = *reinterpret_cast<const float4*>(wrow);
                const float4 w1 = *reinterpret_cast<const float4*>(wrow + CO_TILE / 2);
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    acc[i][0] = fmaf(a[i], w0.x, acc[i][0]); acc[i][1] = fmaf(a[i], w0.y, acc[i][1]);
                    acc[i][2] = fmaf(a[i], w0.z, acc[i][2]); acc[i][3] = fmaf(a[i], w0.w, acc[i][3]);
                    acc[i][4] = fmaf(a[i], w1.x, acc[i][4]); acc[i][5] = fmaf(a[i], w1.y, acc[i][5]);
                    acc[i][6] = fmaf(a[i], w1.z, acc[i][6]); acc[i][7] = fmaf(a[i], w1.w, acc[i][7]);
                }
            }
        }
    }

    // ---- epilogue: bias, (phase-scattered) raw store, GroupNorm partial statistics
    float s = 0.f, ss = 0.f;
    const int coA = co0 + tx * 4, coB = co0 + CO_TILE / 2 + tx * 4;
    const int F2 = p.F_out * p.FR, T2 = p.T_out * p.TR;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int t = t0 + ty + i * TY;
        if (t >= p.T_out) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int co = half ? coB : coA;
            if (co >= p.C_out_eff) continue;
            // 4 consecutive output channels never straddle a phase boundary (Cc % 4 == 0)
            const int ph = co / p.Cc, cch = co - ph * p.Cc;
            const int pf = ph / p.TR, pt = ph - pf * p.TR;
            float* dst = p.out + (((long long)b * F2 + (long long)f_out * p.FR + pf) * T2 + (long long)t * p.TR + pt) * p.Cc + cch;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (co + j < p.C_out_eff) {
                    const float o = acc[i][4 * half + j] + __ldg(p.bias + co + j);
                    dst[j] = o;
                    s += o; ss = fmaf(o, o, ss);
                }
            }
        }
    }
    if (p.partials) {
        __shared__ double red[64];
        double ds = (double)s, dss = (double)ss;
        block_reduce_2d(ds, dss, red);
        if (tid == 0) {
            const int nparts = gridDim.x * gridDim.y;
            double* dst = p.partials + ((long long)blockIdx.z * nparts + blockIdx.y * gridDim.x + blockIdx.x) * 2;
            dst[0] = ds; dst[1] = dss;
        }
    }
}

static void conv2d_pick(const Conv2dParams& p, int* tx, int* tm) {
    *tx = p.C_out_eff >= 128 ? 16 : (p.C_out_eff >= 64 ? 8 : (p.C_out_eff >= 32 ? 4 : 2));
    *tm = 8;
    const int CO_TILE = *tx * 8, T_TILE = (256 / *tx) * 8;
    const long long ctas = (long long)((p.T_out + T_TILE - 1) / T_TILE) * ((p.C_out_eff + CO_TILE - 1) / CO_TILE) * p.B * p.F_out;
    if (ctas < 2 * 148) *tm = 4;
}

int conv2d_num_parts(const Conv2dParams& p) {
    int tx, tm;
    conv2d_pick(p, &tx, &tm);
    const int CO_TILE = tx * 8, T_TILE = (256 / tx) * tm;
    return ((p.T_out + T_TILE - 1) / T_TILE) * ((p.C_out_eff + CO_TILE - 1) / CO_TILE);
}

template <int TX, int TM>
static cudaError_t launch2d_cfg(Conv2dParams p, cudaStream_t st) {
    constexpr int CO_TILE = TX * 8, T_TILE = (256 / TX) * TM;
    int cic = 32;
    while (cic > 1 && (p.C_in % cic != 0)) cic >>= 1;
    auto bytes = [&](int c) {
        const int R = (T_TILE - 1) * p.ST + (p.KT - 1) + 1;
        return ((size_t)p.KT * c * CO_TILE + (size_t)R * (c + 1)) * sizeof(float);
    };
    while (cic > 1 && bytes(cic) > 100 * 1024) cic >>= 1;
    p.cic = cic;
    const size_t smem = bytes(cic);
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    auto kern = conv2d_cl_kernel<TX, TM>;
    cudaError_t e = ensure_dynamic_smem((const void*)kern, 200 * 1024);
    if (e != cudaSuccess) return e;
    dim3 grid((p.T_out + T_TILE - 1) / T_TILE, (p.C_out_eff + CO_TILE - 1) / CO_TILE, p.B * p.F_out);
    kern<<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_conv2d(const Conv2dParams& p, cudaStream_t st) {
    if (p.Cc % 4 != 0 && p.FR * p.TR > 1) return cudaErrorInvalidValue;
    int tx, tm;
    conv2d_pick(p, &tx, &tm);
#define FCB_CASE2(TX_, TM_) if (tx == TX_ && tm == TM_) return launch2d_cfg<TX_, TM_>(p, st);
    FCB_CASE2(2, 8) FCB_CASE2(2, 4) FCB_CASE2(4, 8) FCB_CASE2(4, 4) FCB_CASE2(8, 8) FCB_CASE2(8, 4) FCB_CASE2(16, 8) FCB_CASE2(16, 4)
#undef FCB_CASE2
    return cudaErrorInvalidConfiguration;
}

// =============================================================================================== STFT front end
// One CTA = 8 frames of one clip.  X[k] = sum_n (x[n]/scale) w[n] e^{-2 pi i k n / N}, then the mag_phase features
// (codec_freq.py:365-373) written channels-last as [B][N/2+1][T_s][cpad] = (log max(|X|,1e-6), Re X/max(|X|,1e-6), Im ...)
// followed by cpad - 3 zero channels (cpad = 4: one element is one aligned 16-byte load for the first conv).
constexpr int STFT_FR = 8;

__global__ void __launch_bounds__(256) stft_magphase_kernel(const float* __restrict__ wav, const float* __restrict__ scale, int L,
                                                            int n_fft, int hop, int n_frames, int cpad, float* __restrict__ feats) {
    extern __shared__ __align__(16) float smem[];
    float* cs = smem;                 // [n_fft] cos(2 pi j / N)
    float* sn = cs + n_fft;           // [n_fft] sin(2 pi j / N)
    float* win = sn + n_fft;          // [n_fft] periodic hann
    float* xs = win + n_fft;          // [n_fft + (STFT_FR-1)*hop] samples of this frame group
    const int b = blockIdx.y, fr0 = blockIdx.x * STFT_FR;
    const int tid = threadIdx.x;
    for (int j = tid; j < n_fft; j += 256) {
        float s, c;
        sincospif(2.0f * (float)j / (float)n_fft, &s, &c);
        cs[j] = c; sn[j] = s;
        win[j] = 0.5f - 0.5f * c;
    }
    const int span = n_fft + (STFT_FR - 1) * hop;
    const float sc = scale ? scale[b] : 1.0f;
    for (int i = tid; i < span; i += 256) {
        const int g = fr0 * hop - n_fft / 2 + i;          // center=True: frame m covers [m*hop - N/2, m*hop + N/2)
        const int src = reflect_index(g, L);
        float v = 0.f;
        if (src >= 0 && src < L) v = wav[(long long)b * L + src] / sc;
        xs[i] = v;
    }
    __syncthreads();
    const int n_bins = n_fft / 2 + 1;
    for (int o = tid; o < n_bins * STFT_FR; o += 256) {
        const int fi = o / n_bins, k = o - fi * n_bins;
        const int m = fr0 + fi;
        if (m >= n_frames) continue;
        const float* xf = xs + fi * hop;
        float re = 0.f, im = 0.f;
        int idx = 0;                                       // (k * n) mod n_fft, incrementally
        for (int n = 0; n < n_fft; ++n) {
            const float xw = xf[n] * win[n];
            re = fmaf(xw, cs[idx], re);
            im = fmaf(-xw, sn[idx], im);
            idx += k;
            if (idx >= n_fft) idx -= n_fft;
        }
        const float mag = hypotf(re, im);
        const float cl = fmaxf(mag, 1e-6f);
        float* dst = feats + (((long long)b * n_bins + k) * n_frames + m) * cpad;
        dst[0] = logf(cl);
        dst[1] = re / cl;
        dst[2] = im / cl;
        for (int c = 3; c < cpad; ++c) dst[c] = 0.f;
    }
}

cudaError_t launch_stft_magphase(const float* wav, const float* scale, int B, int L, int n_fft, int hop, int n_frames,
                                 int cpad, float* feats, cudaStream_t st) {
    if (cpad < 3) return cudaErrorInvalidValue;
    const size_t smem = ((size_t)4 * n_fft + (STFT_FR - 1) * hop) * sizeof(float);
    cudaError_t e = ensure_dynamic_smem((const void*)stft_magphase_kernel, 100 * 1024);
    if (e != cudaSuccess) return e;
    stft_magphase_kernel<<<dim3((n_frames + STFT_FR - 1) / STFT_FR, B), 256, smem, st>>>(wav, scale, L, n_fft, hop, n_frames, cpad, feats);
    return cudaGetLastError();
}

// =============================================================================================== iSTFT back end
// Frame synthesis: (deferred GroupNorm of the decoder's last conv) -> softplus(mag) * (re + i im) (codec_freq.py:417-425)
// -> irfft (DC / Nyquist imaginary parts ignored) -> x hann window, one CTA per (frame, clip) -> frames [B][T_s][N].
__global__ void __launch_bounds__(256) istft_frames_kernel(const float* __restrict__ raw, const float* __restrict__ coef, int F_raw,
                                                           int T_raw, int n_fft, int n_frames, float* __restrict__ frames) {
    extern __shared__ __align__(16) float smem[];
    float* cs = smem;
    float* sn = cs + n_fft;
    float* xr = sn + n_fft;           // [n_bins] Re X
    float* xi = xr + n_fft / 2 + 1;   // [n_bins] Im X
    const int m = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int n_bins = n_fft / 2 + 1;
    for (int j = tid; j < n_fft; j += 256) {
        float s, c;
        sincospif(2.0f * (float)j / (float)n_fft, &s, &c);
        cs[j] = c; sn[j] = s;
    }
    const float* cf = coef + (long long)b * 6;           // [2][3]: a0 a1 a2 b0 b1 b2
    for (int k = tid; k < n_bins; k += 256) {
        const float* src = raw + (((long long)b * F_raw + k) * T_raw + m) * 3;
        const float y0 = fmaf(src[0], cf[0], cf[3]);
        const float y1 = fmaf(src[1], cf[1], cf[4]);
        const float y2 = fmaf(src[2], cf[2], cf[5]);
        const float mag = y0 > 20.f ? y0 : log1pf(expf(y0));      // F.softplus(beta=1, threshold=20)
        xr[k] = mag * y1;
        xi[k] = mag * y2;
    }
    __syncthreads();
    const float inv_n = 1.0f / (float)n_fft;
    for (int j = tid; j < n_fft; j += 256) {
        float acc = 0.f;
        int idx = j;                                               // (k * j) mod n_fft for k = 1
        for (int k = 1; k < n_bins - 1; ++k) {
            acc = fmaf(xr[k], cs[idx], acc);
            acc = fmaf(-xi[k], sn[idx], acc);
            idx += j;
            if (idx >= n_fft) idx -= n_fft;
        }
        const float nyq = (j & 1) ? -xr[n_bins - 1] : xr[n_bins - 1];
        const float v = (xr[0] + nyq + 2.0f * acc) * inv_n;
        const float w = 0.5f - 0.5f * cs[j];
        frames[((long long)b * n_frames + m) * n_fft + j] = v * w;
    }
}

// Overlap-add, window-envelope normalisation, center trim, optional * scale, keep out_len samples (torch.istft + the
// reference's `[:, :, :L]`).
__global__ void istft_ola_kernel(const float* __restrict__ frames, const float* __restrict__ scale, int n_fft, int hop, int n_frames,
                                 int out_len, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= out_len) return;
    const int np = n + n_fft / 2;
    int m_hi = np / hop;
    if (m_hi > n_frames - 1) m_hi = n_frames - 1;
    float acc = 0.f, env = 0.f;
    for (int m = m_hi; m >= 0; --m) {
        const int j = np - m * hop;
        if (j >= n_fft) break;
        float s, c;
        sincospif(2.0f * (float)j / (float)n_fft, &s, &c);
        const float w = 0.5f - 0.5f * c;
        acc += frames[((long long)b * n_frames + m) * n_fft + j];
        env = fmaf(w, w, env);
    }
    float v = acc / env;
    if (scale) v *= scale[b];
    out[(long long)b * out_len + n] = v;
}

cudaError_t launch_istft(const float* raw, const float* coef, int B, int F_raw, int T_raw, int n_fft, int hop, int n_frames,
                         const float* scale, float* frames, float* out, int out_len, cudaStream_t st) {
    const size_t smem = ((size_t)2 * n_fft + 2 * (n_fft / 2 + 1)) * sizeof(float);
    istft_frames_kernel<<<dim3(n_frames, B), 256, smem, st>>>(raw, coef, F_raw, T_raw, n_fft, n_frames, frames);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    istft_ola_kernel<<<dim3((out_len + 255) / 256, B), 256, 0, st>>>(frames, scale, n_fft, hop, n_frames, out_len, out);
    return cudaGetLastError();
}

}  // namespace fcb
