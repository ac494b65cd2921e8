// SLSTM recurrent step (fp32 SIMT).  Reference: funcodec/modules/normed_modules/lstm.py:12-28
// (nn.LSTM(dim, dim, num_layers), gate order i,f,g,o, zero initial state, y = lstm(x) + x).
//
// The input projections x_t W_ih^T + b_ih + b_hh for all t are one GEMM (conv_simt.cu as a 1x1 conv),
// written as gx[B][T][4H] with unit-major packed columns (n' = 4*j + gate) so that the CTA owning hidden
// units [j0, j0+8) finds its 32 gate columns contiguous.  One launch per timestep:
//     gates = gx[:, t] + h_{t-1} W_hh^T ;  c = sig(f) c + sig(i) tanh(g) ;  h = sig(o) tanh(c)
// grid = (H/8, ceil(B/16)); each CTA splits K=H over its 8 warps, lanes own the 32 gate columns,
// partial sums are reduced through shared memory and 128 threads do the cell update.
// Bytes per step: W_hh (16.8 MB at H=1024, L2-resident) + B*H*4*3; latency-bound by design (sequential in t).
#include "common.cuh"
#include "kernels.h"

namespace fcb {

constexpr int LSTM_UNITS = 8;     // hidden units per CTA
constexpr int LSTM_BG = 16;       // clips per CTA

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256) lstm_step_kernel(const LstmStepParams p) {
    extern __shared__ __align__(16) float smem[];
    const int H = p.H, T = p.T, t = p.t;
    float* Hs = smem;                              // [LSTM_BG][H]
    float* red = Hs + LSTM_BG * H;                 // [8 warps][LSTM_BG][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j0 = blockIdx.x * LSTM_UNITS;
    const int b0 = blockIdx.y * LSTM_BG;
    const int nb = min(LSTM_BG, p.B - b0);

    // stage h_{t-1} for this CTA's clips
    for (int e = tid; e < LSTM_BG * H; e += 256) {
        const int bb = e / H, k = e - bb * H;
        float v = 0.f;
        if (t > 0 && bb < nb) v = p.h_seq[((long long)(b0 + bb) * T + (t - 1)) * H + k];
        Hs[e] = v;
    }
    __syncthreads();

    float acc[LSTM_BG];
#pragma unroll
    for (int i = 0; i < LSTM_BG; ++i) acc[i] = 0.f;
    if (t > 0) {
        const int kslice = H / 8;
        const int kbeg = warp * kslice;
        const float* wcol = p.whh + (long long)j0 * 4 + lane;      // column of this lane
        for (int k = kbeg; k < kbeg + kslice; k += 4) {
            const float w0 = __ldg(wcol + (long long)(k + 0) * 4 * H);
            const float w1 = __ldg(wcol + (long long)(k + 1) * 4 * H);
            const float w2 = __ldg(wcol + (long long)(k + 2) * 4 * H);
            const float w3 = __ldg(wcol + (long long)(k + 3) * 4 * H);
#pragma unroll
            for (int i = 0; i < LSTM_BG; ++i) {
                const float4 hv = *reinterpret_cast<const float4*>(Hs + i * H + k);
                acc[i] = fmaf(hv.x, w0, acc[i]);
                acc[i] = fmaf(hv.y, w1, acc[i]);
                acc[i] = fmaf(hv.z, w2, acc[i]);
                acc[i] = fmaf(hv.w, w3, acc[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < LSTM_BG; ++i) red[(warp * LSTM_BG + i) * 32 + lane] = acc[i];
    __syncthreads();

    if (tid < LSTM_BG * LSTM_UNITS) {
        const int bb = tid / LSTM_UNITS, u = tid % LSTM_UNITS;
        if (bb < nb) {
            const int b = b0 + bb, j = j0 + u;
            float g4[4];
            const float* gxp = p.gx + ((long long)b * T + t) * 4 * H + (long long)j * 4;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += red[(w * LSTM_BG + bb) * 32 + u * 4 + g];
                g4[g] = gxp[g] + s;
            }
            const float ig = sigmoidf_(g4[0]), fg = sigmoidf_(g4[1]), gg = tanhf(g4[2]), og = sigmoidf_(g4[3]);
            float* cp = p.c_state + (long long)b * H + j;
            const float c = (t > 0 ? fg * (*cp) : 0.f) + ig * gg;
            *cp = c;
            const float h = og * tanhf(c);
            const long long o = ((long long)b * T + t) * H + j;
            p.h_seq[o] = h;
            if (p.y_out) {
                const long long xo = (long long)b * p.skip.clip_stride + ((long long)(p.skip.row_off + t)) * H + j;
                float xv = p.skip.x[xo];
                if (p.skip.stats) {
                    const float mean = p.skip.stats[2 * b], rstd = p.skip.stats[2 * b + 1];
                    const float a = rstd * p.skip.gamma[j];
                    xv = fmaf(xv, a, p.skip.beta[j] - a * mean);
                }
                p.y_out[o] = h + xv;
            }
        }
    }
}

cudaError_t launch_lstm_step(const LstmStepParams& p, cudaStream_t st) {
    if (p.H % 32 != 0) return cudaErrorInvalidValue;   // K split over 8 warps in steps of 4
    const size_t smem = ((size_t)LSTM_BG * p.H + 8 * LSTM_BG * 32) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(lstm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid(p.H / LSTM_UNITS, (p.B + LSTM_BG - 1) / LSTM_BG);
    lstm_step_kernel<<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace fcb
