// SLSTM recurrence as ONE persistent cooperative kernel per layer (fp32 SIMT).
// Reference: funcodec/modules/normed_modules/lstm.py:12-28 (nn.LSTM(dim, dim, num_layers), gate order
// i,f,g,o, zero initial state, y = lstm(x) + x).
//
// The input projections x_t W_ih^T + b_ih + b_hh for all t are one GEMM (conv kernel as a 1x1 conv) written
// as gx[B][T][4H] with unit-major packed columns (n' = 4*j + gate).  The recurrence
//     gates = gx[:, t] + h_{t-1} W_hh^T ;  c = sig(f) c + sig(i) tanh(g) ;  h = sig(o) tanh(c)
// is strictly sequential in t, so the kernel is built around latency:
//   * grid = H / UNITS CTAs (<= 148, one per SM, cooperative launch), CTA j owns hidden units
//     [j*UNITS, (j+1)*UNITS) and keeps its W_hh slice [H][4*UNITS] (128 KB at H=1024) in shared memory for
//     all T steps -- W_hh is read from HBM/L2 exactly once per layer instead of once per step;
//   * per step every CTA needs the whole h_{t-1} [B][H]: it is exchanged through global memory (L2) with
//     a grid-wide release/acquire counter barrier; gx for the step is prefetched before the barrier wait;
//   * a thread accumulates the 4 gates of one unit for 16 clips (64 fp32 accumulators) over an interleaved
//     K slice (W rows via conflict-free LDS.128, h via broadcast LDS.128), K slices are reduced with
//     shuffles + one shared-memory pass, and UNITS*16 threads do the cell update.
// Latency-bound by construction (T' dependent steps); FLOPs = 2*B*T*4H*H per layer.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace fcb {

constexpr int LSTM_BG = 16;       // clips processed together (accumulator tile)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int UNITS>
__global__ void __launch_bounds__(256, 1) lstm_seq_kernel(const LstmSeqParams p) {
    constexpr int COLS = 4 * UNITS;            // gate columns owned by this CTA
    constexpr int KS_PER_WARP = 32 / UNITS;    // K slices inside a warp
    constexpr int NSLICE = 8 * KS_PER_WARP;    // K slices per CTA (interleaved in groups of 4 k)
    extern __shared__ __align__(16) float smem[];
    const int H = p.H, T = p.T, B = p.B;
    float* Ws = smem;                           // [H][COLS]
    float* Hs = Ws + (size_t)H * COLS;          // [LSTM_BG][H]
    float* red = Hs + LSTM_BG * H;              // [8 warps][LSTM_BG][COLS]
    float* cS = red + 8 * COLS * LSTM_BG;       // [nbg][LSTM_BG][UNITS] cell state
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int u = lane % UNITS, ks = lane / UNITS;
    const int slice = warp * KS_PER_WARP + ks;
    const int j0 = blockIdx.x * UNITS;
    const int nbg = (B + LSTM_BG - 1) / LSTM_BG;

    // one-time: W_hh slice -> shared memory (coalesced rows of COLS floats)
    for (int e = tid; e < H * COLS; e += 256) {
        const int k = e / COLS, c = e - k * COLS;
        Ws[e] = __ldg(p.whh + (long long)k * 4 * H + (long long)j0 * 4 + c);
    }
    for (int e = tid; e < nbg * LSTM_BG * UNITS; e += 256) cS[e] = 0.f;

    // finalize-thread identity: (clip bb, unit fu)
    const bool fin = tid < LSTM_BG * UNITS;
    const int fbb = tid / UNITS, fu = tid % UNITS;
    unsigned bar_target = 0;

    for (int t = 0; t < T; ++t) {
        // ---- grid barrier: h_{t-1} of every CTA must be visible (skip at t == 0: h_{-1} = 0)
        if (t > 0) {
            bar_target += gridDim.x;
            if (tid == 0) {
                while (ld_acquire_u32(p.barrier) < bar_target) { }
                __threadfence();
            }
        }
        __syncthreads();
        for (int bg = 0; bg < nbg; ++bg) {
            const int b0 = bg * LSTM_BG;
            const int nb = min(LSTM_BG, B - b0);
            // prefetch this step's input projection for the finalize threads
            float4 gxv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (fin && fbb < nb)
                gxv = __ldcs(reinterpret_cast<const float4*>(p.gx + ((long long)(b0 + fbb) * T + t) * 4 * H + (long long)(j0 + fu) * 4));
            float acc[4][LSTM_BG];
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int i = 0; i < LSTM_BG; ++i) acc[g][i] = 0.f;
            if (t > 0) {
                // stage h_{t-1} (written by other CTAs: bypass L1)
                for (int e = tid * 4; e < LSTM_BG * H; e += 256 * 4) {
                    const int bb = e / H, k = e - bb * H;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bb < nb) v = __ldcg(reinterpret_cast<const float4*>(p.h_seq + ((long long)(b0 + bb) * T + (t - 1)) * H + k));
                    *reinterpret_cast<float4*>(Hs + e) = v;
                }
                __syncthreads();
                for (int k0 = slice * 4; k0 < H; k0 += NSLICE * 4) {
                    float4 w[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(Ws + (k0 + kk) * COLS + u * 4);
#pragma unroll
                    for (int i = 0; i < LSTM_BG; ++i) {
                        const float4 h4 = *reinterpret_cast<const float4*>(Hs + i * H + k0);
                        acc[0][i] = fmaf(h4.x, w[0].x, acc[0][i]); acc[1][i] = fmaf(h4.x, w[0].y, acc[1][i]);
                        acc[2][i] = fmaf(h4.x, w[0].z, acc[2][i]); acc[3][i] = fmaf(h4.x, w[0].w, acc[3][i]);
                        acc[0][i] = fmaf(h4.y, w[1].x, acc[0][i]); acc[1][i] = fmaf(h4.y, w[1].y, acc[1][i]);
                        acc[2][i] = fmaf(h4.y, w[1].z, acc[2][i]); acc[3][i] = fmaf(h4.y, w[1].w, acc[3][i]);
                        acc[0][i] = fmaf(h4.z, w[2].x, acc[0][i]); acc[1][i] = fmaf(h4.z, w[2].y, acc[1][i]);
                        acc[2][i] = fmaf(h4.z, w[2].z, acc[2][i]); acc[3][i] = fmaf(h4.z, w[2].w, acc[3][i]);
                        acc[0][i] = fmaf(h4.w, w[3].x, acc[0][i]); acc[1][i] = fmaf(h4.w, w[3].y, acc[1][i]);
                        acc[2][i] = fmaf(h4.w, w[3].z, acc[2][i]); acc[3][i] = fmaf(h4.w, w[3].w, acc[3][i]);
                    }
                }
                // reduce the K slices that live in the same warp (lanes differing in ks)
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int i = 0; i < LSTM_BG; ++i) {
                        float v = acc[g][i];
#pragma unroll
                        for (int o = UNITS; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        acc[g][i] = v;
                    }
                if (ks == 0) {      // lanes u = 0..UNITS-1 write consecutive float4: conflict-free
#pragma unroll
                    for (int i = 0; i < LSTM_BG; ++i)
                        *reinterpret_cast<float4*>(red + (warp * LSTM_BG + i) * COLS + u * 4) =
                            make_float4(acc[0][i], acc[1][i], acc[2][i], acc[3][i]);
                }
                __syncthreads();
            }
            if (fin && fbb < nb) {
                const int b = b0 + fbb, j = j0 + fu;
                float g4[4] = {gxv.x, gxv.y, gxv.z, gxv.w};
                if (t > 0) {
                    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int w8 = 0; w8 < 8; ++w8) {
                        const float4 r4 = *reinterpret_cast<const float4*>(red + (w8 * LSTM_BG + fbb) * COLS + fu * 4);
                        s4.x += r4.x; s4.y += r4.y; s4.z += r4.z; s4.w += r4.w;
                    }
                    g4[0] += s4.x; g4[1] += s4.y; g4[2] += s4.z; g4[3] += s4.w;
                }
                const float ig = sigmoidf_(g4[0]), fg = sigmoidf_(g4[1]), gg = tanhf(g4[2]), og = sigmoidf_(g4[3]);
                float* cp = cS + (bg * LSTM_BG + fbb) * UNITS + fu;
                const float c = fg * (*cp) + ig * gg;
                *cp = c;
                const float h = og * tanhf(c);
                const long long o = ((long long)b * T + t) * H + j;
                __stcg(p.h_seq + o, h);
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
            if (nbg > 1) __syncthreads();    // Hs / red are reused by the next clip group
        }
        // ---- publish h_t: every thread's stores -> fence -> one release-add per CTA
        if (t + 1 < T) {
            __threadfence();
            __syncthreads();
            if (tid == 0) { __threadfence(); atomicAdd(p.barrier, 1u); }
        }
    }
}

size_t lstm_seq_smem_bytes(int H, int B, int units) {
    const int nbg = (B + LSTM_BG - 1) / LSTM_BG;
    return ((size_t)H * 4 * units + (size_t)LSTM_BG * H + 8 * 4 * units * LSTM_BG + (size_t)nbg * LSTM_BG * units) * sizeof(float);
}

int lstm_pick_units(int H) {
    // largest slice that fits shared memory while keeping >= 64 CTAs busy when H allows it
    if (H % 8 == 0 && lstm_seq_smem_bytes(H, 16, 8) <= 220 * 1024 && H / 8 >= 96) return 8;
    if (H % 4 == 0 && lstm_seq_smem_bytes(H, 16, 4) <= 220 * 1024) return 4;
    return 0;
}

template <int UNITS>
static cudaError_t launch_seq(const LstmSeqParams& p, cudaStream_t st) {
    const size_t smem = lstm_seq_smem_bytes(p.H, p.B, UNITS);
    auto kern = lstm_seq_kernel<UNITS>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    if (smem > 225 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaMemsetAsync(p.barrier, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    dim3 grid(p.H / UNITS), block(256);
    LstmSeqParams pc = p;
    void* args[] = {&pc};
    return cudaLaunchCooperativeKernel((void*)kern, grid, block, args, smem, st);
}

cudaError_t launch_lstm_seq(const LstmSeqParams& p, cudaStream_t st) {
    if (p.H % 4 != 0) return cudaErrorInvalidValue;
    const int units = lstm_pick_units(p.H);
    if (units == 8) return launch_seq<8>(p, st);
    if (units == 4) return launch_seq<4>(p, st);
    return cudaErrorInvalidConfiguration;
}

}  // namespace fcb
