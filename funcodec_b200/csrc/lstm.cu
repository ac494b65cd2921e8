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
//   * per step every CTA needs the whole h_{t-1} of a clip group: it is exchanged through global memory (L2)
//     with a per-group release/acquire counter barrier; independent clip groups (8 clips) are software
//     pipelined so that one group's barrier + broadcast latency hides behind the other group's math;
//   * a thread accumulates the 4 gates of one unit for 8 clips (32 fp32 accumulators) over an interleaved
//     K slice (W rows via conflict-free LDS.128, h via broadcast LDS.128), K slices are reduced with
//     shuffles + one shared-memory pass, and UNITS*8 threads do the cell update.
// Latency-bound by construction (T' dependent steps); FLOPs = 2*B*T*4H*H per layer.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace fcb {

constexpr int LSTM_GB = 8;        // clips per work item (accumulator tile)
constexpr int LSTM_MAX_GROUPS = 64;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Work item (t, g) = timestep t of clip group g (8 clips).  Clip groups are independent sequences, so items are
// software-pipelined: while item i is being reduced / finalised, the h_{t-1} rows of item i+1 (another group,
// published by every CTA at least one item ago) are already in flight from L2 into registers.  Each group has
// its own release/acquire counter, so the grid barrier latency of one group hides behind the other's math.
template <int UNITS>
__global__ void __launch_bounds__(256, 1) lstm_seq_kernel(const LstmSeqParams p) {
    constexpr int COLS = 4 * UNITS;            // gate columns owned by this CTA
    constexpr int KS_PER_WARP = 32 / UNITS;    // K slices inside a warp
    constexpr int NSLICE = 8 * KS_PER_WARP;    // K slices per CTA (interleaved in groups of 4 k)
    constexpr int NLD = 8;                     // float4 prefetch registers per thread (H <= 1024)
    extern __shared__ __align__(16) float smem[];
    const int H = p.H, T = p.T, B = p.B;
    float* Ws = smem;                               // [H][COLS]
    float* Hs = Ws + (size_t)H * COLS;              // [2][LSTM_GB][H]
    float* red = Hs + 2 * LSTM_GB * H;              // [8 warps][LSTM_GB][COLS]
    float* cS = red + 8 * LSTM_GB * COLS;           // [ng][LSTM_GB][UNITS] cell state
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int u = lane % UNITS, ks = lane / UNITS;
    const int slice = warp * KS_PER_WARP + ks;
    const int j0 = blockIdx.x * UNITS;
    const int ng = (B + LSTM_GB - 1) / LSTM_GB;
    const int n_items = T * ng;
    const unsigned nctas = gridDim.x;

    for (int e = tid; e < H * COLS; e += 256) {
        const int k = e / COLS, c = e - k * COLS;
        Ws[e] = __ldg(p.whh + (long long)k * 4 * H + (long long)j0 * 4 + c);
    }
    for (int e = tid; e < ng * LSTM_GB * UNITS; e += 256) cS[e] = 0.f;
    __syncthreads();

    const bool fin = tid < LSTM_GB * UNITS;
    const int fbb = tid / UNITS, fu = tid % UNITS;
    float4 hv[NLD];

    for (int i = 0; i < n_items; ++i) {
        const int t = i / ng, g = i - t * ng;
        const int b0 = g * LSTM_GB;
        const int nb = min(LSTM_GB, B - b0);
        const float* Hc = Hs + (i & 1) * LSTM_GB * H;
        const int inext = i + 1;
        const bool has_next = inext < n_items;
        const int tn = inext / ng, gn = inext - tn * ng;
        const bool early = has_next && gn != g;

        float4 gxv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fin && fbb < nb)
            gxv = __ldcs(reinterpret_cast<const float4*>(p.gx + ((long long)(b0 + fbb) * T + t) * 4 * H + (long long)(j0 + fu) * 4));

        float acc[4][LSTM_GB];
#pragma unroll
        for (int gg = 0; gg < 4; ++gg)
#pragma unroll
            for (int bb = 0; bb < LSTM_GB; ++bb) acc[gg][bb] = 0.f;
        if (t > 0) {
            for (int k0 = slice * 4; k0 < H; k0 += NSLICE * 4) {
                float4 w[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(Ws + (k0 + kk) * COLS + u * 4);
#pragma unroll
                for (int bb = 0; bb < LSTM_GB; ++bb) {
                    const float4 h4 = *reinterpret_cast<const float4*>(Hc + bb * H + k0);
                    acc[0][bb] = fmaf(h4.x, w[0].x, acc[0][bb]); acc[1][bb] = fmaf(h4.x, w[0].y, acc[1][bb]);
                    acc[2][bb] = fmaf(h4.x, w[0].z, acc[2][bb]); acc[3][bb] = fmaf(h4.x, w[0].w, acc[3][bb]);
                    acc[0][bb] = fmaf(h4.y, w[1].x, acc[0][bb]); acc[1][bb] = fmaf(h4.y, w[1].y, acc[1][bb]);
                    acc[2][bb] = fmaf(h4.y, w[1].z, acc[2][bb]); acc[3][bb] = fmaf(h4.y, w[1].w, acc[3][bb]);
                    acc[0][bb] = fmaf(h4.z, w[2].x, acc[0][bb]); acc[1][bb] = fmaf(h4.z, w[2].y, acc[1][bb]);
                    acc[2][bb] = fmaf(h4.z, w[2].z, acc[2][bb]); acc[3][bb] = fmaf(h4.z, w[2].w, acc[3][bb]);
                    acc[0][bb] = fmaf(h4.w, w[3].x, acc[0][bb]); acc[1][bb] = fmaf(h4.w, w[3].y, acc[1][bb]);
                    acc[2][bb] = fmaf(h4.w, w[3].z, acc[2][bb]); acc[3][bb] = fmaf(h4.w, w[3].w, acc[3][bb]);
                }
            }
        }

        // ---- prefetch of the next item's h_{t-1} (issued here when it belongs to another clip group)
        auto fetch_next = [&]() {
            if (tn > 0) {
                if (tid == 0) {
                    while (ld_acquire_u32(p.barrier + gn) < (unsigned)tn * nctas) { }
                    __threadfence();
                }
                __syncthreads();
                const int bn0 = gn * LSTM_GB;
#pragma unroll
                for (int r = 0; r < NLD; ++r) {
                    const int e = (r * 256 + tid) * 4;
                    hv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e < LSTM_GB * H) {
                        const int bb = e / H, k = e - bb * H;
                        if (bn0 + bb < B)
                            hv[r] = __ldcg(reinterpret_cast<const float4*>(p.h_seq + ((long long)(bn0 + bb) * T + (tn - 1)) * H + k));
                    }
                }
            }
        };
        if (early) fetch_next();

        // ---- reduce the K slices, cell update
        if (t > 0) {
#pragma unroll
            for (int gg = 0; gg < 4; ++gg)
#pragma unroll
                for (int bb = 0; bb < LSTM_GB; ++bb) {
                    float v = acc[gg][bb];
#pragma unroll
                    for (int o = UNITS; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    acc[gg][bb] = v;
                }
            if (ks == 0) {
#pragma unroll
                for (int bb = 0; bb < LSTM_GB; ++bb)
                    *reinterpret_cast<float4*>(red + (warp * LSTM_GB + bb) * COLS + u * 4) =
                        make_float4(acc[0][bb], acc[1][bb], acc[2][bb], acc[3][bb]);
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
                    const float4 r4 = *reinterpret_cast<const float4*>(red + (w8 * LSTM_GB + fbb) * COLS + fu * 4);
                    s4.x += r4.x; s4.y += r4.y; s4.z += r4.z; s4.w += r4.w;
                }
                g4[0] += s4.x; g4[1] += s4.y; g4[2] += s4.z; g4[3] += s4.w;
            }
            const float ig = sigmoidf_(g4[0]), fg = sigmoidf_(g4[1]), gg = tanhf(g4[2]), og = sigmoidf_(g4[3]);
            float* cp = cS + (g * LSTM_GB + fbb) * UNITS + fu;
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
            __threadfence();
        }
        // ---- publish h_t of this group
        __syncthreads();
        if (tid == 0 && t + 1 < T) { __threadfence(); atomicAdd(p.barrier + g, 1u); }
        if (has_next && !early) fetch_next();
        if (has_next && tn > 0) {
            float* Hn = Hs + (inext & 1) * LSTM_GB * H;
#pragma unroll
            for (int r = 0; r < NLD; ++r) {
                const int e = (r * 256 + tid) * 4;
                if (e < LSTM_GB * H) *reinterpret_cast<float4*>(Hn + e) = hv[r];
            }
        }
        __syncthreads();
    }
}

size_t lstm_seq_smem_bytes(int H, int B, int units) {
    const int ng = (B + LSTM_GB - 1) / LSTM_GB;
    return ((size_t)H * 4 * units + (size_t)2 * LSTM_GB * H + 8 * 4 * units * LSTM_GB + (size_t)ng * LSTM_GB * units) * sizeof(float);
}

int lstm_pick_units(int H) {
    // largest slice that fits shared memory while keeping >= 96 CTAs busy when H allows it
    if (H > 1024) return 0;                                   // prefetch registers cover 8 clips x 1024
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
    if ((p.B + LSTM_GB - 1) / LSTM_GB > LSTM_MAX_GROUPS) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(p.barrier, 0, LSTM_MAX_GROUPS * sizeof(unsigned), st);
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
