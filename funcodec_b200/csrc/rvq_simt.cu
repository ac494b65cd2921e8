// Residual vector quantizer, all n_q stages in one kernel (fp32 SIMT reference implementation).
//
// Reference: DistributedResidualVectorQuantization.forward (eval) funcodec/modules/quantization/ddp_core_vq.py:367-418,
// EuclideanCodebook.quantize :180-188 (dist = -(|x|^2 - 2 x.C^T + |c|^2), first maximal index),
// dequantize :190-192, residual update :407-408, decode :442-453.
//
// One CTA owns 32 frames (rows) for every stage: the residual and the running quantized sum stay in shared
// memory, the stage's [K][D] codebook streams through shared memory in 128-codeword chunks (all 16.8 MB of
// codebooks are L2-resident), warp w scores rows 4w..4w+3 against the chunk with 4x4 register tiles, and the
// (value, index) argmin is a warp-shuffle reduction with the reference's first-index tie-break.
// The fp32 expression order of the reference is kept: t = (|x|^2 - 2*dot) + |c|^2, argmin t.
// The one-hot [M, K] tensor the reference materialises and discards in eval (:221) is never built.
// FLOPs per launch: 2 * (B*T') * K * D * n_q; bytes: x in, codes/quant out (tiny) -> tensor/FMA-bound.
#include "common.cuh"
#include "kernels.h"

namespace fcb {

constexpr int RVQ_ROWS = 32;
constexpr int RVQ_CHUNK = 128;

// SLICED (D too wide for a whole [128][D] codebook chunk next to the residual and the running sum, e.g. the SoundStream YAMLs'
// D = 512): the chunk is staged `ds` columns at a time and the 4x4 dot products keep accumulating across the slices -- the k order
// of every dot product, hence every bit of the result, is the same as in the unsliced kernel.
template <bool SLICED>
__global__ void __launch_bounds__(256, 2) rvq_kernel(const RvqParams p, const int ds) {
    extern __shared__ __align__(16) float smem[];
    const int D = p.D, K = p.K, T = p.T;
    const int pitch = D + 4;
    const int cpitch = SLICED ? ds + 4 : pitch;
    float* Xs = smem;                       // [32][pitch] residual
    float* Os = Xs + RVQ_ROWS * pitch;      // [32][pitch] quantized_out
    float* Cs = Os + RVQ_ROWS * pitch;      // [128][cpitch] codebook chunk (SLICED: `ds` of its D columns)
    float* xx = Cs + RVQ_CHUNK * cpitch;    // [32]
    int* best = reinterpret_cast<int*>(xx + RVQ_ROWS);   // [32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long M = (long long)p.B * T;
    const long long row0 = (long long)blockIdx.x * RVQ_ROWS;

    // ---- load the encoder output (GroupNorm applied on load), init quantized_out = 0
    for (int e = tid; e < RVQ_ROWS * D; e += 256) {
        const int r = e / D, d = e - r * D;
        const long long row = row0 + r;
        float v = 0.f;
        if (row < M) {
            const int b = (int)(row / T), t = (int)(row - (long long)b * T);
            v = p.in.x[(long long)b * p.in.clip_stride + (long long)(p.in.row_off + t) * D + d];
            if (p.in.stats) {
                const float mean = p.in.stats[2 * b], rstd = p.in.stats[2 * b + 1];
                const float a = rstd * p.in.gamma[d];
                v = fmaf(v, a, p.in.beta[d] - a * mean);
            }
            if (p.enc_out) p.enc_out[row * D + d] = v;
        }
        Xs[r * pitch + d] = v;
        Os[r * pitch + d] = 0.f;
    }
    __syncthreads();

    for (int q = 0; q < p.n_q; ++q) {
        const float* E = p.embed + (long long)q * K * D;
        const float* cn = p.cnorm + (long long)q * K;
        // |x|^2 per row: 8 lanes per row, fixed order
        {
            const int r = tid >> 3, part = tid & 7;
            float s = 0.f;
            for (int d = part; d < D; d += 8) { const float v = Xs[r * pitch + d]; s = fmaf(v, v, s); }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (part == 0) xx[r] = s;
        }
        float bval[4];
        int bidx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { bval[i] = 3.402823466e38f; bidx[i] = 0x7fffffff; }

        for (int c0 = 0; c0 < K; c0 += RVQ_CHUNK) {
            float dot[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dot[i][j] = 0.f;
            if constexpr (!SLICED) {
                __syncthreads();   // previous chunk consumed (and xx / residual updates visible)
                for (int e = tid * 4; e < RVQ_CHUNK * D; e += 256 * 4) {
                    const int c = e / D, d = e - c * D;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c0 + c < K) v = __ldg(reinterpret_cast<const float4*>(E + (long long)(c0 + c) * D + d));
                    *reinterpret_cast<float4*>(Cs + c * pitch + d) = v;
                }
                __syncthreads();
                const float* xr = Xs + (warp * 4) * pitch;
                const float* cr = Cs + lane * pitch;
                for (int k = 0; k < D; k += 4) {
                    float4 xv[4], cv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xr + i * pitch + k);
#pragma unroll
                    for (int j = 0; j < 4; ++j) cv[j] = *reinterpret_cast<const float4*>(cr + j * 32 * pitch + k);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            dot[i][j] = fmaf(xv[i].x, cv[j].x, dot[i][j]);
                            dot[i][j] = fmaf(xv[i].y, cv[j].y, dot[i][j]);
                            dot[i][j] = fmaf(xv[i].z, cv[j].z, dot[i][j]);
                            dot[i][j] = fmaf(xv[i].w, cv[j].w, dot[i][j]);
                        }
                }
            } else {
                for (int d0 = 0; d0 < D; d0 += ds) {
                    const int dw = min(ds, D - d0);          // columns of this slice (a multiple of 4)
                    __syncthreads();   // previous slice / chunk consumed (and xx / residual updates visible)
                    for (int e = tid * 4; e < RVQ_CHUNK * dw; e += 256 * 4) {
                        const int c = e / dw, d = e - c * dw;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c0 + c < K) v = __ldg(reinterpret_cast<const float4*>(E + (long long)(c0 + c) * D + d0 + d));
                        *reinterpret_cast<float4*>(Cs + c * cpitch + d) = v;
                    }
                    __syncthreads();
                    const float* xr = Xs + (warp * 4) * pitch + d0;
                    const float* cr = Cs + lane * cpitch;
                    for (int k = 0; k < dw; k += 4) {
                        float4 xv[4], cv[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xr + i * pitch + k);
#pragma unroll
                        for (int j = 0; j < 4; ++j) cv[j] = *reinterpret_cast<const float4*>(cr + j * 32 * cpitch + k);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                dot[i][j] = fmaf(xv[i].x, cv[j].x, dot[i][j]);
                                dot[i][j] = fmaf(xv[i].y, cv[j].y, dot[i][j]);
                                dot[i][j] = fmaf(xv[i].z, cv[j].z, dot[i][j]);
                                dot[i][j] = fmaf(xv[i].w, cv[j].w, dot[i][j]);
                            }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + lane + 32 * j;
                if (c < K) {
                    const float cc = __ldg(cn + c);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // reference order: (|x|^2 - 2*x.c) + |c|^2 ; 2*dot is exact
                        const float tval = __fadd_rn(__fsub_rn(xx[warp * 4 + i], 2.0f * dot[i][j]), cc);
                        if (tval < bval[i] || (tval == bval[i] && c < bidx[i])) { bval[i] = tval; bidx[i] = c; }
                    }
                }
            }
        }
        // ---- argmin across the 32 lanes, first index on ties
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = bval[i];
            int ix = bidx[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
                if (ov < v || (ov == v && oi < ix)) { v = ov; ix = oi; }
            }
            // a row whose distances are all NaN never replaces the sentinel: the reference's max() returns index 0 there
            if (lane == 0) best[warp * 4 + i] = (ix < 0 || ix >= K) ? 0 : ix;
        }
        __syncthreads();
        // ---- dequantize + residual update (ddp_core_vq.py:407-408)
        for (int e = tid; e < RVQ_ROWS * D; e += 256) {
            const int r = e / D, d = e - r * D;
            const long long row = row0 + r;
            if (row < M) {
                const int ix = best[r];
                const float cv = __ldg(E + (long long)ix * D + d);
                Xs[r * pitch + d] = Xs[r * pitch + d] - cv;
                Os[r * pitch + d] = Os[r * pitch + d] + cv;
                if (p.sub_quants) {
                    const int b = (int)(row / T), t = (int)(row - (long long)b * T);
                    p.sub_quants[(((long long)q * p.B + b) * D + d) * T + t] = cv;
                }
            }
        }
        if (tid < RVQ_ROWS && row0 + tid < M) p.codes[(long long)q * M + row0 + tid] = (long long)best[tid];
        __syncthreads();
    }
    if (p.quant) {
        for (int e = tid; e < RVQ_ROWS * D; e += 256) {
            const int r = e / D, d = e - r * D;
            if (row0 + r < M) p.quant[(row0 + r) * D + d] = Os[r * pitch + d];
        }
    }
}

__global__ void code_norms_kernel(const float* __restrict__ embed, float* __restrict__ cnorm, int rows, int D) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float s = 0.f;
    for (int d = 0; d < D; ++d) { const float v = embed[(long long)r * D + d]; s = fmaf(v, v, s); }
    cnorm[r] = s;
}

// DistributedResidualVectorQuantization.decode (ddp_core_vq.py:442-453): ((0 + C_0[i0]) + C_1[i1]) + ...
// codes layout: q_major ? [n_q][M] (encode side) : [M][n_q] (decode side, codec_basic.py:789)
__global__ void embed_sum_kernel(const long long* __restrict__ codes, int q_major, const float* __restrict__ embed, long long M,
                                 int n_q, int K, int D, float* __restrict__ out, int* __restrict__ err_flag) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * D) return;
    const long long row = e / D;
    const int d = (int)(e - row * D);
    float acc = 0.f;
    for (int q = 0; q < n_q; ++q) {
        const long long ix = q_major ? codes[(long long)q * M + row] : codes[row * n_q + q];
        if (ix < 0 || ix >= K) { if (err_flag) atomicExch(err_flag, 1); continue; }
        acc = acc + __ldg(embed + ((long long)q * K + ix) * D + d);
    }
    out[e] = acc;
}

constexpr size_t RVQ_SMEM_MAX = 200 * 1024;
static size_t rvq_smem_bytes(int D, int ds) {
    return ((size_t)2 * RVQ_ROWS * (D + 4) + (size_t)RVQ_CHUNK * (ds + 4) + RVQ_ROWS) * sizeof(float) + RVQ_ROWS * sizeof(int);
}
// 0: D does not fit at all; D: the whole-chunk kernel; else the slice width of the sliced kernel
int rvq_simt_slice(int D) {
    if (D % 4 != 0 || D < 4) return 0;
    if (rvq_smem_bytes(D, D) <= RVQ_SMEM_MAX) return D;
    for (int ds = 256; ds >= 32; ds >>= 1)
        if (rvq_smem_bytes(D, ds) <= RVQ_SMEM_MAX) return ds;
    return 0;
}

cudaError_t launch_rvq(const RvqParams& p, cudaStream_t st) {
    const int ds = rvq_simt_slice(p.D);
    if (ds == 0 || (ds != p.D && !p.allow_sliced)) return cudaErrorInvalidValue;
    const size_t smem = rvq_smem_bytes(p.D, ds);
    const long long M = (long long)p.B * p.T;
    const unsigned grid = (unsigned)((M + RVQ_ROWS - 1) / RVQ_ROWS);
    if (ds == p.D) {
        cudaError_t e = ensure_dynamic_smem((const void*)rvq_kernel<false>, (int)RVQ_SMEM_MAX);
        if (e != cudaSuccess) return e;
        rvq_kernel<false><<<grid, 256, smem, st>>>(p, ds);
    } else {
        cudaError_t e = ensure_dynamic_smem((const void*)rvq_kernel<true>, (int)RVQ_SMEM_MAX);
        if (e != cudaSuccess) return e;
        rvq_kernel<true><<<grid, 256, smem, st>>>(p, ds);
    }
    return cudaGetLastError();
}

cudaError_t launch_code_norms(const float* embed, float* cnorm, int rows, int D, cudaStream_t st) {
    code_norms_kernel<<<(rows + 127) / 128, 128, 0, st>>>(embed, cnorm, rows, D);
    return cudaGetLastError();
}

cudaError_t launch_embed_sum(const long long* codes, int q_major, const float* embed, int B, int T, int n_q, int K, int D,
                             float* out, int* err_flag, cudaStream_t st) {
    const long long M = (long long)B * T;
    const long long n = M * D;
    embed_sum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(codes, q_major, embed, M, n_q, K, D, out, err_flag);
    return cudaGetLastError();
}

}  // namespace fcb
