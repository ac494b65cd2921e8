// Residual vector quantizer with the nearest-codeword search on the tensor cores (tcgen05.mma kind::tf32,
// 3xTF32 split) and EXACT fp32 re-scoring of near-ties, all n_q stages in one kernel.
//
// Reference: DistributedResidualVectorQuantization.forward (eval) funcodec/modules/quantization/ddp_core_vq.py:367-418,
// EuclideanCodebook.quantize :180-188 (dist = -(|x|^2 - 2 x.C^T + |c|^2), first maximal index).
//
// One CTA owns 128 frames (rows) for every stage.
//   * the residual lives in shared memory as the 3xTF32 operand itself: hi and lo slabs (4 chunks of 32 dims,
//     canonical SWIZZLE_128B K-major), and hi + lo == the fp32 residual EXACTLY, so no separate copy is kept;
//   * per stage the [K][D] codebook streams through a shared-memory ring as pre-split, pre-swizzled slab images
//     (one cp.async.bulk per 128-codeword x 32-dim slab), the MMA warp produces dot[128 rows x 128 codewords]
//     tiles into two ping-pong TMEM accumulators (48 chained MMAs each), and the 4 epilogue warps (TMEM lane ==
//     row) evaluate t = (|x|^2 - 2*dot) + |c|^2 in the reference's fp32 order and keep the two best (value, index);
//   * the tensor-core dot carries ~1e-5 absolute error, so whenever best and runner-up are closer than
//     RESCORE_TOL both are re-scored with the exact sequential-fp32 dot product of the SIMT kernel (rvq_simt.cu) and
//     compared with the first-index tie-break: decisions equal the fp32 path's unless three candidates fall inside
//     the tolerance band;
//   * dequantize + residual update (ddp_core_vq.py:407-408) re-split the residual in place; the quantized sum is
//     rebuilt afterwards from the codes by embed_sum_kernel in the reference's accumulation order.
// FLOPs per launch: 2 * rows * K * D * n_q (x3 tensor passes); bytes: rows*D*4 in, codes out -> tensor-bound.
#include "common.cuh"
#include "kernels.h"
#include "tc_sm100.cuh"

namespace fcb {

using namespace tc;

constexpr int RQ_M = 128;           // rows per CTA
constexpr int RQ_N = RVQ_TC_N;      // codewords per MMA tile (measured: 64-wide tiles with a 4-deep ring are 35% slower)
constexpr int RQ_NB = 2;            // codebook slab ring depth
constexpr int RQ_THREADS = 192;     // 4 epilogue warps, copy warp, MMA warp
constexpr float RQ_RESCORE_TOL = 4e-3f;

struct RqSmem {
    int a_slab;        // bytes of one (hi or lo) chunk slab: 128 rows x 128 B
    int off_b, off_cc, off_xx, off_idx, off_bar, total;
};

__host__ __device__ inline RqSmem rq_layout(int D, int K) {
    RqSmem L;
    const int n_chunks = D / 32;
    L.a_slab = RQ_M * 128;
    L.off_b = 2 * n_chunks * L.a_slab;                // A: [chunk][hi|lo]
    L.off_cc = L.off_b + RQ_NB * 2 * RQ_N * 128;      // B ring: [stage][hi|lo][128 x 128 B]
    L.off_xx = L.off_cc + K * 4;
    L.off_idx = L.off_xx + RQ_M * 4;
    L.off_bar = (L.off_idx + RQ_M * 4 + 15) & ~15;
    L.total = L.off_bar + 8 * (2 * RQ_NB + 4 + 1) + 16;
    return L;
}

__global__ void __launch_bounds__(RQ_THREADS, 1) rvq_tc_kernel(const RvqParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D, K = p.K, T = p.T;
    const int n_chunks = D / 32;
    const int n_nt = K / RQ_N;
    const RqSmem L = rq_layout(D, K);
    const long long M = (long long)p.B * T;
    const long long row0 = (long long)blockIdx.x * RQ_M;

    uint8_t* smA = smem_raw;
    uint8_t* smB = smem_raw + L.off_b;
    float* cc_s = reinterpret_cast<float*>(smem_raw + L.off_cc);
    float* xx_s = reinterpret_cast<float*>(smem_raw + L.off_xx);
    int* idx_s = reinterpret_cast<int*>(smem_raw + L.off_idx);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);
    uint64_t* b_full = bars;                    // [RQ_NB]
    uint64_t* b_empty = b_full + RQ_NB;         // [RQ_NB]
    uint64_t* acc_full = b_empty + RQ_NB;       // [2]
    uint64_t* acc_empty = acc_full + 2;         // [2]
    uint64_t* a_ready = acc_empty + 2;          // [1] residual slabs (re)written for the stage
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_ready + 1);

    if (tid == 0) {
        for (int i = 0; i < RQ_NB; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 128); }
        mbar_init(a_ready, 128);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(tmem_ptr, 2 * RQ_N);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    const long long slab_bytes = 2LL * RQ_N * 128;      // one (n-tile, chunk) hi+lo image

    if (warp < 4) {
        const int jchunk = tid & 7, rsub = tid >> 3;    // (row, 16-byte chunk) mapping: 16 rows per pass
        // ---- load the encoder output (GroupNorm applied on load) into the hi/lo slabs
        for (int ch = 0; ch < n_chunks; ++ch) {
            uint8_t* hi = smA + (2 * ch) * L.a_slab;
            uint8_t* lo = hi + L.a_slab;
            for (int r = rsub; r < RQ_M; r += 16) {
                const long long row = row0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int d = ch * 32 + jchunk * 4;
                if (row < M) {
                    const int b = (int)(row / T), t = (int)(row - (long long)b * T);
                    v = __ldg(reinterpret_cast<const float4*>(p.in.x + (long long)b * p.in.clip_stride + (long long)(p.in.row_off + t) * D + d));
                    if (p.in.stats) {
                        const float mean = p.in.stats[2 * b], rstd = p.in.stats[2 * b + 1];
                        const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.in.gamma + d));
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.in.beta + d));
                        float a;
                        a = rstd * g4.x; v.x = fmaf(v.x, a, b4.x - a * mean);
                        a = rstd * g4.y; v.y = fmaf(v.y, a, b4.y - a * mean);
                        a = rstd * g4.z; v.z = fmaf(v.z, a, b4.z - a * mean);
                        a = rstd * g4.w; v.w = fmaf(v.w, a, b4.w - a * mean);
                    }
                    if (p.enc_out) *reinterpret_cast<float4*>(p.enc_out + row * D + d) = v;
                }
                float4 h, l;
                split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
                const uint32_t o = (uint32_t)r * 128u + (uint32_t)((jchunk ^ (r & 7)) << 4);
                *reinterpret_cast<float4*>(hi + o) = h;
                *reinterpret_cast<float4*>(lo + o) = l;
            }
        }
        const int quad = warp & 3;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const int myrow = quad * 32 + lane;                 // TMEM lane == row of this thread
        long long tcount = 0;                                 // global dist-tile counter (stage-major)
        for (int q = 0; q < p.n_q; ++q) {
            const float* E = p.embed + (long long)q * K * D;
            // ---- |x|^2 in the SIMT kernel's order (8 lanes per row, stride-8 dims, xor-shuffle 1,2,4) and |c|^2
            asm volatile("bar.sync 1, 128;" ::: "memory");    // previous stage's residual update is complete
            for (int r = rsub; r < RQ_M; r += 16) {
                float s = 0.f;
                for (int d = jchunk; d < D; d += 8) {
                    const uint32_t o = (uint32_t)r * 128u + (uint32_t)((((d & 31) >> 2) ^ (r & 7)) << 4) + (uint32_t)((d & 3) << 2);
                    const uint8_t* hi = smA + (2 * (d >> 5)) * L.a_slab;
                    const float v = *reinterpret_cast<const float*>(hi + o) + *reinterpret_cast<const float*>(hi + L.a_slab + o);
                    s = fmaf(v, v, s);
                }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                if (jchunk == 0) xx_s[r] = s;
            }
            for (int c = tid; c < K; c += 128) cc_s[c] = __ldg(p.cnorm + (long long)q * K + c);
            fence_proxy_async_smem();
            mbar_arrive(a_ready);                             // slabs of this stage are final -> MMA may start
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const float xx = xx_s[myrow];
            float v1 = 3.402823466e38f, v2 = 3.402823466e38f;
            int i1 = 0, i2 = -1;       // a NaN row replaces nothing: index 0, like the reference's max() over NaNs; no OOB gather
            for (int nt = 0; nt < n_nt; ++nt, ++tcount) {
                const int buf = (int)(tcount & 1);
                mbar_wait_backoff(acc_full + buf, (uint32_t)((tcount >> 1) & 1), 64);
                tc_fence_after_sync();
#pragma unroll
                for (int c0 = 0; c0 < RQ_N; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_base + lane_base + (uint32_t)(buf * RQ_N + c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int c = nt * RQ_N + c0 + j;
                        const float tv = __fadd_rn(__fsub_rn(xx, 2.0f * __uint_as_float(v[j])), cc_s[c]);
                        if (tv < v1) { v2 = v1; i2 = i1; v1 = tv; i1 = c; }
                        else if (tv < v2) { v2 = tv; i2 = c; }
                    }
                }
                tc_fence_before_sync();
                mbar_arrive(acc_empty + buf);
            }
            // ---- exact fp32 re-scoring of near-ties (sequential fmaf chain == rvq_simt.cu)
            int best = i1;
            if (row0 + myrow < M && v2 - v1 < RQ_RESCORE_TOL + 2e-5f * fabsf(v1) && i2 >= 0) {
                float d1 = 0.f, d2 = 0.f;
                const float* c1 = E + (long long)i1 * D;
                const float* c2 = E + (long long)i2 * D;
                for (int d = 0; d < D; ++d) {
                    const uint32_t o = (uint32_t)myrow * 128u + (uint32_t)((((d & 31) >> 2) ^ (myrow & 7)) << 4) + (uint32_t)((d & 3) << 2);
                    const uint8_t* hi = smA + (2 * (d >> 5)) * L.a_slab;
                    const float x = *reinterpret_cast<const float*>(hi + o) + *reinterpret_cast<const float*>(hi + L.a_slab + o);
                    d1 = fmaf(x, __ldg(c1 + d), d1);
                    d2 = fmaf(x, __ldg(c2 + d), d2);
                }
                const float t1 = __fadd_rn(__fsub_rn(xx, 2.0f * d1), cc_s[i1]);
                const float t2 = __fadd_rn(__fsub_rn(xx, 2.0f * d2), cc_s[i2]);
                if (t2 < t1 || (t2 == t1 && i2 < i1)) best = i2;
            }
            idx_s[myrow] = best;
            if (row0 + myrow < M) p.codes[(long long)q * M + row0 + myrow] = (long long)best;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // ---- dequantize + residual update (all MMAs of the stage have completed: last acc_full was waited on).
            // 8 passes of 16 rows; the codeword reads of pass i+1 (L2 latency) are in flight while pass i is re-split.
            {
                constexpr int MAXCH = 4;                       // D <= 128 on this path
                float4 cv[MAXCH], nv[MAXCH];
                auto load_c = [&](int r, float4 (&dst)[MAXCH]) {
#pragma unroll
                    for (int ch = 0; ch < MAXCH; ++ch) {
                        dst[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ch < n_chunks && row0 + r < M)
                            dst[ch] = __ldg(reinterpret_cast<const float4*>(E + (long long)idx_s[r] * D + ch * 32 + jchunk * 4));
                    }
                };
                load_c(rsub, cv);
                for (int r = rsub; r < RQ_M; r += 16) {
                    if (r + 16 < RQ_M) load_c(r + 16, nv);
                    const long long row = row0 + r;
                    if (row < M) {
                        const uint32_t o = (uint32_t)r * 128u + (uint32_t)((jchunk ^ (r & 7)) << 4);
#pragma unroll
                        for (int ch = 0; ch < MAXCH; ++ch) {
                            if (ch < n_chunks) {
                                uint8_t* hi = smA + (2 * ch) * L.a_slab;
                                uint8_t* lo = hi + L.a_slab;
                                const float4 h0 = *reinterpret_cast<const float4*>(hi + o);
                                const float4 l0 = *reinterpret_cast<const float4*>(lo + o);
                                const float4 c4 = cv[ch];
                                float4 x;
                                x.x = (h0.x + l0.x) - c4.x; x.y = (h0.y + l0.y) - c4.y; x.z = (h0.z + l0.z) - c4.z; x.w = (h0.w + l0.w) - c4.w;
                                float4 h, l;
                                split_tf32(x.x, h.x, l.x); split_tf32(x.y, h.y, l.y); split_tf32(x.z, h.z, l.z); split_tf32(x.w, h.w, l.w);
                                *reinterpret_cast<float4*>(hi + o) = h;
                                *reinterpret_cast<float4*>(lo + o) = l;
                                if (p.sub_quants) {
                                    const int b = (int)(row / T), t = (int)(row - (long long)b * T);
                                    const int d = ch * 32 + jchunk * 4;
                                    float* sq = p.sub_quants + (((long long)q * p.B + b) * D + d) * T + t;
                                    sq[0] = c4.x; sq[(long long)T] = c4.y; sq[2LL * T] = c4.z; sq[3LL * T] = c4.w;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int ch = 0; ch < MAXCH; ++ch) cv[ch] = nv[ch];
                }
            }
        }
    } else if (warp == 4) {
        // =========================================================== codebook slabs via the bulk-copy engine
        if (lane == 0) {
            long long it = 0;
            for (int q = 0; q < p.n_q; ++q) {
                const uint8_t* qbase = reinterpret_cast<const uint8_t*>(p.embed_tc) + (long long)q * n_nt * n_chunks * slab_bytes;
                for (int nt = 0; nt < n_nt; ++nt)
                    for (int ch = 0; ch < n_chunks; ++ch, ++it) {
                        const int bs = (int)(it % RQ_NB);
                        mbar_wait_backoff(b_empty + bs, (uint32_t)((it / RQ_NB) & 1) ^ 1, 64);
                        mbar_arrive_expect_tx(b_full + bs, (uint32_t)slab_bytes);
                        bulk_g2s(smB + bs * slab_bytes, qbase + ((long long)nt * n_chunks + ch) * slab_bytes, (uint32_t)slab_bytes, b_full + bs);
                    }
            }
        }
    } else {
        // =========================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(RQ_M, RQ_N);
            const uint32_t a_base = smem_u32(smA), b_base = smem_u32(smB);
            long long it = 0, tcount = 0;
            for (int q = 0; q < p.n_q; ++q) {
                mbar_wait(a_ready, (uint32_t)(q & 1));
                tc_fence_after_sync();
                for (int nt = 0; nt < n_nt; ++nt, ++tcount) {
                    const int buf = (int)(tcount & 1);
                    mbar_wait(acc_empty + buf, (uint32_t)((tcount >> 1) & 1) ^ 1);
                    tc_fence_after_sync();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * RQ_N);
                    uint32_t accum = 0;
                    for (int ch = 0; ch < n_chunks; ++ch, ++it) {
                        const int bs = (int)(it % RQ_NB);
                        mbar_wait(b_full + bs, (uint32_t)((it / RQ_NB) & 1));
                        tc_fence_after_sync();
                        const uint32_t a_hi0 = a_base + (2 * ch) * L.a_slab, a_lo0 = a_hi0 + L.a_slab;
                        const uint32_t b_hi0 = b_base + bs * (uint32_t)slab_bytes, b_lo0 = b_hi0 + RQ_N * 128;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t da_hi = make_desc_k_sw128(a_hi0 + ks * 32), da_lo = make_desc_k_sw128(a_lo0 + ks * 32);
                            const uint64_t db_hi = make_desc_k_sw128(b_hi0 + ks * 32), db_lo = make_desc_k_sw128(b_lo0 + ks * 32);
                            mma_tf32_ss(d_tmem, da_lo, db_hi, idesc, accum);
                            accum = 1;
                            mma_tf32_ss(d_tmem, da_hi, db_lo, idesc, 1);
                            mma_tf32_ss(d_tmem, da_hi, db_hi, idesc, 1);
                        }
                        mma_commit(b_empty + bs);
                    }
                    mma_commit(acc_full + buf);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 2 * RQ_N);
    }
}

bool rvq_tc_supported(int D, int K) {
    return D % 32 == 0 && D >= 32 && D <= 128 && K % RQ_N == 0 && rq_layout(D, K).total <= 225 * 1024;
}

cudaError_t launch_rvq_tc(const RvqParams& p, cudaStream_t st) {
    if (!rvq_tc_supported(p.D, p.K) || !p.embed_tc) return cudaErrorInvalidValue;
    const RqSmem L = rq_layout(p.D, p.K);
    {
        cudaError_t e = ensure_dynamic_smem((const void*)rvq_tc_kernel, 225 * 1024);
        if (e != cudaSuccess) return e;
    }
    const long long M = (long long)p.B * p.T;
    rvq_tc_kernel<<<(unsigned)((M + RQ_M - 1) / RQ_M), RQ_THREADS, L.total, st>>>(p);
    return cudaGetLastError();
}

}  // namespace fcb
