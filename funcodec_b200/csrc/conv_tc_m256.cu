// EXPERIMENTAL (option "tc_m256", off by default, not yet run on hardware): the deep-layer variant of conv_tc.cu.
//
// Why: for C_in >= 256 the N = 128 kernel is limited by weight ingress -- one (chunk, tap) slab is 32 KB (hi + lo) and
// feeds only 12 MMAs = 768 tensor cycles, 42 B/clk per SM, more than L2 can deliver to 148 SMs (tensor pipe 44 %,
// profiles/conv_ncu_r1n.txt).  Here a CTA owns M = 256 time rows as two 128-row halves, one per producer group, and every
// weight slab (N = 64: 16 KB) is used for both halves: 24 MMAs (768 tensor cycles) per 16 KB = 21 B/clk per SM.
// TMEM: per half two ping-pong accumulators + running totals, 6 x 64 = 384 columns (N = 128 would need 768).
// Known cap: an M128 N64 K8 tf32 MMA reads 4 KB of A + 2 KB of B from shared memory for 32.8 tensor cycles; at 128 B/clk
// of operand bandwidth that is ~48 cycles, so this variant tops out near 68 % of the tensor pipe (vs the 44 % measured
// with the weight stream as the limit); the real fix for 100 % is cta_group::2, where the CTA pair shares the operands.
//
// Same contract as conv1d_tc_kernel (1-D layers only, no FREQ / STAGE modes, weights always streamed through the ring):
// fused [GroupNorm apply + resblock add + ELU + reflect pad] on the input, bias + raw store + GroupNorm partials on the
// output, 3xTF32 split, accumulation chains cut every ~48 MMAs and folded with round-to-nearest adds.
// Reference semantics: funcodec/modules/normed_modules/conv.py:243-261 / :281-305.
#include "common.cuh"
#include "kernels.h"
#include "tc_sm100.cuh"

namespace fcb {

using namespace tc;

namespace {

constexpr int M2_HALF = 128;       // rows per half (one MMA M)
constexpr int M2_N = 64;           // output channels per CTA tile
constexpr int M2_KC = 32;
constexpr int M2_THREADS = 704;    // 2 x 8 producer warps (half 0 / half 1), copy warp, MMA warp, 4 accumulator warps
constexpr int M2_PROD = 256;
constexpr uint32_t M2_TMEM_COLS = 512;

struct M2Layout {
    int a_rows, a_stage, b_stage, na, nb, off_b, off_bar, total;
};

__host__ __device__ inline M2Layout m2_layout(int K, int S, int na, int nb) {
    M2Layout L;
    const int qmax = (K - 1) / S;
    L.a_rows = ((M2_HALF + qmax + 7) / 8) * 8;
    L.a_stage = 2 * L.a_rows * 128;
    L.b_stage = 2 * M2_N * 128;
    L.na = na; L.nb = nb;
    L.off_b = na * L.a_stage;
    L.off_bar = L.off_b + nb * L.b_stage;
    L.total = L.off_bar + 8 * (2 * na + 2 * nb + 4) + 96;
    return L;
}

__host__ __device__ inline int m2_units_per_group(int K, int S, int group_mmas) {
    const int taps = (K + S - 1) / S;
    const int g = group_mmas / (12 * taps);
    return g < 1 ? 1 : g;
}

__device__ __forceinline__ float m2_elu(float v) { return v > 0.f ? v : (__expf(v) - 1.0f); }

__global__ void __launch_bounds__(M2_THREADS, 1) conv1d_tc_m256_kernel(const ConvParams p, const int na_stages, const int nb_stages,
                                                                      const int n_tiles, const int group_mmas) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C_in = p.C_in, K = p.K, S = p.S;
    const bool has1 = p.in1.x != nullptr;
    const M2Layout L = m2_layout(K, S, na_stages, nb_stages);
    const int n_chunks = (C_in + M2_KC - 1) / M2_KC;
    const int n_units = n_chunks * S;
    const int upg = m2_units_per_group(K, S, group_mmas);
    const int n_groups = (n_units + upg - 1) / upg;
    const int n_tt2 = (p.T_out + 2 * M2_HALF - 1) / (2 * M2_HALF);
    const int n_nt = p.C_out / M2_N;
    const int n_pairs = na_stages / 2;                  // A ring depth per half

    uint8_t* smA = smem_raw;
    uint8_t* smB = smem_raw + L.off_b;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.off_bar);
    uint64_t* a_full = bars;                       // [na]  stage 2*j + half: 256 producer arrivals of that half's group
    uint64_t* a_empty = a_full + na_stages;        // [na]  tcgen05.commit
    uint64_t* b_full = a_empty + na_stages;        // [nb]  expect_tx
    uint64_t* b_empty = b_full + nb_stages;        // [nb]  tcgen05.commit
    uint64_t* acc_full = b_empty + nb_stages;      // [2]   tcgen05.commit (both halves of the group)
    uint64_t* acc_empty = acc_full + 2;            // [2]   128 accumulator-warp arrivals
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
    double* red = reinterpret_cast<double*>(tmem_ptr + 2);

    if (tid == 0) {
        for (int i = 0; i < na_stages; ++i) { mbar_init(a_full + i, M2_PROD); mbar_init(a_empty + i, 1); }
        for (int i = 0; i < nb_stages; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 128); }
        mbar_fence_init();
    }
    if (warp == 16) tmem_alloc(tmem_ptr, M2_TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    // TMEM column of region r (0/1 ping-pong, 2 totals) of half h
    auto region = [&](int h, int r) -> uint32_t { return tmem_base + (uint32_t)((h * 3 + r) * M2_N); };

    if (warp < 16) {
        // =========================================================== producers: group = half of the 256-row tile
        const int half = warp >> 3;
        const int ptid = tid & (M2_PROD - 1);
        const int jchunk = ptid & 7, rsub = ptid >> 3;
        const int gt_max = (p.T_out - 1) * S - p.pad_l + (K - 1);
        int pj = 0;                                    // ring index of this half's next stage
        uint32_t aphase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int r0 = tile / n_nt;
            const int tt2 = r0 % n_tt2, b = r0 / n_tt2;
            const int t0 = tt2 * 2 * M2_HALF + half * M2_HALF;
            const float* x0 = p.in0.x + (long long)b * p.in0.clip_stride + (long long)p.in0.row_off * C_in;
            const float* x1 = has1 ? p.in1.x + (long long)b * p.in1.clip_stride + (long long)p.in1.row_off * C_in : nullptr;
            const float* cf0 = p.in0.coef ? p.in0.coef + (long long)b * 2 * C_in : nullptr;
            const float* cf1 = (has1 && p.in1.coef) ? p.in1.coef + (long long)b * 2 * C_in : nullptr;
            for (int unit = 0; unit < n_units; ++unit) {
                const int chunk = unit / S, ph = unit - chunk * S;
                const int as = 2 * pj + half;
                uint8_t* hi = smA + as * L.a_stage;
                uint8_t* lo = hi + L.a_rows * 128;
                const int c = chunk * M2_KC + jchunk * 4;
                const bool c_ok = c < C_in;
                float4 a0 = make_float4(1.f, 1.f, 1.f, 1.f), b0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b1 = b0;
                if (!c_ok) { a0 = b0; a1 = b0; }
                else if (cf0) { a0 = __ldg(reinterpret_cast<const float4*>(cf0 + c)); b0 = __ldg(reinterpret_cast<const float4*>(cf0 + C_in + c)); }
                if (c_ok && cf1) { a1 = __ldg(reinterpret_cast<const float4*>(cf1 + c)); b1 = __ldg(reinterpret_cast<const float4*>(cf1 + C_in + c)); }
                constexpr int NR = 5;
                float4 xa[NR], xb[NR];
                bool okr[NR];
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int u = rsub + 32 * i;
                    const int gt = (t0 + u) * S + ph - p.pad_l;
                    bool ok = c_ok && u < L.a_rows && gt <= gt_max;
                    int src = gt;
                    if (p.pad_zero) ok = ok && gt >= 0 && gt < p.T_in;
                    else { src = reflect_index(gt, p.T_ext); ok = ok && src < p.T_in && src >= 0; }
                    okr[i] = ok;
                    xa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    xb[i] = xa[i];
                    if (ok) {
                        const long long off = (long long)src * C_in + c;
                        xa[i] = __ldg(reinterpret_cast<const float4*>(x0 + off));
                        if (has1) xb[i] = __ldg(reinterpret_cast<const float4*>(x1 + off));
                    }
                }
                mbar_wait_backoff(a_empty + as, aphase ^ 1, 64);
#pragma unroll
                for (int i = 0; i < NR; ++i) {
                    const int u = rsub + 32 * i;
                    if (u < L.a_rows) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (okr[i]) {
                            const float4 xv = xa[i];
                            v.x = fmaf(xv.x, a0.x, b0.x); v.y = fmaf(xv.y, a0.y, b0.y);
                            v.z = fmaf(xv.z, a0.z, b0.z); v.w = fmaf(xv.w, a0.w, b0.w);
                            if (has1) {
                                const float4 yv = xb[i];
                                v.x = v.x + fmaf(yv.x, a1.x, b1.x); v.y = v.y + fmaf(yv.y, a1.y, b1.y);
                                v.z = v.z + fmaf(yv.z, a1.z, b1.z); v.w = v.w + fmaf(yv.w, a1.w, b1.w);
                            }
                            if (p.elu) { v.x = m2_elu(v.x); v.y = m2_elu(v.y); v.z = m2_elu(v.z); v.w = m2_elu(v.w); }
                        }
                        float4 h, l;
                        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y);
                        split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
                        const uint32_t o = (uint32_t)u * 128u + (uint32_t)((jchunk ^ (u & 7)) << 4);
                        *reinterpret_cast<float4*>(hi + o) = h;
                        *reinterpret_cast<float4*>(lo + o) = l;
                    }
                }
                fence_proxy_async_smem();
                mbar_arrive(a_full + as);
                if (++pj == n_pairs) { pj = 0; aphase ^= 1; }
            }
        }
    } else if (warp == 16) {
        // =========================================================== weight slabs via the bulk-copy engine
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)L.b_stage;
            int bs = 0;
            uint32_t bphase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int nt = tile % n_nt;
                const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.w_tc) + (long long)nt * n_chunks * K * bytes;
                int chunk = 0, ph = 0;
                for (int unit = 0; unit < n_units; ++unit) {
                    for (int k = ph; k < K; k += S) {
                        mbar_wait_backoff(b_empty + bs, bphase ^ 1, 64);
                        mbar_arrive_expect_tx(b_full + bs, bytes);
                        bulk_g2s(smB + bs * L.b_stage, wbase + ((long long)chunk * K + k) * bytes, bytes, b_full + bs);
                        if (++bs == nb_stages) { bs = 0; bphase ^= 1; }
                    }
                    if (++ph == S) { ph = 0; ++chunk; }
                }
            }
        }
    } else if (warp == 17) {
        // =========================================================== MMA issuer: every weight slab feeds both halves
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(M2_HALF, M2_N);
            const uint32_t a_base = smem_u32(smA), b_base = smem_u32(smB);
            int pj = 0, bs = 0;
            uint32_t aphase = 0, bphase = 0, gcount = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int ph = 0;
                for (int g = 0; g < n_groups; ++g, ++gcount) {
                    const int buf = (int)(gcount & 1);
                    mbar_wait(acc_empty + buf, ((gcount >> 1) & 1) ^ 1);
                    tc_fence_after_sync();
                    const uint32_t d0 = region(0, buf), d1 = region(1, buf);
                    uint32_t accum = 0;
                    const int u_end = min(n_units, (g + 1) * upg);
                    for (int unit = g * upg; unit < u_end; ++unit) {
                        mbar_wait(a_full + 2 * pj, aphase);
                        mbar_wait(a_full + 2 * pj + 1, aphase);
                        tc_fence_after_sync();
                        const uint32_t a0_hi = a_base + (2 * pj) * L.a_stage, a0_lo = a0_hi + L.a_rows * 128;
                        const uint32_t a1_hi = a_base + (2 * pj + 1) * L.a_stage, a1_lo = a1_hi + L.a_rows * 128;
                        int q = 0;
                        for (int k = ph; k < K; k += S, ++q) {
                            mbar_wait(b_full + bs, bphase);
                            tc_fence_after_sync();
                            const uint32_t b_hi0 = b_base + bs * L.b_stage;
                            const uint32_t b_lo0 = b_hi0 + M2_N * 128;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint64_t db_hi = make_desc_k_sw128(b_hi0 + ks * 32);
                                const uint64_t db_lo = make_desc_k_sw128(b_lo0 + ks * 32);
                                const uint64_t da0_hi = make_desc_k_sw128(a0_hi + q * 128 + ks * 32);
                                const uint64_t da0_lo = make_desc_k_sw128(a0_lo + q * 128 + ks * 32);
                                const uint64_t da1_hi = make_desc_k_sw128(a1_hi + q * 128 + ks * 32);
                                const uint64_t da1_lo = make_desc_k_sw128(a1_lo + q * 128 + ks * 32);
                                mma_tf32_ss(d0, da0_lo, db_hi, idesc, accum);
                                mma_tf32_ss(d0, da0_hi, db_lo, idesc, 1);
                                mma_tf32_ss(d0, da0_hi, db_hi, idesc, 1);
                                mma_tf32_ss(d1, da1_lo, db_hi, idesc, accum);
                                mma_tf32_ss(d1, da1_hi, db_lo, idesc, 1);
                                mma_tf32_ss(d1, da1_hi, db_hi, idesc, 1);
                                accum = 1;
                            }
                            mma_commit(b_empty + bs);
                            if (++bs == nb_stages) { bs = 0; bphase ^= 1; }
                        }
                        mma_commit(a_empty + 2 * pj);
                        mma_commit(a_empty + 2 * pj + 1);
                        if (++pj == n_pairs) { pj = 0; aphase ^= 1; }
                        if (++ph == S) ph = 0;
                    }
                    mma_commit(acc_full + buf);
                }
            }
        }
    } else {
        // =========================================================== accumulator warps: fold groups, epilogue (both halves)
        const int quad = warp & 3;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        uint32_t gcount = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int nt = tile % n_nt;
            const int r0 = tile / n_nt;
            const int tt2 = r0 % n_tt2, b = r0 / n_tt2;
            const float* bias = p.bias + nt * M2_N;
            float s = 0.f, ss = 0.f;
            for (int g = 0; g < n_groups; ++g, ++gcount) {
                const int buf = (int)(gcount & 1);
                const bool last = (g == n_groups - 1);
                mbar_wait_backoff(acc_full + buf, (gcount >> 1) & 1, 128);
                tc_fence_after_sync();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int t = tt2 * 2 * M2_HALF + h * M2_HALF + quad * 32 + lane;
                    const bool row_ok = t < p.T_out;
                    float* orow = p.out + (long long)b * p.out_clip_stride + (long long)t * p.C_out + (long long)nt * M2_N;
#pragma unroll
                    for (int c0 = 0; c0 < M2_N; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(region(h, buf) + lane_base + (uint32_t)c0, v);
                        if (g > 0) {
                            uint32_t tv[32];
                            tmem_ld_32x32b_x32(region(h, 2) + lane_base + (uint32_t)c0, tv);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(tv[j]) + __uint_as_float(v[j]));
                        } else {
                            tmem_ld_wait();
                        }
                        if (!last) {
                            tmem_st_32x32b_x32(region(h, 2) + lane_base + (uint32_t)c0, v);
                        } else if (row_ok) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 o;
                                o.x = __uint_as_float(v[j + 0]) + __ldg(bias + c0 + j + 0);
                                o.y = __uint_as_float(v[j + 1]) + __ldg(bias + c0 + j + 1);
                                o.z = __uint_as_float(v[j + 2]) + __ldg(bias + c0 + j + 2);
                                o.w = __uint_as_float(v[j + 3]) + __ldg(bias + c0 + j + 3);
                                s += (o.x + o.y) + (o.z + o.w);
                                ss = fmaf(o.x, o.x, ss); ss = fmaf(o.y, o.y, ss); ss = fmaf(o.z, o.z, ss); ss = fmaf(o.w, o.w, ss);
                                *reinterpret_cast<float4*>(orow + c0 + j) = o;
                            }
                        }
                    }
                }
                if (!last) tmem_st_wait();
                tc_fence_before_sync();
                mbar_arrive(acc_empty + buf);
            }
            if (p.partials) {
                double ds = (double)s, dss = (double)ss;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    ds += __shfl_xor_sync(0xffffffffu, ds, o);
                    dss += __shfl_xor_sync(0xffffffffu, dss, o);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (lane == 0) { red[quad * 2] = ds; red[quad * 2 + 1] = dss; }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (quad == 0 && lane == 0) {
                    const int nparts = n_nt * n_tt2;
                    double* dst = p.partials + ((long long)b * nparts + nt * n_tt2 + tt2) * 2;
                    dst[0] = (red[0] + red[2]) + (red[4] + red[6]);
                    dst[1] = (red[1] + red[3]) + (red[5] + red[7]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, M2_TMEM_COLS);
    }
}

int g_m2_sms = 0;

}  // namespace

bool conv_tc_m256_supported(int C_in, int C_out_eff, int K, int S, int D) {
    return D == 1 && C_in % M2_KC == 0 && C_in >= 256 && C_out_eff % M2_N == 0 && K >= 1 && S >= 1 && ((K - 1) / S) <= 16;
}

int conv_tc_m256_num_parts(int T_out, int C_out_eff) {
    return ((T_out + 2 * M2_HALF - 1) / (2 * M2_HALF)) * (C_out_eff / M2_N);
}

// p.w_tc must be the n_tile = 64 slab image of the layer (engine.cu pack_tc), p.n_tile = 64.
cudaError_t launch_conv_tc_m256(const ConvParams& p, int B, cudaStream_t st, int* nparts) {
    if (p.n_tile != M2_N || p.fq.KF > 0) return cudaErrorInvalidValue;
    if (g_m2_sms == 0) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&g_m2_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    const int limit = 225 * 1024;
    int na = 4, nb = 2;
    M2Layout L = m2_layout(p.K, p.S, na, nb);
    if (L.total > limit) { na = 2; L = m2_layout(p.K, p.S, na, nb); }
    if (L.total > limit) return cudaErrorInvalidConfiguration;
    const int n_slabs = ((p.C_in + M2_KC - 1) / M2_KC) * p.K;
    while (nb < 12 && nb < n_slabs && m2_layout(p.K, p.S, na, nb + 1).total <= limit) L = m2_layout(p.K, p.S, na, ++nb);
    auto kern = conv1d_tc_m256_kernel;
    cudaError_t e = ensure_dynamic_smem((const void*)kern, limit);
    if (e != cudaSuccess) return e;
    const int n_tt2 = (p.T_out + 2 * M2_HALF - 1) / (2 * M2_HALF), n_nt = p.C_out / M2_N;
    *nparts = n_tt2 * n_nt;
    const int n_tiles = n_tt2 * n_nt * B;
    const int grid = n_tiles < g_m2_sms ? n_tiles : g_m2_sms;
    kern<<<grid, M2_THREADS, L.total, st>>>(p, na, nb, n_tiles, 48);
    return cudaGetLastError();
}

}  // namespace fcb
