// Small elementwise kernels around the stacks.
#include "common.cuh"
#include "kernels.h"

#include <map>
#include <mutex>
#include <utility>

namespace fcb {

cudaError_t ensure_dynamic_smem(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> done;     // (kernel, device) -> bytes granted
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(kernel, dev);
    auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done[key] = bytes;
    return e;
}

// Final GroupNorm(1,1) of the decoder's last conv (seanet_decoder.py:160-164 -> conv.py:162), optional
// `out * scale` (codec_basic.py:405-407) and the `[:, :, :L]` trim (codec_basic.py:711) in one pass.
__global__ void final_output_kernel(const float* __restrict__ raw, const float* __restrict__ stats,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ scale, int T_raw, int out_len, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= out_len) return;
    const float mean = stats[2 * b], rstd = stats[2 * b + 1];
    const float a = rstd * gamma[0];
    float v = fmaf(raw[(long long)b * T_raw + t], a, beta[0] - a * mean);
    if (scale) v = v * scale[b];
    out[(long long)b * out_len + t] = v;
}

__global__ void fill_kernel(float* p, float v, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

cudaError_t launch_final_output(const float* raw, const float* stats, const float* gamma, const float* beta,
                                const float* scale, int B, int T_raw, int out_len, float* out, cudaStream_t st) {
    final_output_kernel<<<dim3((out_len + 255) / 256, B), 256, 0, st>>>(raw, stats, gamma, beta, scale, T_raw, out_len, out);
    return cudaGetLastError();
}

cudaError_t launch_fill(float* p, float v, long long n, cudaStream_t st) {
    fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, v, n);
    return cudaGetLastError();
}

}  // namespace fcb
