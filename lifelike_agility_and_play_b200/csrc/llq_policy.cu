// llq_policy.cu -- on-device forward of the PMC policy (include/llq_policy.h), sm_100a.
//
// One CTA (256 threads) owns a tile of M = 32 observation rows.  Activations live in shared memory, transposed
// ([feature][row], so one float4 load hands a thread four rows of the same feature); the weights (0.96 MB fp32 in total)
// stream through L2, coalesced: in a layer with `nout` outputs thread j < nout owns output neuron j for all 32 rows
// (32 fp32 accumulators), i.e. per input feature one weight load, eight LDS.128 and 32 FFMA.  Two register-tiled variants
// (4 rows x 8 outputs per thread; two neurons x 32 rows on 128 threads) measured slower on B200 (0.21 / 0.25 ms against
// 0.137 ms at 4096 rows): with one 92 KB CTA per SM the kernel lives on latency hiding across its 8 warps, not on LSU economy.  At 4096 envs that is one
// wave of 128 CTAs, 2.1 GFLOP per launch in fp32 -- the actions feed the physics, so the layers stay in fp32 rather than
// TF32 tensor-core arithmetic (next step: 3xTF32 on tcgen05).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <new>
#include <string>
#include "../../include/llq.h"
#include "../../include/llq_policy.h"

namespace {

constexpr int M = 32, THREADS = 256;
constexpr int N_PROP = 135, N_FUT = 72, N_OBS = 207, H = 256, Z = 32, NCODE = 256, PE = 64, ZE = 32, NACT = 12;

struct Weights {   // device pointers into one allocation
  const float *prop_mean, *prop_std, *fut_mean, *fut_std;
  const float *e1w, *e1b, *e2w, *e2b, *e3w, *e3b, *code;
  const float *pew, *peb, *zew, *zeb;
  const float *d1w, *d1b, *d2w, *d2b, *d3w, *d3b;
};

// out[j][m] = act(b[j] + sum_k in[k][m] * W[k][j]),  in / out in shared memory ([feature][M])
template <bool RELU>
__device__ __forceinline__ void dense(const float* in, int K, const float* __restrict__ W, const float* __restrict__ b, int nout, float* out) {
  const int j = threadIdx.x;
  if (j < nout) {
    float acc[M];
    const float bj = b[j];
#pragma unroll
    for (int m = 0; m < M; m++) acc[m] = bj;
    // the weights come from L2 (~300 cycles): keep PF of them in flight per thread
    constexpr int PF = 16;
    int k0 = 0;
    for (; k0 + PF <= K; k0 += PF) {
      float wv[PF];
#pragma unroll
      for (int t = 0; t < PF; t++) wv[t] = __ldg(W + (size_t)(k0 + t) * nout + j);
#pragma unroll
      for (int t = 0; t < PF; t++) {
        const float w = wv[t];
        const float4* a4 = reinterpret_cast<const float4*>(in + (k0 + t) * M);
#pragma unroll
        for (int m4 = 0; m4 < M / 4; m4++) {
          const float4 a = a4[m4];
          acc[4 * m4] = fmaf(a.x, w, acc[4 * m4]); acc[4 * m4 + 1] = fmaf(a.y, w, acc[4 * m4 + 1]);
          acc[4 * m4 + 2] = fmaf(a.z, w, acc[4 * m4 + 2]); acc[4 * m4 + 3] = fmaf(a.w, w, acc[4 * m4 + 3]);
        }
      }
    }
    for (int k = k0; k < K; k++) {
      const float w = __ldg(W + (size_t)k * nout + j);
      const float4* a4 = reinterpret_cast<const float4*>(in + k * M);
#pragma unroll
      for (int m4 = 0; m4 < M / 4; m4++) {
        const float4 a = a4[m4];
        acc[4 * m4] = fmaf(a.x, w, acc[4 * m4]); acc[4 * m4 + 1] = fmaf(a.y, w, acc[4 * m4 + 1]);
        acc[4 * m4 + 2] = fmaf(a.z, w, acc[4 * m4 + 2]); acc[4 * m4 + 3] = fmaf(a.w, w, acc[4 * m4 + 3]);
      }
    }
    float4* o4 = reinterpret_cast<float4*>(out + j * M);
#pragma unroll
    for (int m4 = 0; m4 < M / 4; m4++) {
      float4 v = make_float4(acc[4 * m4], acc[4 * m4 + 1], acc[4 * m4 + 2], acc[4 * m4 + 3]);
      if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      o4[m4] = v;
    }
  }
}

__global__ void __launch_bounds__(THREADS) pmc_policy_kernel(const float* __restrict__ obs, long long ld, int n, Weights w,
                                                             float* __restrict__ act, int* __restrict__ codes) {
  extern __shared__ __align__(16) float sm[];
  float* x = sm;                       // [207][M] normalised observation
  float* h1 = x + N_OBS * M;           // [256][M]
  float* h2 = h1 + H * M;              // [256][M]
  __shared__ int s_code[M];
  const int row0 = blockIdx.x * M;
  const int tid = threadIdx.x;
  // ---- normalise + clip (pmc_net.py:130-137): coalesced along the row, transposed into shared memory
  for (int idx = tid; idx < M * N_OBS; idx += THREADS) {
    const int m = idx / N_OBS, k = idx - m * N_OBS;
    const int row = row0 + m < n ? row0 + m : n - 1;
    const float v = obs[(size_t)row * ld + k];
    const float mean = k < N_PROP ? w.prop_mean[k] : w.fut_mean[k - N_PROP];
    const float sd = k < N_PROP ? w.prop_std[k] : w.fut_std[k - N_PROP];
    x[k * M + m] = fminf(fmaxf((v - mean) / (sd + 1e-8f), -5.0f), 5.0f);
  }
  __syncthreads();
  // ---- VQ encoder 207 -> 256 -> 256 -> 32
  dense<true>(x, N_OBS, w.e1w, w.e1b, H, h1);
  __syncthreads();
  dense<true>(h1, H, w.e2w, w.e2b, H, h2);
  __syncthreads();
  dense<false>(h2, H, w.e3w, w.e3b, Z, h1);              // z in h1[0..31][M]
  __syncthreads();
  // ---- nearest code: thread c owns code c, squared distance to all M rows; then per-row argmin (first index wins ties)
  {
    float d[M];
#pragma unroll
    for (int m = 0; m < M; m++) d[m] = 0.f;
    for (int k = 0; k < Z; k++) {
      const float c = w.code[k * NCODE + tid];
      const float4* z4 = reinterpret_cast<const float4*>(h1 + k * M);
#pragma unroll
      for (int m4 = 0; m4 < M / 4; m4++) {
        const float4 z = z4[m4];
        float t;
        t = z.x - c; d[4 * m4] = fmaf(t, t, d[4 * m4]);
        t = z.y - c; d[4 * m4 + 1] = fmaf(t, t, d[4 * m4 + 1]);
        t = z.z - c; d[4 * m4 + 2] = fmaf(t, t, d[4 * m4 + 2]);
        t = z.w - c; d[4 * m4 + 3] = fmaf(t, t, d[4 * m4 + 3]);
      }
    }
#pragma unroll
    for (int m = 0; m < M; m++) h2[tid * M + m] = d[m];   // h2[c][m]
  }
  __syncthreads();
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int m = warp; m < M; m += THREADS / 32) {
      float best = 3.4e38f; int bi = 0;
      for (int c = lane; c < NCODE; c += 32) {
        const float v = h2[c * M + m];
        if (v < best) { best = v; bi = c; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) { s_code[m] = bi; if (codes && row0 + m < n) codes[row0 + m] = bi; }
    }
  }
  __syncthreads();
  // quantised code -> h2[0..31][M]; z_embed 32 -> 32 into h1[64..95][M]; prop_embed 135 -> 64 into h1[0..63][M]
  for (int idx = tid; idx < Z * M; idx += THREADS) {
    const int k = idx / M, m = idx - k * M;
    h2[k * M + m] = w.code[k * NCODE + s_code[m]];
  }
  __syncthreads();
  dense<true>(x, N_PROP, w.pew, w.peb, PE, h1);
  dense<true>(h2, Z, w.zew, w.zeb, ZE, h1 + PE * M);
  __syncthreads();
  // ---- decoder 96 -> 256 -> 256 -> 12
  dense<true>(h1, PE + ZE, w.d1w, w.d1b, H, h2);
  __syncthreads();
  dense<true>(h2, H, w.d2w, w.d2b, H, h1);
  __syncthreads();
  dense<false>(h1, H, w.d3w, w.d3b, NACT, h2);
  __syncthreads();
  for (int idx = tid; idx < M * NACT; idx += THREADS) {
    const int m = idx / NACT, j = idx - m * NACT;
    if (row0 + m < n) act[(size_t)(row0 + m) * NACT + j] = h2[j * M + m];
  }
}

thread_local std::string g_err;
int fail(int code, const char* msg) { g_err = msg; return code; }

}  // namespace

struct llq_policy { int device; float* d_w; Weights w; };

extern "C" {

const char* llq_policy_last_error(void) { return g_err.c_str(); }

int llq_policy_create(const float* weights, int64_t n_weights, int32_t device, llq_policy_handle* out) {
  if (!weights || !out) return fail(LLQ_EINVAL, "null argument");
  if (n_weights != LLQ_POLICY_N_WEIGHTS) return fail(LLQ_EINVAL, "weight blob has the wrong length (include/llq_policy.h)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(LLQ_ECUDA, "no CUDA device visible (no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(LLQ_EINVAL, "device ordinal out of range");
  llq_policy* h = new (std::nothrow) llq_policy();
  if (!h) return fail(LLQ_ENOMEM, "out of memory");
  h->device = device;
  cudaSetDevice(device);
  // device copy with every array padded to a 16-byte boundary (the tiled layers read the weights as float4)
  const size_t sizes[21] = {N_PROP, N_PROP, N_FUT, N_FUT, (size_t)N_OBS * H, H, (size_t)H * H, H, (size_t)H * Z, Z, (size_t)Z * NCODE,
                            (size_t)N_PROP * PE, PE, (size_t)Z * ZE, ZE, (size_t)(PE + ZE) * H, H, (size_t)H * H, H, (size_t)H * NACT, NACT};
  size_t off[22]; off[0] = 0;
  for (int i = 0; i < 21; i++) off[i + 1] = off[i] + ((sizes[i] + 3) & ~(size_t)3);
  if (cudaMalloc(&h->d_w, sizeof(float) * off[21]) != cudaSuccess) { delete h; return fail(LLQ_ECUDA, "cudaMalloc failed"); }
  {
    const float* src = weights;
    for (int i = 0; i < 21; i++) {
      if (cudaMemcpy(h->d_w + off[i], src, sizeof(float) * sizes[i], cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(h->d_w); delete h; return fail(LLQ_ECUDA, "weight upload failed");
      }
      src += sizes[i];
    }
  }
  int ai = 0;
  auto take = [&](size_t) { return (const float*)(h->d_w + off[ai++]); };
  Weights& w = h->w;
  w.prop_mean = take(N_PROP); w.prop_std = take(N_PROP); w.fut_mean = take(N_FUT); w.fut_std = take(N_FUT);
  w.e1w = take(N_OBS * H); w.e1b = take(H); w.e2w = take(H * H); w.e2b = take(H); w.e3w = take(H * Z); w.e3b = take(Z);
  w.code = take(Z * NCODE);
  w.pew = take(N_PROP * PE); w.peb = take(PE); w.zew = take(Z * ZE); w.zeb = take(ZE);
  w.d1w = take((PE + ZE) * H); w.d1b = take(H); w.d2w = take(H * H); w.d2b = take(H); w.d3w = take(H * NACT); w.d3b = take(NACT);
  const int smem = (int)(sizeof(float) * (N_OBS + 2 * H) * M);
  if (cudaFuncSetAttribute(pmc_policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
    cudaFree(h->d_w); delete h; return fail(LLQ_ECUDA, "cannot reserve shared memory for the policy kernel");
  }
  *out = h;
  return LLQ_OK;
}

int llq_policy_destroy(llq_policy_handle h) {
  if (!h) return LLQ_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_w);
  delete h;
  return LLQ_OK;
}

int llq_policy_forward(llq_policy_handle h, const float* d_obs, int64_t obs_ld, int32_t n, float* d_actions, int32_t* d_codes, void* stream) {
  if (!h || !d_obs || !d_actions || n <= 0 || obs_ld < N_OBS) return fail(LLQ_EINVAL, "bad arguments");
  cudaSetDevice(h->device);
  const int smem = (int)(sizeof(float) * (N_OBS + 2 * H) * M);
  pmc_policy_kernel<<<(n + M - 1) / M, THREADS, smem, (cudaStream_t)stream>>>(d_obs, (long long)obs_ld, n, h->w, d_actions, d_codes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LLQ_ECUDA, cudaGetErrorString(e));
  return LLQ_OK;
}

}  // extern "C"
