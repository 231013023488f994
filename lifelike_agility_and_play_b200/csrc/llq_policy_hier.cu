// On-device inference of the reference's environmental- and strategic-level policies (include/llq_policy.h).
//
//   networks/legged_robot/epmc_net/epmc_net.py:86-135   perception encoders (2-D / 1-D convolutions, SAME + ReLU) and their fusion
//   networks/legged_robot/epmc_net/epmc_net.py:138-177  mlc_encoder: prop 135 -> 64 | command 120 -> 64, concat -> 256 -> LSTM(32, layer norm)
//                                                       -> 256 logits -> argmax -> column of the primitive-level codebook
//   networks/legged_robot/sepmc_net/sepmc_net.py:122-146 hlc_encoder: prop | perception 88 -> 64 | game vector 29 -> 64 -> 64, concat 192 ->
//                                                       256 -> LSTM(32) -> heading angle; (cos, sin, commanded speed) = the mlc target
//   networks/legged_robot/pmc_net/pmc_net.py:99-112      llc: the frozen primitive-level decoder
// The LSTM is `tpolicies`' layer-norm LSTM (absent from the reference tree), restated as in lifelike_agility_and_play_b200/policy_epmc.py,
// which is the host statement of the same nets and the checker of this kernel (tests/test_policy_epmc.py).
//
// One CTA (256 threads) per observation row; activations in shared memory, weights (1.2 MB, fp32) streamed from L2 with every
// thread of a layer reading consecutive columns; 0.23 M MAC per row on the CUDA cores.  The recurrent states live in device memory
// next to the engine's arrays and are wiped where the `done` flag of the previous step is set.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/llq.h"
#include "../../include/llq_policy.h"

namespace {

thread_local std::string g_err_h;
int fail_h(int code, const char* msg) { g_err_h = msg; return code; }

constexpr int kThreads = 256;
// roles of the weight arrays (index into the offset table the host builds from the model file)
enum Role {
  R_MEAN = 0, R_STD, R_MPROP_W, R_MPROP_B, R_MENC /* 28 arrays */, R_MEMB_W = R_MENC + 28, R_MEMB_B, R_MLSTM /* 9 */, R_LOGIT_W = R_MLSTM + 9, R_LOGIT_B,
  R_CODEBOOK, R_LLC /* 10 */, R_N_MLC = R_LLC + 10,
  R_HPROP_W = R_N_MLC, R_HPROP_B, R_HENC /* 26 */, R_HVEC = R_HENC + 26 /* 4 */, R_HEMB_W = R_HVEC + 4, R_HEMB_B, R_HLSTM /* 9 */, R_HMU_W = R_HLSTM + 9, R_HMU_B,
  R_N_ALL
};
static_assert(R_N_MLC == LLQ_HIER_ROLES_MLC && R_N_ALL == LLQ_HIER_ROLES_ALL, "role table (include/llq_policy.h)");

struct Net { const float* w; const int* off; };
__device__ __forceinline__ const float* arr(const Net& n, int role) { return n.w + n.off[role]; }

// out[j] = act(b[j] + sum_k in[k] W[k][j]); W row major [K][N]; the 256 threads split K into 256 / N parts (N <= 256)
__device__ void dense(const float* in, int K, const float* W, const float* b, int N, float* out, float* scratch, bool relu) {
  const int t = threadIdx.x;
  int parts = kThreads / N; if (parts < 1) parts = 1; if (parts > 8) parts = 8;
  const int j = t % N, p = t / N;
  if (p < parts) {
    float acc = 0.f;
    for (int k = p; k < K; k += parts) acc = fmaf(in[k], W[(size_t)k * N + j], acc);
    scratch[p * N + j] = acc;
  }
  __syncthreads();
  if (t < N) {
    float acc = b ? b[t] : 0.f;
    for (int q = 0; q < parts; q++) acc += scratch[q * N + t];
    out[t] = relu ? fmaxf(acc, 0.f) : acc;
  }
  __syncthreads();
}
// TF 'SAME' convolution + ReLU on a [H][W][C] tensor in shared memory (conv1d: H = 1, kh = 1); w [kh][kw][C][O]
__device__ void conv_same_relu(const float* in, int H, int W, int C, const float* w, const float* b, int kh, int kw, int O, int stride, float* out) {
  const int oh = (H + stride - 1) / stride, ow = (W + stride - 1) / stride;
  int th = (oh - 1) * stride + kh - H; if (th < 0) th = 0;
  int tw = (ow - 1) * stride + kw - W; if (tw < 0) tw = 0;
  const int pt = th / 2, pl = tw / 2;
  for (int idx = threadIdx.x; idx < oh * ow * O; idx += kThreads) {
    const int o = idx % O, x = (idx / O) % ow, y = idx / (O * ow);
    float acc = b[o];
    for (int di = 0; di < kh; di++) {
      const int yy = y * stride + di - pt;
      if (yy < 0 || yy >= H) continue;
      for (int dj = 0; dj < kw; dj++) {
        const int xx = x * stride + dj - pl;
        if (xx < 0 || xx >= W) continue;
        const float* ip = in + (yy * W + xx) * C;
        const float* wp = w + ((di * kw + dj) * C) * O + o;
        for (int c = 0; c < C; c++) acc = fmaf(ip[c], wp[c * O], acc);
      }
    }
    out[idx] = fmaxf(acc, 0.f);
  }
  __syncthreads();
}
// the three perception encoders of one usr_cmd_encoder: enc[0..8) 2-D map, [8..16) lidar, [16..24) front map; results appended to `cat`
__device__ void perception(const Net& n, int enc, const float* obs, float* bufa, float* bufb, float* cat) {
  for (int map = 0; map < 2; map++) {
    const float* src = obs + (map == 0 ? 135 : 588);
    const int e = enc + (map == 0 ? 0 : 16);
    for (int i = threadIdx.x; i < 325; i += kThreads) bufa[i] = src[i];
    __syncthreads();
    conv_same_relu(bufa, 25, 13, 1, arr(n, e), arr(n, e + 1), 1, 1, 4, 1, bufb);            // 25 x 13 x 4
    conv_same_relu(bufb, 25, 13, 4, arr(n, e + 2), arr(n, e + 3), 4, 4, 4, 2, bufa);        // 13 x 7 x 4
    conv_same_relu(bufa, 13, 7, 4, arr(n, e + 4), arr(n, e + 5), 2, 2, 4, 2, bufb);         // 7 x 4 x 4
    conv_same_relu(bufb, 7, 4, 4, arr(n, e + 6), arr(n, e + 7), 2, 2, 1, 1, cat + (map == 0 ? 0 : 60));   // 28
  }
  {   // lidar: periodic padding 4, conv(4) SAME, crop, two stride-2 convs, one to a single channel (epmc_net.py:97-117)
    const float* src = obs + 460;
    const int e = enc + 8;
    for (int i = threadIdx.x; i < 136; i += kThreads) bufa[i] = src[(i + 124) & 127];
    __syncthreads();
    conv_same_relu(bufa, 1, 136, 1, arr(n, e), arr(n, e + 1), 1, 4, 4, 1, bufb);             // 136 x 4
    conv_same_relu(bufb + 16, 1, 128, 4, arr(n, e + 2), arr(n, e + 3), 1, 4, 4, 2, bufa);    // crop 4 positions (x 4 channels) -> 64 x 4
    conv_same_relu(bufa, 1, 64, 4, arr(n, e + 4), arr(n, e + 5), 1, 4, 4, 2, bufb);          // 32 x 4
    conv_same_relu(bufb, 1, 32, 4, arr(n, e + 6), arr(n, e + 7), 1, 4, 1, 1, cat + 28);      // 32
  }
}
// layer norm over n (<= 128) values in shared memory: v <- (v - mean) / sqrt(var + 1e-12) * g + b (tf.contrib.layers.layer_norm)
__device__ void layer_norm(float* v, int n, const float* beta, const float* gamma, float* red) {
  const int t = threadIdx.x;
  if (t < 32) {
    float s = 0.f;
    for (int i = t; i < n; i += 32) s += v[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float m = s / (float)n;
    float q = 0.f;
    for (int i = t; i < n; i += 32) { const float d = v[i] - m; q = fmaf(d, d, q); }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (t == 0) { red[0] = m; red[1] = 1.0f / sqrtf(q / (float)n + 1e-12f); }
  }
  __syncthreads();
  if (t < n) v[t] = (v[t] - red[0]) * red[1] * gamma[t] + beta[t];
  __syncthreads();
}
// one step of the layer-norm LSTM (nh = 32): x [256] in shared memory, state [c(32), h(32)] in global memory (updated in place);
// arrays lstm + 0..8 = wx, wh, b, beta_x, gamma_x, beta_h, gamma_h, beta_c, gamma_c.  Leaves h in hout[32].
__device__ void lstm_step(const Net& n, int lstm, const float* x, float* state, bool wipe, float* zx, float* zh, float* cbuf, float* hout,
                          float* scratch, float* red) {
  const int t = threadIdx.x;
  if (t < 64) cbuf[t] = wipe ? 0.f : state[t];              // cbuf[0..32) = c, [32..64) = h
  __syncthreads();
  dense(x, 256, arr(n, lstm), nullptr, 128, zx, scratch, false);
  dense(cbuf + 32, 32, arr(n, lstm + 1), nullptr, 128, zh, scratch, false);
  layer_norm(zx, 128, arr(n, lstm + 3), arr(n, lstm + 4), red);
  layer_norm(zh, 128, arr(n, lstm + 5), arr(n, lstm + 6), red);
  if (t < 32) {
    const float* b = arr(n, lstm + 2);
    const float gi = zx[t] + zh[t] + b[t], gf = zx[32 + t] + zh[32 + t] + b[32 + t];
    const float go = zx[64 + t] + zh[64 + t] + b[64 + t], gu = zx[96 + t] + zh[96 + t] + b[96 + t];
    const float c = (1.0f / (1.0f + expf(-(gf + 1.0f)))) * cbuf[t] + (1.0f / (1.0f + expf(-gi))) * tanhf(gu);     // forget_bias 1
    cbuf[t] = c;
    zx[t] = c;                                               // layer norm of the new cell state (32 values)
    zh[t] = 1.0f / (1.0f + expf(-go));
  }
  __syncthreads();
  if (t < 32) state[t] = cbuf[t];
  layer_norm(zx, 32, arr(n, lstm + 7), arr(n, lstm + 8), red);
  if (t < 32) { const float h = zh[t] * tanhf(zx[t]); hout[t] = h; state[32 + t] = h; }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) hier_policy_kernel(Net net, int strategic, const float* __restrict__ obs, long long obs_ld, int n_rows,
                                                               const unsigned char* __restrict__ done, float* __restrict__ state, float* __restrict__ actions,
                                                               int* __restrict__ codes, float* __restrict__ heading) {
  __shared__ float s_obs[965];
  __shared__ float s_p[135];
  __shared__ float s_a[1312], s_b[1312];      // convolution ping-pong (25 x 13 x 4 = 1300)
  __shared__ float s_cat[256];
  __shared__ float s_x[256], s_y[256];
  __shared__ float s_zx[128], s_zh[128], s_c[64], s_h[32];
  __shared__ float s_scr[8 * 256];
  __shared__ float s_red[4];
  __shared__ int s_code;
  const int row = blockIdx.x, t = threadIdx.x;
  if (row >= n_rows) return;
  const int ow = strategic ? 965 : 916;
  for (int i = t; i < ow; i += kThreads) s_obs[i] = obs[(size_t)row * obs_ld + i];
  __syncthreads();
  if (t < 135) s_p[t] = fminf(fmaxf((s_obs[t] - arr(net, R_MEAN)[t]) / (arr(net, R_STD)[t] + 1e-8f), -5.f), 5.f);
  __syncthreads();
  const bool wipe = done != nullptr && done[row] != 0;
  const int ssz = strategic ? 128 : 64;
  float* st = state + (size_t)row * ssz;
  float tgt0, tgt1, tgt2;
  if (strategic) {
    // ---- heading controller
    dense(s_p, 135, arr(net, R_HPROP_W), arr(net, R_HPROP_B), 64, s_cat, s_scr, true);                    // cat[0..64)
    perception(net, R_HENC, s_obs, s_a, s_b, s_x);                                                          // x[0..88)
    dense(s_x, 88, arr(net, R_HENC + 24), arr(net, R_HENC + 25), 64, s_cat + 64, s_scr, true);             // cat[64..128)
    if (t < 29) s_x[t] = t < 5 ? s_obs[913 + t] : (t < 20 ? s_obs[918 + t - 5] : (t < 27 ? s_obs[948 + t - 20] : s_obs[962 + t - 27]));
    __syncthreads();
    dense(s_x, 29, arr(net, R_HVEC), arr(net, R_HVEC + 1), 64, s_y, s_scr, true);
    dense(s_y, 64, arr(net, R_HVEC + 2), arr(net, R_HVEC + 3), 64, s_cat + 128, s_scr, true);              // cat[128..192)
    dense(s_cat, 192, arr(net, R_HEMB_W), arr(net, R_HEMB_B), 256, s_x, s_scr, true);
    lstm_step(net, R_HLSTM, s_x, st, wipe, s_zx, s_zh, s_c, s_h, s_scr, s_red);
    if (t == 0) {
      float a = arr(net, R_HMU_B)[0];
      for (int k = 0; k < 32; k++) a = fmaf(s_h[k], arr(net, R_HMU_W)[k], a);
      a = fminf(fmaxf(a, -3.14159265358979f), 3.14159265358979f);
      s_red[2] = a;
      if (heading) heading[row] = a;
    }
    __syncthreads();
    tgt0 = cosf(s_red[2]); tgt1 = sinf(s_red[2]); tgt2 = s_obs[964];
    st += 64;
  } else {
    tgt0 = s_obs[913]; tgt1 = s_obs[914]; tgt2 = s_obs[915];
  }
  // ---- code controller (environmental level)
  dense(s_p, 135, arr(net, R_MPROP_W), arr(net, R_MPROP_B), 64, s_cat, s_scr, true);                       // cat[0..64)
  if (t < 3) s_y[t] = t == 0 ? tgt0 : (t == 1 ? tgt1 : tgt2);
  __syncthreads();
  dense(s_y, 3, arr(net, R_MENC + 24), arr(net, R_MENC + 25), 32, s_x, s_scr, true);                       // x[0..32) = target embedding
  perception(net, R_MENC, s_obs, s_a, s_b, s_x + 32);                                                       // x[32..120)
  dense(s_x, 120, arr(net, R_MENC + 26), arr(net, R_MENC + 27), 64, s_cat + 64, s_scr, true);              // cat[64..128)
  dense(s_cat, 128, arr(net, R_MEMB_W), arr(net, R_MEMB_B), 256, s_x, s_scr, true);
  lstm_step(net, R_MLSTM, s_x, st, wipe, s_zx, s_zh, s_c, s_h, s_scr, s_red);
  dense(s_h, 32, arr(net, R_LOGIT_W), arr(net, R_LOGIT_B), 256, s_y, s_scr, false);                        // logits
  if (t < 32) {                                             // argmax, first occurrence
    float best = -3.4e38f; int bi = 0;
    for (int i = t; i < 256; i += 32) if (s_y[i] > best) { best = s_y[i]; bi = i; }
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (t == 0) { s_code = bi; if (codes) codes[row] = bi; }
  }
  __syncthreads();
  // ---- frozen primitive-level decoder
  if (t < 32) s_y[t] = arr(net, R_CODEBOOK)[t * 256 + s_code];
  __syncthreads();
  dense(s_p, 135, arr(net, R_LLC), arr(net, R_LLC + 1), 64, s_cat, s_scr, true);
  dense(s_y, 32, arr(net, R_LLC + 2), arr(net, R_LLC + 3), 32, s_cat + 64, s_scr, true);
  dense(s_cat, 96, arr(net, R_LLC + 4), arr(net, R_LLC + 5), 256, s_x, s_scr, true);
  dense(s_x, 256, arr(net, R_LLC + 6), arr(net, R_LLC + 7), 256, s_y, s_scr, true);
  dense(s_y, 256, arr(net, R_LLC + 8), arr(net, R_LLC + 9), 12, s_x, s_scr, false);
  if (t < 12) actions[(size_t)row * 12 + t] = s_x[t];
}

}  // namespace

struct llq_hier_policy {
  int device = 0, strategic = 0;
  float* d_w = nullptr;
  int* d_off = nullptr;
};

extern "C" {

int llq_hier_policy_create(const float* weights, int64_t n_weights, const int32_t* offsets, int32_t n_roles, int32_t strategic, int32_t device,
                           llq_hier_policy_handle* out) {
  if (!weights || !offsets || !out) return fail_h(LLQ_EINVAL, "null argument");
  if (n_roles != (strategic ? LLQ_HIER_ROLES_ALL : LLQ_HIER_ROLES_MLC)) return fail_h(LLQ_EINVAL, "role table has the wrong length (include/llq_policy.h)");
  for (int i = 0; i < n_roles; i++) if (offsets[i] < 0 || offsets[i] >= n_weights) return fail_h(LLQ_EINVAL, "role offset outside the weight blob");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail_h(LLQ_ECUDA, "no CUDA device visible (no CPU fallback)");
  if (device < 0 || device >= ndev) return fail_h(LLQ_EINVAL, "device ordinal out of range");
  if (cudaSetDevice(device) != cudaSuccess) return fail_h(LLQ_ECUDA, "cudaSetDevice failed");
  llq_hier_policy* h = new (std::nothrow) llq_hier_policy();
  if (!h) return fail_h(LLQ_ENOMEM, "out of memory");
  h->device = device; h->strategic = strategic ? 1 : 0;
  std::vector<int> off(LLQ_HIER_ROLES_ALL, 0);
  for (int i = 0; i < n_roles; i++) off[i] = offsets[i];
  if (cudaMalloc(&h->d_w, sizeof(float) * (size_t)n_weights) != cudaSuccess || cudaMalloc(&h->d_off, sizeof(int) * off.size()) != cudaSuccess ||
      cudaMemcpy(h->d_w, weights, sizeof(float) * (size_t)n_weights, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->d_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(h->d_w); cudaFree(h->d_off); delete h;
    return fail_h(LLQ_ECUDA, "weight upload failed");
  }
  *out = h;
  return LLQ_OK;
}

int llq_hier_policy_destroy(llq_hier_policy_handle h) {
  if (!h) return LLQ_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_w); cudaFree(h->d_off);
  delete h;
  return LLQ_OK;
}

int llq_hier_policy_forward(llq_hier_policy_handle h, const float* d_obs, int64_t obs_ld, int32_t n, const uint8_t* d_done, float* d_state,
                            float* d_actions, int32_t* d_codes, float* d_heading, void* stream) {
  if (!h || !d_obs || !d_state || !d_actions) return fail_h(LLQ_EINVAL, "null argument");
  if (n <= 0 || obs_ld < (h->strategic ? 965 : 916)) return fail_h(LLQ_EINVAL, "bad row count or row stride");
  if (cudaSetDevice(h->device) != cudaSuccess) return fail_h(LLQ_ECUDA, "cudaSetDevice failed");
  Net net{h->d_w, h->d_off};
  hier_policy_kernel<<<n, kThreads, 0, (cudaStream_t)stream>>>(net, h->strategic, d_obs, obs_ld, n, d_done, d_state, d_actions, d_codes, d_heading);
  if (cudaGetLastError() != cudaSuccess) return fail_h(LLQ_ECUDA, "hier_policy_kernel launch failed");
  return LLQ_OK;
}

const char* llq_hier_policy_last_error(void) { return g_err_h.c_str(); }

}  // extern "C"
