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
// One CTA (256 threads) per 8 observation rows; activations in shared memory (89 kB), weights (1.2 MB, fp32) streamed from L2 with
// every thread of a layer reading consecutive columns and using each weight for all 8 rows; 0.23 M MAC per row on the CUDA cores.
// The recurrent states live in device memory next to the engine's arrays and are wiped where the `done` flag of the previous
// step is set.
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
constexpr int kRows = 8;                 // observation rows per CTA: every weight is read once per CTA and used for 8 rows (one warp per row
                                         // in the row-wise stages: layer norms, gates, argmax)
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

// out[r][j] = act(b[j] + sum_k in[r][k] W[k][j]) for the CTA's kRows rows; W row major [K][N], N a multiple of 4, every array 16-byte
// aligned in the blob.  A thread owns FOUR consecutive columns of all rows (one 16-byte weight load and kRows shared-memory broadcasts
// per 4 x kRows FMAs); the input dimension is split into `parts` interleaved slices over the thread groups, partial sums meet in `scratch`
// (parts * kRows * N <= 4096 floats).
__device__ void dense(const float* in, int in_ld, int K, const float* W, const float* b, int N, float* out, int out_ld, float* scratch, bool relu) {
  const int t = threadIdx.x;
  const int quads = N >> 2;
  int parts = kThreads / quads; if (parts > 8) parts = 8; if (parts > 512 / N) parts = 512 / N; if (parts < 1) parts = 1;
  const int jq = t % quads, p = t / quads;
  if (p < parts) {
    float acc[kRows][4];
#pragma unroll
    for (int r = 0; r < kRows; r++) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
    const float4* Wq = reinterpret_cast<const float4*>(W) + jq;
    int k = p;
    for (; k + 3 * parts < K; k += 4 * parts) {             // four 16-byte weight loads in flight
      float4 w[4];
#pragma unroll
      for (int u = 0; u < 4; u++) w[u] = Wq[(size_t)(k + u * parts) * quads];
#pragma unroll
      for (int u = 0; u < 4; u++) {
#pragma unroll
        for (int r = 0; r < kRows; r++) {
          const float x = in[r * in_ld + k + u * parts];
          acc[r][0] = fmaf(x, w[u].x, acc[r][0]); acc[r][1] = fmaf(x, w[u].y, acc[r][1]);
          acc[r][2] = fmaf(x, w[u].z, acc[r][2]); acc[r][3] = fmaf(x, w[u].w, acc[r][3]);
        }
      }
    }
    for (; k < K; k += parts) {
      const float4 w = Wq[(size_t)k * quads];
#pragma unroll
      for (int r = 0; r < kRows; r++) {
        const float x = in[r * in_ld + k];
        acc[r][0] = fmaf(x, w.x, acc[r][0]); acc[r][1] = fmaf(x, w.y, acc[r][1]);
        acc[r][2] = fmaf(x, w.z, acc[r][2]); acc[r][3] = fmaf(x, w.w, acc[r][3]);
      }
    }
#pragma unroll
    for (int r = 0; r < kRows; r++)
      *reinterpret_cast<float4*>(scratch + (p * kRows + r) * N + 4 * jq) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
  __syncthreads();
  for (int idx = t; idx < kRows * N; idx += kThreads) {
    const int r = idx / N, jj = idx - r * N;
    float v = b ? b[jj] : 0.f;
    for (int q = 0; q < parts; q++) v += scratch[(q * kRows + r) * N + jj];
    out[r * out_ld + jj] = relu ? fmaxf(v, 0.f) : v;
  }
  __syncthreads();
}
// TF 'SAME' convolution + ReLU on the [H][W][C] tensors of the CTA's rows in shared memory (conv1d: H = 1, kh = 1); w [kh][kw][C][O]
__device__ void conv_same_relu(const float* in, int in_ld, int H, int W, int C, const float* w, const float* b, int kh, int kw, int O, int stride,
                               float* out, int out_ld) {
  const int oh = (H + stride - 1) / stride, ow = (W + stride - 1) / stride;
  int th = (oh - 1) * stride + kh - H; if (th < 0) th = 0;
  int tw = (ow - 1) * stride + kw - W; if (tw < 0) tw = 0;
  const int pt = th / 2, pl = tw / 2, per_row = oh * ow * O;
  for (int idx = threadIdx.x; idx < kRows * per_row; idx += kThreads) {
    const int r = idx / per_row, q = idx - r * per_row;
    const int o = q % O, x = (q / O) % ow, y = q / (O * ow);
    const float* irow = in + r * in_ld;
    float acc = b[o];
    for (int di = 0; di < kh; di++) {
      const int yy = y * stride + di - pt;
      if (yy < 0 || yy >= H) continue;
      for (int dj = 0; dj < kw; dj++) {
        const int xx = x * stride + dj - pl;
        if (xx < 0 || xx >= W) continue;
        const float* ip = irow + (yy * W + xx) * C;
        const float* wp = w + ((di * kw + dj) * C) * O + o;
        for (int c = 0; c < C; c++) acc = fmaf(ip[c], wp[c * O], acc);
      }
    }
    out[r * out_ld + q] = fmaxf(acc, 0.f);
  }
  __syncthreads();
}
// first two layers of percep_2d_encoder in one pass over the raw 25 x 13 maps of all rows: 1 x 1 conv to 4 channels (+ ReLU), then 4 x 4
// stride 2 SAME (padding 1 before, 2 after; the padding belongs to the SECOND layer: padded taps contribute 0) -> [13][7][4] per row
__device__ void conv2d_first_two(const float* in, int in_ld, const float* w1, const float* b1, const float* w2, const float* b2, float* out, int out_ld) {
  for (int idx = threadIdx.x; idx < kRows * 364; idx += kThreads) {
    const int r = idx / 364, q = idx - r * 364;
    const int o = q & 3, x = (q >> 2) % 7, y = q / 28;
    const float* m = in + r * in_ld;
    float acc = b2[o];
    for (int di = 0; di < 4; di++) {
      const int yy = 2 * y + di - 1;
      if (yy < 0 || yy >= 25) continue;
      for (int dj = 0; dj < 4; dj++) {
        const int xx = 2 * x + dj - 1;
        if (xx < 0 || xx >= 13) continue;
        const float v = m[yy * 13 + xx];
        const float* wp = w2 + ((di * 4 + dj) * 4) * 4 + o;
#pragma unroll
        for (int c = 0; c < 4; c++) acc = fmaf(fmaxf(fmaf(v, w1[c], b1[c]), 0.f), wp[c * 4], acc);
      }
    }
    out[r * out_ld + q] = fmaxf(acc, 0.f);
  }
  __syncthreads();
}
// first two layers of percep_1d_encoder on the 128 lidar rays of all rows: periodic padding 4, conv1d(4 channels, k = 4, SAME), crop of the
// padded positions, conv1d(4, k = 4, stride 2, SAME: padding 1 before, 1 after) -> [64][4] per row (epmc_net.py:97-115)
__device__ void conv1d_first_two(const float* in, int in_ld, const float* w1, const float* b1, const float* w2, const float* b2, float* out, int out_ld) {
  for (int idx = threadIdx.x; idx < kRows * 256; idx += kThreads) {
    const int r = idx >> 8, q = idx & 255;
    const int o = q & 3, x = q >> 2;
    const float* ray = in + r * in_ld;
    float acc = b2[o];
    for (int dj = 0; dj < 4; dj++) {
      const int ii = 2 * x + dj - 1;                        // position in the cropped first-layer output
      if (ii < 0 || ii >= 128) continue;
      // first layer at padded position ii + 4: taps at padded positions ii + 3 .. ii + 6, i.e. rays (ii - 1 .. ii + 2) mod 128
      const float p0 = ray[(ii + 127) & 127], p1 = ray[ii], p2 = ray[(ii + 1) & 127], p3 = ray[(ii + 2) & 127];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const float a = fmaxf(fmaf(p0, w1[c], fmaf(p1, w1[4 + c], fmaf(p2, w1[8 + c], fmaf(p3, w1[12 + c], b1[c])))), 0.f);
        acc = fmaf(a, w2[(dj * 4 + c) * 4 + o], acc);
      }
    }
    out[r * out_ld + q] = fmaxf(acc, 0.f);
  }
  __syncthreads();
}
// the three perception encoders of one usr_cmd_encoder for every row of the CTA: enc[0..8) 2-D map, [8..16) lidar, [16..24) front map;
// row r's 88 features go to cat + r * cat_ld.  bufa [kRows][364], bufb [kRows][128].
__device__ void perception(const Net& n, int enc, const float* obs, int obs_ld, float* bufa, float* bufb, float* cat, int cat_ld) {
  for (int map = 0; map < 2; map++) {
    const int e = enc + (map == 0 ? 0 : 16);
    conv2d_first_two(obs + (map == 0 ? 135 : 588), obs_ld, arr(n, e), arr(n, e + 1), arr(n, e + 2), arr(n, e + 3), bufa, 364);      // 13 x 7 x 4
    conv_same_relu(bufa, 364, 13, 7, 4, arr(n, e + 4), arr(n, e + 5), 2, 2, 4, 2, bufb, 128);                                       // 7 x 4 x 4
    conv_same_relu(bufb, 128, 7, 4, 4, arr(n, e + 6), arr(n, e + 7), 2, 2, 1, 1, cat + (map == 0 ? 0 : 60), cat_ld);                // 28
  }
  const int e = enc + 8;
  conv1d_first_two(obs + 460, obs_ld, arr(n, e), arr(n, e + 1), arr(n, e + 2), arr(n, e + 3), bufa, 364);                           // 64 x 4
  conv_same_relu(bufa, 364, 1, 64, 4, arr(n, e + 4), arr(n, e + 5), 1, 4, 4, 2, bufb, 128);                                         // 32 x 4
  conv_same_relu(bufb, 128, 1, 32, 4, arr(n, e + 6), arr(n, e + 7), 1, 4, 1, 1, cat + 28, cat_ld);                                  // 32
}
// layer norm over n values of every row (warp r normalises row r): v <- (v - mean) / sqrt(var + 1e-12) * g + b (tf.contrib.layers.layer_norm)
__device__ void layer_norm(float* v, int ld, int n, const float* beta, const float* gamma) {
  const int r = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (r < kRows) {
    float* x = v + r * ld;
    float s = 0.f;
    for (int i = l; i < n; i += 32) s += x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float m = s / (float)n;
    float q = 0.f;
    for (int i = l; i < n; i += 32) { const float d = x[i] - m; q = fmaf(d, d, q); }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float inv = 1.0f / sqrtf(q / (float)n + 1e-12f);
    for (int i = l; i < n; i += 32) x[i] = (x[i] - m) * inv * gamma[i] + beta[i];
  }
  __syncthreads();
}
// one step of the layer-norm LSTM (nh = 32) for every row: x [kRows][256] in shared memory, state [c(32), h(32)] per row in global memory
// (updated in place; rows with live[r] == 0 are not stored); arrays lstm + 0..8 = wx, wh, b, beta_x, gamma_x, beta_h, gamma_h, beta_c,
// gamma_c.  Leaves h in hout[kRows][32].
__device__ void lstm_step(const Net& n, int lstm, const float* x, float* state, int state_ld, const int* live, const int* wipe, float* zx, float* zh,
                          float* cbuf, float* hout, float* scratch) {
  const int t = threadIdx.x;
  for (int idx = t; idx < kRows * 64; idx += kThreads) {
    const int r = idx >> 6, i = idx & 63;
    cbuf[idx] = (live[r] && !wipe[r]) ? state[(size_t)r * state_ld + i] : 0.f;          // cbuf[r][0..32) = c, [32..64) = h
  }
  __syncthreads();
  dense(x, 256, 256, arr(n, lstm), nullptr, 128, zx, 128, scratch, false);
  dense(cbuf + 32, 64, 32, arr(n, lstm + 1), nullptr, 128, zh, 128, scratch, false);
  layer_norm(zx, 128, 128, arr(n, lstm + 3), arr(n, lstm + 4));
  layer_norm(zh, 128, 128, arr(n, lstm + 5), arr(n, lstm + 6));
  const int r = t >> 5, u = t & 31;                         // 8 rows x 32 units
  {
    const float* b = arr(n, lstm + 2);
    float* px = zx + r * 128; float* ph = zh + r * 128;
    const float gi = px[u] + ph[u] + b[u], gf = px[32 + u] + ph[32 + u] + b[32 + u];
    const float go = px[64 + u] + ph[64 + u] + b[64 + u], gu = px[96 + u] + ph[96 + u] + b[96 + u];
    const float c = (1.0f / (1.0f + expf(-(gf + 1.0f)))) * cbuf[r * 64 + u] + (1.0f / (1.0f + expf(-gi))) * tanhf(gu);      // forget_bias 1
    __syncwarp();
    cbuf[r * 64 + u] = c;
    px[u] = c;                                               // layer norm of the new cell state (32 values)
    ph[u] = 1.0f / (1.0f + expf(-go));
    if (live[r]) state[(size_t)r * state_ld + u] = c;
  }
  __syncthreads();
  layer_norm(zx, 128, 32, arr(n, lstm + 7), arr(n, lstm + 8));
  {
    const float h = zh[r * 128 + u] * tanhf(zx[r * 128 + u]);
    hout[r * 32 + u] = h;
    if (live[r]) state[(size_t)r * state_ld + 32 + u] = h;
  }
  __syncthreads();
}

constexpr int kObsLd = 968;      // shared-memory row stride of the observation block
struct alignas(16) Smem {
  float scr[4096];                 // partial sums of the split layers (parts * rows * N <= 4096); first member: 16-byte aligned
  float obs[kRows][kObsLd];
  float p[kRows][136];
  float cat[kRows][256], x[kRows][256], y[kRows][256];
  float zx[kRows][128], zh[kRows][128], c[kRows][64], h[kRows][32];
  float a[kRows * 364], b[kRows * 128];   // convolution buffers: [13][7][4] / [64][4] and [7][4][4] / [32][4] per row
  float ang[kRows];
  int code[kRows], live[kRows], wipe[kRows];
};

__global__ void __launch_bounds__(kThreads) hier_policy_kernel(Net net, int strategic, const float* __restrict__ obs, long long obs_ld, int n_rows,
                                                               const unsigned char* __restrict__ done, float* __restrict__ state, float* __restrict__ actions,
                                                               int* __restrict__ codes, float* __restrict__ heading) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int row0 = blockIdx.x * kRows, t = threadIdx.x;
  const int ow = strategic ? 965 : 916;
  if (t < kRows) {
    const int live = row0 + t < n_rows;
    S.live[t] = live;
    S.wipe[t] = live && done != nullptr && done[row0 + t] != 0;
  }
  for (int idx = t; idx < kRows * kObsLd; idx += kThreads) {
    const int r = idx / kObsLd, i = idx - r * kObsLd;
    S.obs[r][i] = (row0 + r < n_rows && i < ow) ? obs[(size_t)(row0 + r) * obs_ld + i] : 0.f;
  }
  __syncthreads();
  for (int idx = t; idx < kRows * 135; idx += kThreads) {
    const int r = idx / 135, i = idx - r * 135;
    S.p[r][i] = fminf(fmaxf((S.obs[r][i] - arr(net, R_MEAN)[i]) / (arr(net, R_STD)[i] + 1e-8f), -5.f), 5.f);
  }
  __syncthreads();
  const int ssz = strategic ? 128 : 64;
  float* st = state + (size_t)row0 * ssz;
  if (strategic) {
    // ---- heading controller
    dense(&S.p[0][0], 136, 135, arr(net, R_HPROP_W), arr(net, R_HPROP_B), 64, &S.cat[0][0], 256, S.scr, true);              // cat[0..64)
    perception(net, R_HENC, &S.obs[0][0], kObsLd, S.a, S.b, &S.x[0][0], 256);                                                // x[0..88)
    dense(&S.x[0][0], 256, 88, arr(net, R_HENC + 24), arr(net, R_HENC + 25), 64, &S.cat[0][64], 256, S.scr, true);          // cat[64..128)
    for (int idx = t; idx < kRows * 29; idx += kThreads) {
      const int r = idx / 29, i = idx - r * 29;
      S.x[r][i] = i < 5 ? S.obs[r][913 + i] : (i < 20 ? S.obs[r][918 + i - 5] : (i < 27 ? S.obs[r][948 + i - 20] : S.obs[r][962 + i - 27]));
    }
    __syncthreads();
    dense(&S.x[0][0], 256, 29, arr(net, R_HVEC), arr(net, R_HVEC + 1), 64, &S.y[0][0], 256, S.scr, true);
    dense(&S.y[0][0], 256, 64, arr(net, R_HVEC + 2), arr(net, R_HVEC + 3), 64, &S.cat[0][128], 256, S.scr, true);           // cat[128..192)
    dense(&S.cat[0][0], 256, 192, arr(net, R_HEMB_W), arr(net, R_HEMB_B), 256, &S.x[0][0], 256, S.scr, true);
    lstm_step(net, R_HLSTM, &S.x[0][0], st, ssz, S.live, S.wipe, &S.zx[0][0], &S.zh[0][0], &S.c[0][0], &S.h[0][0], S.scr);
    if (t < kRows) {
      float a = arr(net, R_HMU_B)[0];
      for (int k = 0; k < 32; k++) a = fmaf(S.h[t][k], arr(net, R_HMU_W)[k], a);
      a = fminf(fmaxf(a, -3.14159265358979f), 3.14159265358979f);
      S.ang[t] = a;
      if (heading && S.live[t]) heading[row0 + t] = a;
    }
    __syncthreads();
    st += 64;
  }
  // ---- code controller (environmental level)
  dense(&S.p[0][0], 136, 135, arr(net, R_MPROP_W), arr(net, R_MPROP_B), 64, &S.cat[0][0], 256, S.scr, true);                // cat[0..64)
  if (t < kRows * 3) {
    const int r = t / 3, i = t - 3 * r;
    S.y[r][i] = strategic ? (i == 0 ? cosf(S.ang[r]) : (i == 1 ? sinf(S.ang[r]) : S.obs[r][964])) : S.obs[r][913 + i];
  }
  __syncthreads();
  dense(&S.y[0][0], 256, 3, arr(net, R_MENC + 24), arr(net, R_MENC + 25), 32, &S.x[0][0], 256, S.scr, true);                // x[0..32) = target embedding
  perception(net, R_MENC, &S.obs[0][0], kObsLd, S.a, S.b, &S.x[0][32], 256);                                                // x[32..120)
  dense(&S.x[0][0], 256, 120, arr(net, R_MENC + 26), arr(net, R_MENC + 27), 64, &S.cat[0][64], 256, S.scr, true);           // cat[64..128)
  dense(&S.cat[0][0], 256, 128, arr(net, R_MEMB_W), arr(net, R_MEMB_B), 256, &S.x[0][0], 256, S.scr, true);
  lstm_step(net, R_MLSTM, &S.x[0][0], st, ssz, S.live, S.wipe, &S.zx[0][0], &S.zh[0][0], &S.c[0][0], &S.h[0][0], S.scr);
  dense(&S.h[0][0], 32, 32, arr(net, R_LOGIT_W), arr(net, R_LOGIT_B), 256, &S.y[0][0], 256, S.scr, false);                  // logits
  {                                                         // argmax per row (warp r), first occurrence
    const int r = t >> 5, l = t & 31;
    float best = -3.4e38f; int bi = 0;
    for (int i = l; i < 256; i += 32) if (S.y[r][i] > best) { best = S.y[r][i]; bi = i; }
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (l == 0) { S.code[r] = bi; if (codes && S.live[r]) codes[row0 + r] = bi; }
  }
  __syncthreads();
  // ---- frozen primitive-level decoder
  { const int r = t >> 5, u = t & 31; S.y[r][u] = arr(net, R_CODEBOOK)[u * 256 + S.code[r]]; }
  __syncthreads();
  dense(&S.p[0][0], 136, 135, arr(net, R_LLC), arr(net, R_LLC + 1), 64, &S.cat[0][0], 256, S.scr, true);
  dense(&S.y[0][0], 256, 32, arr(net, R_LLC + 2), arr(net, R_LLC + 3), 32, &S.cat[0][64], 256, S.scr, true);
  dense(&S.cat[0][0], 256, 96, arr(net, R_LLC + 4), arr(net, R_LLC + 5), 256, &S.x[0][0], 256, S.scr, true);
  dense(&S.x[0][0], 256, 256, arr(net, R_LLC + 6), arr(net, R_LLC + 7), 256, &S.y[0][0], 256, S.scr, true);
  dense(&S.y[0][0], 256, 256, arr(net, R_LLC + 8), arr(net, R_LLC + 9), 12, &S.x[0][0], 256, S.scr, false);
  if (t < kRows * 12) {
    const int r = t / 12, i = t - 12 * r;
    if (S.live[r]) actions[(size_t)(row0 + r) * 12 + i] = S.x[r][i];
  }
}

}  // namespace

struct llq_hier_policy {
  int device = 0, strategic = 0;
  bool attr_set = false;
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
  if (!h->attr_set) {
    if (cudaFuncSetAttribute(hier_policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)) != cudaSuccess)
      return fail_h(LLQ_ECUDA, "cudaFuncSetAttribute failed");
    h->attr_set = true;                                      // per handle = per device
  }
  hier_policy_kernel<<<(n + kRows - 1) / kRows, kThreads, sizeof(Smem), (cudaStream_t)stream>>>(net, h->strategic, d_obs, obs_ld, n, d_done, d_state,
                                                                                              d_actions, d_codes, d_heading);
  if (cudaGetLastError() != cudaSuccess) return fail_h(LLQ_ECUDA, "hier_policy_kernel launch failed");
  return LLQ_OK;
}

const char* llq_hier_policy_last_error(void) { return g_err_h.c_str(); }

}  // extern "C"
