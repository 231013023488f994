/* llq_policy.h -- C ABI of the on-device PMC policy forward.
 *
 * Replaces, for the batched actor loop, the TensorFlow graph of the reference's primitive-level policy evaluated once per
 * env step in `PGAgent.step(obs, argmax=True)` (test_scripts/primitive_level/test_primitive_level_env.py:75-88):
 *   networks/legged_robot/pmc_net/pmc_net.py:33-58   vq_encoder / decoder (fully connected, ReLU)
 *   networks/legged_robot/pmc_net/pmc_net.py:99-114  llc: prop 135 -> 64, z 32 -> 32, concat -> 256 -> 256 -> 12 (mean)
 *   networks/legged_robot/pmc_net/pmc_net.py:130-137 running-mean normalisation, clip to +-5 (networks/layers.py:55)
 *   networks/legged_robot/pmc_net/pmc_net.py:155-171 nearest code of the 32 x 256 codebook
 * The observation rows are read in place from the engine's device buffer (or a trajectory slab: `obs_ld` floats per row) and
 * the 12 actions per env are written to device memory that llq_step_ex(LLQ_IO_DEVICE) consumes: no host round trip.
 *
 * `weights`: the 28 arrays of a shipped *.model file, fp32, concatenated in their stored order:
 *   0-3   prop_mean[135] prop_std[135] future_mean[72] future_std[72]
 *   4-9   value head W1[207x256] b1[256] W2[256x256] b2[256] W3[256x1] b3[1]   (tanh; pmc_net.py:139-144)
 *   10-15 enc W1[207x256] b1[256] W2[256x256] b2[256] W3[256x32] b3[32]          16 codebook[32x256]
 *   17-20 prop_embed W[135x64] b[64], z_embed W[32x32] b[32]
 *   21-26 dec W1[96x256] b1[256] W2[256x256] b2[256] W3[256x12] b3[12]           27 logstd[12]
 * All matrices row major [in][out] as TensorFlow stores them. */
#ifndef LLQ_POLICY_H
#define LLQ_POLICY_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LLQ_POLICY_N_WEIGHTS 358647   /* total floats of the list above */

typedef struct llq_policy* llq_policy_handle;

/* Uploads the weights to `device`.  Returns 0 or a negative LLQ_E* code (llq.h); message via llq_policy_last_error(). */
int llq_policy_create(const float* weights, int64_t n_weights, int32_t device, llq_policy_handle* out);
int llq_policy_destroy(llq_policy_handle h);
/* actions[n,12] = mean action for obs[n, >=207] (device pointers; obs_ld = row stride in floats); `codes` (int32[n], device,
 * nullable) receives the selected codebook index.  Asynchronous on `stream` (a cudaStream_t, 0 = the default stream). */
int llq_policy_forward(llq_policy_handle h, const float* d_obs, int64_t obs_ld, int32_t n, float* d_actions, int32_t* d_codes, void* stream);
/* Full actor step for rollouts: as llq_policy_forward, plus
 *   d_values  (float[n], device, nullable): the value head's V(obs)  -> PMCInputs.V (pmc_net_data.py:7-16);
 *   d_neglogp (float[n], device, nullable): when given, d_actions receives a SAMPLE a = mean + exp(logstd) * eps of the diagonal
 *     Gaussian head (agent.step(argmax=False); pmc_net.py:107-113) and d_neglogp its -log p(a) = 0.5 sum eps^2 + sum logstd +
 *     6 log(2 pi) -> PMCInputs.neglogp; eps ~ N(0,1) from Philox4x32-10 keyed by `seed`, counter (row, draw, `counter`): pass a
 *     different `counter` every step.  NULL: d_actions is the mean (argmax=True). */
int llq_policy_forward_ex(llq_policy_handle h, const float* d_obs, int64_t obs_ld, int32_t n, float* d_actions, int32_t* d_codes,
                          float* d_values, float* d_neglogp, uint64_t seed, uint64_t counter, void* stream);
/* Record variant for the rollout worker (the kernels write the whole trajectory record): as llq_policy_forward_ex, but
 * value and -log p of row i go to d_values[i * out_ld] / d_neglogp[i * out_ld] (out_ld = the slab's row stride puts them straight into
 * the value / neglogp columns of a [N, 223] record row), and the Gaussian noise is keyed by the GLOBAL row row_gid0 + i so that
 * shards with equal seeds draw different noise (one actor process per shard in the reference draws from its own np.random). */
int llq_policy_forward_rec(llq_policy_handle h, const float* d_obs, int64_t obs_ld, int32_t n, float* d_actions, int32_t* d_codes,
                           float* d_values, float* d_neglogp, int64_t out_ld, uint64_t seed, uint64_t counter, int64_t row_gid0, void* stream);
const char* llq_policy_last_error(void);

/* ---- environmental- and strategic-level policies (csrc/llq_policy_hier.cu): conv encoders + layer-norm LSTMs + the frozen
 * primitive-level decoder, one CTA per observation row, fp32 on the CUDA cores.  Replaces `PGAgent.step(obs, argmax=True)` of
 * test_scripts/environmental_level/test_environmental_level_env.py:95-100 and test_scripts/strategic_level/test_strategic_level_env.py:96
 * (mean heading, argmax code, mean action) for a whole batch of envs; nets: networks/legged_robot/epmc_net/epmc_net.py:86-177,
 * networks/legged_robot/sepmc_net/sepmc_net.py:122-203, networks/legged_robot/pmc_net/pmc_net.py:99-112.
 * `weights`: all arrays of the shipped *.model file, fp32, concatenated; `offsets[role]`: start of the array that plays `role`
 * (the host-side table is lifelike_agility_and_play_b200/policy_epmc.py::hier_role_arrays):
 *   0 prop mean, 1 prop std | code controller: 2-3 prop embed W b, 4-31 usr_cmd_encoder (2-D map 8, lidar 8, front map 8, target fc 2,
 *   fusion fc 2), 32-33 embed, 34-42 LSTM (wx wh b beta_x gamma_x beta_h gamma_h beta_c gamma_c), 43-44 logits, 45 codebook,
 *   46-55 low-level controller | heading controller (strategic level only): 56-57 prop embed, 58-83 perception encoder (24 + fusion fc 2),
 *   84-87 game-vector fc x 2, 88-89 embed, 90-98 LSTM, 99-100 heading fc. */
#define LLQ_HIER_ROLES_MLC 56
#define LLQ_HIER_ROLES_ALL 101
typedef struct llq_hier_policy* llq_hier_policy_handle;
int llq_hier_policy_create(const float* weights, int64_t n_weights, const int32_t* offsets, int32_t n_roles, int32_t strategic, int32_t device,
                           llq_hier_policy_handle* out);
int llq_hier_policy_destroy(llq_hier_policy_handle h);
/* d_actions[n,12] = mean action for d_obs[n, >= 916 (environmental) / 965 (strategic)] (device pointers, obs_ld = row stride in floats).
 * d_state [n, 64 / 128] floats: the LSTM states ([c, h] of the heading LSTM first at the strategic level), updated in place; rows whose
 * d_done[i] != 0 (uint8, nullable: the done flags of the step that produced these observations) start from a zero state.
 * d_codes (int32[n]) / d_heading (float[n], strategic level) are optional outputs.  Asynchronous on `stream`. */
int llq_hier_policy_forward(llq_hier_policy_handle h, const float* d_obs, int64_t obs_ld, int32_t n, const uint8_t* d_done, float* d_state,
                            float* d_actions, int32_t* d_codes, float* d_heading, void* stream);
const char* llq_hier_policy_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
