/* copo_b200.h - C ABI of libcopo_b200.so (B200-native CoPO hot path).
 *
 * The reference (decisionforce/CoPO) is pure Python and has no FFI for this path; the interfaces each
 * entry point replaces are the Python call sites cited per function (paths relative to
 * copo_code/copo/torch_copo/).  Conventions: plain pointers and sizes only, every array pointer is a
 * DEVICE pointer owned by the caller unless the name ends in _host, all work is enqueued on the
 * caller's stream (a cudaStream_t passed as void*), nothing synchronises unless stated.  Every function
 * returns 0 (B2C_OK) or a negative b2c_status; the message is available from b2c_last_error()
 * (thread-local).  Handles are not thread-safe; distinct handles are independent.
 */
#ifndef COPO_B200_H
#define COPO_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef enum { B2C_OK = 0, B2C_ERR_ARG = -1, B2C_ERR_CUDA = -2, B2C_ERR_STATE = -3 } b2c_status;

const char* b2c_last_error(void);
int b2c_version(void);
/* 1 when a CUDA device of compute capability 10.x is current, else 0 (never falls back to the CPU). */
int b2c_device_ok(void);

/* ---------------------------------------------------------------------------------------------------
 * Batched environment: S scenes x A agent slots.  Replaces
 *   metadrive MultiAgent*Env.reset/step           (called at utils/env_wrappers.py:95, 309, 277)
 *   CCEnv._update_distance_map/_find_in_range     (utils/env_wrappers.py:125-158)
 *   LCFEnv.step reward bookkeeping and _add_lcf    (utils/env_wrappers.py:307-361, 393-418)
 *   LCFEnv.set_lcf_dist / set_force_lcf            (utils/env_wrappers.py:420-430)
 * ------------------------------------------------------------------------------------------------- */
typedef struct b2c_env b2c_env;

typedef struct {
    int32_t num_scenes;           /* S */
    int32_t num_slots;            /* A, <= 64 */
    int32_t num_agents;           /* live population per scene (<= A); 0 means A */
    int32_t delay_done;           /* steps a finished vehicle stays as an obstacle (MetaDrive MARL: 25) */
    int32_t horizon;              /* scene horizon in env steps (1000) */
    int32_t agent_horizon;        /* per-agent max_step (1000) */
    int32_t allow_respawn;
    int32_t auto_reset;           /* restart a scene in the step it reaches its horizon */
    int32_t append_lcf;           /* LCFEnv: append (lcf+1)/2 to the observation (enable_copo) */
    int32_t lcf_uniform;          /* lcf_dist == "uniform" */
    int32_t scene_offset;         /* global index of scene 0 (multi-GPU sharding keeps RNG streams distinct) */
    uint32_t seed;
    float neighbours_distance;    /* strict '<' radius, env_wrappers.py:133 */
    float mf_nei_distance;        /* mean-field radius, '<=' kept, algo_ccppo.py:283 */
    float lcf_mean, lcf_std;      /* current LCF distribution, env_wrappers.py:200-201 */
    float force_lcf;              /* -100 disables, env_wrappers.py:183 */
} b2c_env_config;

typedef struct {
    float* obs;                   /* [S][A][D]   required */
    float* reward;                /* [S][A]      required; native reward (return_native_reward=True) */
    uint8_t* flags;               /* [S][A]      required; B2C_FLAG_* bits */
    uint64_t* nei_mask;           /* [S][A]      bit j: slot j is a neighbour (info["neighbours"]) */
    uint64_t* mf_mask;            /* [S][A]      neighbours with distance <= mf_nei_distance */
    float* nei_reward;            /* [S][A]      info["nei_rewards"] */
    float* global_reward;         /* [S]         info["global_rewards"] */
    int8_t* nei_list;             /* [S][A][4]   nearest neighbours in (distance, slot) order, -1 = none */
    int32_t* agent_id;            /* [S][A]      running agent number ("agent{id}") */
    float* lcf;                   /* [S][A]      info["lcf"] in [-1, 1] */
    uint8_t* scene_done;          /* [S]         done["__all__"] */
    uint16_t* obs_split;          /* [S][A][W]   optional: obs as the [hi | lo] bf16 operand of b2c_tc_linear,
                                                  W = b2c_env_obs_split_width() (saves the b2c_tc_split_rows pass) */
} b2c_env_io;

enum {
    B2C_FLAG_VALID = 1,           /* the slot's agent acted in this step: reward / done are meaningful */
    B2C_FLAG_DONE = 2,
    B2C_FLAG_ARRIVE = 4,
    B2C_FLAG_CRASH = 8,
    B2C_FLAG_OUT = 16,
    B2C_FLAG_MAXSTEP = 32,
    B2C_FLAG_SPAWNED = 64,        /* a new agent entered this slot; obs is its first observation */
    B2C_FLAG_ALIVE = 128          /* the slot will act in the next step */
};

int b2c_env_create(const b2c_env_config* cfg, const uint32_t* map_blob_host, int map_words, b2c_env** out);
int b2c_env_destroy(b2c_env* env);
int b2c_env_reset(b2c_env* env, const b2c_env_io* out, int new_episode, void* stream);
int b2c_env_step(b2c_env* env, const float* actions /* [S][A][2] */, const b2c_env_io* out, void* stream);
/* Steps only scenes [first_scene, first_scene + num_scenes): `actions` and every buffer of `out` hold just those scenes
 * (the caller passes the sub-range of its arrays).  Scenes are independent, so stepping the ranges of a partition one
 * after the other equals b2c_env_step; a host-facing caller uses it to start copying a range's results while the next
 * range is still being computed. */
int b2c_env_step_scenes(b2c_env* env, const float* actions, const b2c_env_io* out, int first_scene, int num_scenes,
                        void* stream);
int b2c_env_set_lcf_dist(b2c_env* env, float mean, float std);
int b2c_env_set_force_lcf(b2c_env* env, float value);
int b2c_env_set_num_agents(b2c_env* env, int num_agents);   /* curriculum: ChangeNEnv, env_wrappers.py:450 */
/* two-kernel mode only: runs the lidar kernel again on the poses / pairs of the last step (idempotent; profiling) */
int b2c_env_relaunch_lidar(b2c_env* env, const b2c_env_io* out, void* stream);
int b2c_env_obs_dim(const b2c_env* env);
int b2c_env_obs_split_width(const b2c_env* env);
int b2c_env_kernels_per_step(const b2c_env* env);            /* 1: fused kernel, 2: state kernel + lidar kernel */
int b2c_env_shape_specialised(const b2c_env* env);           /* 1: (slots, obs width) has a compile-time specialised kernel */
int b2c_env_state_words(const b2c_env* env);                /* u32 words per scene tile */
int b2c_env_slots_padded(const b2c_env* env);
int b2c_env_get_state(b2c_env* env, uint32_t* dst_host, void* stream);       /* synchronises the stream */
int b2c_env_set_state(b2c_env* env, const uint32_t* src_host, void* stream); /* synchronises the stream */


/* ---------------------------------------------------------------------------------------------------
 * Policy / value networks: 3-layer tanh MLPs, weights in torch Linear layout W[out][in] (row-major), fp32.
 * Replaces SlimFC stacks of CCModel / CoPOModel (algo_ccppo.py:74-219, algo_copo.py:96-182) and their autograd
 * backward inside {IPPO,CCPPO,CoPO}Policy.loss (algo_ippo.py:79-172, algo_ccppo.py:376-472, algo_copo.py:311-424).
 * ------------------------------------------------------------------------------------------------- */
/* y[M][N] = act(x[M][K] W^T + b); act: 0 none, 1 tanh.  ldx / ldy are row strides in floats. */
int b2c_linear_forward(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int M, int K, int N,
                       int act, void* stream);
/* dx[M][K] = (dy[M][N] W) * (1 - h_prev^2)   (h_prev = tanh output that fed this layer; NULL: no factor) */
int b2c_linear_backward_input(const float* dy, int ldy, const float* W, const float* h_prev, int ldh, float* dx, int ldx,
                              int M, int K, int N, void* stream);
/* dW[N][K] += dy^T x, db[N] += column sums of dy (db may be NULL).  Accumulates: zero dW / db first. */
int b2c_linear_backward_weight(const float* dy, int ldy, const float* x, int ldx, float* dW, float* db, int M, int K,
                               int N, void* stream);
/* db[N] += column sums of dy[M][N] (bias gradient on its own) */
int b2c_colsum(const float* dy, int ldy, float* db, int M, int N, void* stream);
/* narrow output layers (logits: N = 4, value: N = 1), N <= 8 */
int b2c_head_forward(const float* h, int ldh, const float* W, const float* b, float* y, int ldy, int M, int K, int N,
                     void* stream);
/* dz[M][K] = (dy W) * (dtanh ? 1 - h^2 : 1), dW += dy^T h, db += colsum(dy); dz / dW / db may be NULL */
int b2c_head_backward(const float* dy, int ldy, const float* h, int ldh, const float* W, float* dz, int ldz, float* dW,
                      float* db, int M, int K, int N, int dtanh, void* stream);
/* same, and dz also as the [hi | lo] bf16 operand of the tensor-core kernels (dz_split [M][2*K], K % 64 == 0) */
int b2c_head_backward_split(const float* dy, int ldy, const float* h, int ldh, const float* W, float* dz, int ldz,
                            uint16_t* dz_split, float* dW, float* db, int M, int K, int N, int dtanh, void* stream);
/* TorchDiagGaussian (rllib): logits = (mean[2] | log_std[2]); action = mean + exp(log_std) * eps, logp(action).
 * eps_in NULL: standard normals from the counter-based generator keyed by (seed, step, row); eps_out optional. */
int b2c_gaussian_sample(const float* logits, const float* eps_in, float* actions, float* logp, float* eps_out, int M,
                        uint32_t seed, uint32_t step, int deterministic, void* stream);

/* Per-row forward + backward of the PPO objective (mode 0) or of mean(logp) (mode 1, the theta_old term of
 * CoPOPolicy.meta_update, algo_copo.py:265-272).  Gradients are those of
 *   mean(-surr + vf_loss_coeff * sum_h vf_loss_h - entropy_coeff * H) + kl_coeff * mean(KL(old || new))
 * with the clipped "old" value loss (algo_copo.py:360-367; or the plain clamped one, see plain_value_loss), including
 * autograd's tie rule for min / max.
 * stats[8] (+=): sum(-surr), sum vf_loss[0..2], sum entropy, sum KL, sum logp, rows. */
typedef struct {
    const float* logits;          /* [rows][4] current policy output */
    const float* actions;         /* [rows][2] */
    const float* old_logp;        /* [rows]    SampleBatch.ACTION_LOGP */
    const float* old_logits;      /* [rows][4] SampleBatch.ACTION_DIST_INPUTS (NULL allowed when kl_coeff == 0) */
    const float* adv;             /* [rows]    advantages / normalized_advantages / global_advantages */
    const float* v_cur[3];        /* value heads (native, neighbourhood, global): current prediction */
    const float* v_old[3];        /*   prediction stored at sampling time (VF_PREDS, NEI_VALUES, GLOBAL_VALUES) */
    const float* v_tgt[3];        /*   targets (VALUE_TARGETS, NEI_TARGET, GLOBAL_TARGET) */
    float* dlogits;               /* [rows][4] out */
    float* dv[3];                 /* [rows]    out */
    double* stats;                /* [8] accumulated, may be NULL */
    int32_t rows, n_heads, mode;
    float clip_param, vf_clip_param, vf_loss_coeff, entropy_coeff, kl_coeff;
    int32_t norm_rows;            /* rows the means are taken over; 0 = `rows`.  Data parallel: the GLOBAL minibatch row
                                     count, so that the all-reduced SUM of the ranks' gradients is the gradient of the
                                     whole-minibatch mean even when ranks hold different numbers of rows */
    int32_t plain_value_loss;     /* 1: old_value_loss=False, vf_loss = clamp((v - target)^2, 0, vf_clip_param)
                                     (algo_ippo.py:146-148, algo_copo.py:358-363) */
    const float* dyn_coeffs;      /* NULL, or device [2] = {kl_coeff, entropy_coeff} read by the kernel at run time instead
                                     of the two by-value fields (KLCoeffMixin.update_kl changes kl_coeff between
                                     iterations; a launch captured in a CUDA graph keeps working) */
} b2c_ppo_head_args;
int b2c_ppo_head(const b2c_ppo_head_args* args, void* stream);

/* LCF side of the meta-gradient (algo_copo.py:280-287, compute_coordinated algo_copo.py:155-161), with
 * phi = (lcf_mean + lcf_std * eps) * pi/2:  out3 += { sum(cos(phi) adv + sin(phi) nei),
 * sum d, sum d * eps },  d = (-sin(phi) adv + cos(phi) nei) * pi/2. */
int b2c_lcf_meta_terms(const float* adv, const float* nei_adv, const float* eps, int rows, float lcf_mean, float lcf_std,
                       double* out3, void* stream);
/* same, reading the model's raw lcf_parameters[2] on the DEVICE (lcf_mean = clamp(tanh(p0)), lcf_std = exp(clamp(p1)),
 * algo_copo.py:171-177): no host read of the current LCF between meta-update minibatches */
int b2c_lcf_meta_terms_params(const float* adv, const float* nei_adv, const float* eps, int rows,
                              const float* lcf_parameters, double* out3, void* stream);

/* The same sums plus out4[3] += sum(global_adv) (the `global_adv` statistic meta_update logs, algo_copo.py:304). */
int b2c_lcf_meta_sums(const float* adv, const float* nei_adv, const float* eps, const float* global_adv, int rows,
                      const float* lcf_parameters, double* out4, void* stream);
/* Everything CoPOPolicy.meta_update does after the two policy gradients, their dot product and the sums above
 * (algo_copo.py:264-309), in one launch: lcf_adv_loss = (mean coordinated adv - raw_mean) / raw_std, the gradient of
 * grad_value * lcf_adv_loss w.r.t. lcf_parameters through lcf_mean = clamp(tanh(p0)) and lcf_std = exp(clamp(p1))
 * (algo_copo.py:171-177; zero outside the clamps), torch.optim.Adam's step on the two parameters, and the 13 logged
 * statistics in the order of meta_update's return dict (new_policy_ego_loss, old_policy_logp_loss, lcf_lcf_adv_loss,
 * lcf_final_loss, grad_value, lcf, lcf_deg, lcf_param, coordinated_adv, global_adv, lcf_std, lcf_std_deg, lcf_std_param).
 * st_new / st_old: the 8 statistics b2c_ppo_head accumulated for the new policy (mode 0) / the old one (mode 1); sums:
 * b2c_lcf_meta_sums' out4; all of them over the GLOBAL minibatch of `rows` rows (all-reduce them first). */
typedef struct {
    const double* grad_value;     /* [1] <g_new, g_old> */
    const double* st_new;         /* [8] */
    const double* st_old;         /* [8] */
    const double* sums;           /* [4] */
    double rows, raw_mean, raw_std;
    float* lcf_parameters;        /* [2] in / out */
    float* exp_avg;               /* [2] Adam state */
    float* exp_avg_sq;            /* [2] */
    float* lcf_grad;              /* [2] out */
    double* stats;                /* [13] out */
    float lr, beta1, beta2, eps;
    int32_t step;                 /* Adam step count of this update (1-based) */
} b2c_lcf_meta_finish_args;
int b2c_lcf_meta_finish(const b2c_lcf_meta_finish_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Rollout bookkeeping over time-major [T][N] columns (N = scenes x slots); flags are the B2C_FLAG_* bytes the
 * env wrote for the step in which the row's action was applied.
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* flags;         /* [T][N] */
    const float* rewards[3];      /* native, neighbourhood (info["nei_rewards"]), global (info["global_rewards"]) */
    const float* values[3];       /* VF_PREDS, NEI_VALUES, GLOBAL_VALUES */
    float* advantages[3];         /* out */
    float* targets[3];            /* out */
    int32_t T, N, heads;          /* heads = 1 (IPPO / CCPPO) or 3 (CoPO) */
    int32_t global_reward_per_scene;  /* 0: rewards[2] is [T][N]; A > 0: it is [T][N / A], one value per scene */
    float gamma, lambda_;         /* the global head always uses gamma = 1 (algo_copo.py:498-500) */
    const float* bootstrap[3];    /* optional [N] per head: value of the observation AFTER row T-1 (stock rllib PPO,
                                     used by IPPO: last_r = V(NEXT_OBS), algo_ippo.py inherits PPOTorchPolicy's
                                     postprocessing); NULL = the CCPPO / CoPO rule below */
} b2c_gae_args;
/* compute_advantages x3 (rllib postprocessing; algo_copo.py:189-204, 473-502): per (scene, slot) column, consecutive
 * valid rows up to a done row are one trajectory; a trajectory cut by the fragment end bootstraps with the value
 * of its own last row (algo_ccppo.py:362-365).  Scan in float64, results float32. */
int b2c_gae3(const b2c_gae_args* args, void* stream);
/* algo_copo.py:539-551.  out5 += { sum x, sum x^2, rows, sum g, sum g^2 } over valid rows, x = cos(lcf pi/2) adv +
 * sin(lcf pi/2) nei_adv (nei_adv NULL: x = adv, the stock PPO standardisation), g = global_adv (may be NULL). */
int b2c_lcf_mix_stats(const uint8_t* flags, const float* adv, const float* nei_adv, const float* step_lcf,
                      const float* global_adv, size_t rows, double* out5, void* stream);
/* normalized_adv = (x - mean) / max(1e-4, std); global_adv standardised in place with (gmean, gstd) */
int b2c_lcf_mix_apply(const uint8_t* flags, const float* adv, const float* nei_adv, const float* step_lcf,
                      float* global_adv, float* normalized_adv, size_t rows, float mean, float std, float gmean,
                      float gstd, void* stream);
/* centralized critic observation (algo_ccppo.py:225-355): mode 0 none, 1 mean field (mf_mask), 2 concat (nei_list);
 * rows = T * N ordered [t][scene][slot]; a neighbour contributes when it has a row at the same t (its VALID flag). */
int b2c_cc_obs_fuse(const float* obs, const float* actions, const uint8_t* flags, const uint64_t* mf_mask,
                    const int8_t* nei_list, float* cobs, size_t rows, int slots, int obs_dim, int act_dim, int cobs_dim,
                    int mode, int counterfactual, void* stream);
/* The same with a second output for the mean-field mode on whole scenes (rows a multiple of slots): cobs_split [rows][2*kp]
 * bf16 = the row once more as the [hi | lo] tensor-core operand of the central value network's first layer (kp = cobs_dim
 * padded to 64; the bits b2c_tc_split_rows would produce from cobs), so that CCPPOPolicy.postprocess_trajectory's value
 * predictions (algo_ccppo.py:357-371) need no conversion pass over the fused observations.  NULL: as b2c_cc_obs_fuse. */
int b2c_cc_obs_fuse_split(const float* obs, const float* actions, const uint8_t* flags, const uint64_t* mf_mask,
                          const int8_t* nei_list, float* cobs, uint16_t* cobs_split, int kp, size_t rows, int slots,
                          int obs_dim, int act_dim, int cobs_dim, int mode, int counterfactual, void* stream);
/* dst[r][0..width) = src[idx[r]][0..width): minibatch assembly from a shuffled index list */
int b2c_gather_rows(const float* src, size_t ld_src, const int64_t* idx, float* dst, size_t ld_dst, size_t rows, int width,
                    void* stream);
/* dst[c][r] = src[idx[r]][c], dst [width][rows]: the scalar columns of a minibatch, each one a contiguous vector */
int b2c_gather_cols(const float* src, size_t ld_src, const int64_t* idx, float* dst, size_t rows, int width, void* stream);
/* torch.optim.Adam step (no weight decay); grad_scale multiplies the gradient first (1/world_size after an all-reduce) */
int b2c_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                  float beta2, float eps, int step, float grad_scale, void* stream);
/* *out += <a, b> in float64 (algo_copo.py:274-278) */
int b2c_dot(const float* a, const float* b, size_t n, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * tcgen05 tensor-core path for the 256-wide layers ("split bf16": fp32 operands split into bf16 hi + lo, the four
 * products hi*hi + lo*hi + hi*lo + lo*lo accumulated in fp32 in TMEM).  Same layers as b2c_linear_forward /
 * b2c_linear_backward_input.
 *   a_split  [M][2*Kp] bf16: hi in columns [0, Kp), lo in [Kp, 2Kp), Kp = b2c_tc_padded_k(K), zero padded
 *   w_prep   [256][2*Kp] bf16: [hi | lo] of W (forward) or of W^T (input gradient)
 * ------------------------------------------------------------------------------------------------- */
int b2c_tc_padded_k(int K);
int b2c_tc_split_rows(const float* x, int ldx, uint16_t* out, int M, int K, int Kp, void* stream);
/* transpose = 0: W[N][K] -> rows N, reduction K.  transpose = 1: rows K, reduction N (dx = dy W). */
int b2c_tc_prep_weight(const float* W, uint16_t* out, int N, int K, int Kp, int transpose, void* stream);
/* out = epilogue(a W'^T): + bias, tanh (act = 1), * (1 - dtanh_src^2); writes fp32 [M][ld_out] and / or the
 * [hi | lo] bf16 operand of the next layer [M][512].  Output width is 256. */
/* Same layer with the narrow output layer fused into the epilogue (the activated 256-wide result never leaves the
 * SM unless out_f32 is given): head.out[m][j] = head.bias[j] + sum_n y[m][n] head.weight[j][n], n in {1, 4}.
 * With head.actions set (n = 4) the epilogue also draws the action from TorchDiagGaussian(logits) with the
 * counter-based generator keyed by (seed, step, row) and writes its log-probability (rollout step, a7). */
typedef struct {
    const float* weight;          /* [n][256] */
    const float* bias;            /* [n] */
    float* out;                   /* [M][n] */
    int32_t n;
    float* actions;               /* [M][2] or NULL */
    float* logp;                  /* [M] or NULL */
    uint32_t seed, step;
} b2c_tc_head;
int b2c_tc_linear_head(const uint16_t* a_split, const uint16_t* w_prep, const float* bias, float* out_f32, int ld_out,
                       int M, int Kp, int act, const b2c_tc_head* head, void* stream);
/* The whole inference pass of a [K]-256-256-n tanh network in one kernel (n in {1, 4}; replaces CCModel / CoPOModel
 * forward without gradients, algo_ccppo.py:201-219, algo_copo.py:138-153): head.out = head(tanh(tanh(a W1^T + b1) W2^T + b2)),
 * plus the action sample when head.actions is set.  The hidden layers stay in tensor memory / shared memory; results are
 * bit-identical to b2c_tc_linear followed by b2c_tc_linear_head.  w1_prep [256][2*Kp1], w2_prep [256][512]. */
int b2c_tc_mlp2_head(const uint16_t* a_split, int Kp1, const uint16_t* w1_prep, const float* b1, const uint16_t* w2_prep,
                     const float* b2, const b2c_tc_head* head, int M, void* stream);
/* The same kernel as the learner's forward pass (CCModel / CoPOModel forward under `loss`, algo_copo.py:311-330,
 * algo_ccppo.py:376-400): besides head.out it leaves what the backward pass reads again - the first hidden layer as the
 * [hi | lo] operand h1_split [M][512] (layer 2's input; 1 - h1^2 in b2c_tc_linear_dgrad) and the second hidden layer
 * h2 [M][256] fp32 (b2c_head_backward_split) - written from the epilogues, so the hidden layers cross HBM once (out)
 * instead of three times.  Bit-identical to b2c_tc_linear(split out) + b2c_tc_linear_head(fp32 out) in h1_split and h2.
 * Both buffers 32-byte aligned. */
int b2c_tc_mlp2_train(const uint16_t* a_split, int Kp1, const uint16_t* w1_prep, const float* b1, const uint16_t* w2_prep,
                      const float* b2, const b2c_tc_head* head, uint16_t* h1_split, float* h2, int M, void* stream);
/* Diagnostics: CTA 0 of the following b2c_tc_mlp2_head launches stamps clock64() at 16 pipeline events per tile into
 * dev_buffer ([tiles of CTA 0][16] int64); NULL turns it off (tools/fused_trace.py prints the timeline). */
int b2c_tc_mlp2_set_trace(long long* dev_buffer);
/* Weight gradient of a 256-wide layer on the tensor cores: dW[256][K] += dz^T x from the [hi | lo] operands
 * (dz_split [M][512], x_split [M][2*Kp], Kp <= 256).  workspace: b2c_tc_wgrad_parts() * 256 * Kp floats; the
 * per-CTA partial sums are added in a fixed order (deterministic). */
int b2c_tc_wgrad_parts(void);
int b2c_tc_wgrad(const uint16_t* dz_split, const uint16_t* x_split, float* workspace, float* dW, int M, int K, int Kp,
                 void* stream);
int b2c_tc_linear(const uint16_t* a_split, const uint16_t* w_prep, const float* bias, const float* dtanh_src, int ld_src,
                  float* out_f32, int ld_out, uint16_t* out_split, int M, int Kp, int act, void* stream);
/* Learner backward without fp32 intermediates (replaces the autograd of SlimFC layers, algo_copo.py:311-424 loss.backward()):
 *   b2c_tc_split_rows_ones  as b2c_tc_split_rows, with 1.0 in the first padding column (K < Kp)
 *   b2c_tc_linear_dgrad     out = (dz W) * (1 - h^2), h = hi + lo read from the layer's own [hi | lo] operand h_split [M][512]
 *   b2c_tc_wgrad_bias       as b2c_tc_wgrad; db[256] += column K of dz^T x (x from b2c_tc_split_rows_ones)
 *   b2c_head_backward_tc    as b2c_head_backward_split; dz_colsum[K] += column sums of dz (bias gradient of the layer
 *                           below); dz may be NULL (only the tensor-core operand leaves) */
int b2c_tc_split_rows_ones(const float* x, int ldx, uint16_t* out, int M, int K, int Kp, int ones_col, void* stream);
int b2c_tc_linear_dgrad(const uint16_t* dz_split, const uint16_t* wT_prep, const uint16_t* h_split, float* out_f32,
                        int ld_out, uint16_t* out_split, int M, int Kp, void* stream);
int b2c_tc_wgrad_bias(const uint16_t* dz_split, const uint16_t* x_split, float* workspace, float* dW, float* db, int M, int K,
                      int Kp, void* stream);
int b2c_head_backward_tc(const float* dy, int ldy, const float* h, int ldh, const float* W, float* dz, int ldz,
                         uint16_t* dz_split, float* dW, float* db, float* dz_colsum, int M, int K, int N, int dtanh,
                         void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
