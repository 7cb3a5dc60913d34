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
int b2c_env_set_lcf_dist(b2c_env* env, float mean, float std);
int b2c_env_set_force_lcf(b2c_env* env, float value);
int b2c_env_set_num_agents(b2c_env* env, int num_agents);   /* curriculum: ChangeNEnv, env_wrappers.py:450 */
int b2c_env_obs_dim(const b2c_env* env);
int b2c_env_state_words(const b2c_env* env);                /* u32 words per scene tile */
int b2c_env_slots_padded(const b2c_env* env);
int b2c_env_get_state(b2c_env* env, uint32_t* dst_host, void* stream);       /* synchronises the stream */
int b2c_env_set_state(b2c_env* env, const uint32_t* src_host, void* stream); /* synchronises the stream */

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
