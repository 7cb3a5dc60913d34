// env_step.cu - fused scene step kernel (K1 step_state + K2 raycast_obs + K3 neighbour of SURVEY.md 2.2).
//
// One CTA works on one scene at a time (persistent loop over scenes).  The map blob and the scene's
// agent-state tile ([16 fields][AP slots] + 8-word header) are staged into shared memory with 1-D TMA
// bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP); every phase of sim_core.cuh then runs on shared
// memory, the assembled observation tile [A][D] and the updated state tile leave through bulk stores.
// Algorithmic HBM bytes per agent-step: 8 (action) + 64 (state in) + 64 (state out) + 4*D (obs) + 4 + 4 + 1
// + 8 = 4*D + 153.
//
// Replaces: MetaDrive env.step (reference call site env_wrappers.py:95), CCEnv._update_distance_map /
// _find_in_range (env_wrappers.py:125-158), LCFEnv.step reward bookkeeping (env_wrappers.py:313-357) and
// LCFEnv._add_lcf (env_wrappers.py:393-418).  Build with -fmad=false (bit-exact spec, oracle/sim.py).
#include <cuda_runtime.h>
#include <stdint.h>
#include "sim_core.cuh"
#include "b2c_internal.h"

namespace b2c {

struct EnvIO {
    const uint32_t* map;
    uint32_t* state;
    const float* actions;
    float* obs;
    float* reward;
    uint8_t* flags;
    unsigned long long* nei_mask;
    unsigned long long* mf_mask;
    float* nei_reward;
    float* global_reward;
    int8_t* nei_list;
    int32_t* agent_id;
    float* lcf;
    uint8_t* scene_done;
    int map_words;
    int tile_words;
    int obs_bulk;   // 1 when the obs tile can leave through a bulk store (16-byte multiple + aligned base)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

static constexpr int ENV_THREADS = 256;

__global__ void __launch_bounds__(ENV_THREADS)
env_step_kernel(const __grid_constant__ EnvConfig cfg, const __grid_constant__ EnvIO io) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int A = cfg.A, AP = cfg.AP, D = cfg.D;
    const int tid = threadIdx.x;
    // ---- carve shared memory ------------------------------------------------------------------------
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* s_map = reinterpret_cast<uint32_t*>(smem_raw + 16);
    uint32_t* s_st = s_map + io.map_words;
    float* s_obs = reinterpret_cast<float*>(s_st + io.tile_words);
    float* s_f = s_obs + ((A * D + 3) & ~3);
    int* s_i = reinterpret_cast<int*>(s_f + 6 * A);
    int* s_place = s_i + 5 * A;
    uint8_t* s_cand = reinterpret_cast<uint8_t*>(s_place + MAX_SPAWN);
    __shared__ int s_scene_done;

    SceneView v;
    v.map = s_map; v.st = s_st; v.obs = s_obs;
    v.cs = s_f; v.sn = s_f + A; v.rew = s_f + 2 * A; v.long_last = s_f + 3 * A; v.loc_s = s_f + 4 * A;
    v.loc_l = s_f + 5 * A;
    v.flags = s_i; v.crash = s_i + A; v.acted = s_i + 2 * A; v.linger = s_i + 3 * A; v.ncand = s_i + 4 * A;
    v.cand = s_cand; v.place_free = s_place;
    v.A = A; v.AP = AP; v.D = D;

    const uint32_t tile_bytes = (uint32_t)io.tile_words * 4u;
    const uint32_t map_bytes = (uint32_t)io.map_words * 4u;
    uint32_t parity = 0;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    bool first = true;

    for (int scene = blockIdx.x; scene < cfg.S; scene += gridDim.x) {
        uint32_t* g_tile = io.state + (size_t)scene * io.tile_words;
        // ---- stage map (first iteration) + state tile ------------------------------------------------
        if (tid == 0) {
            mbar_expect_tx(bar, tile_bytes + (first ? map_bytes : 0u));
            if (first) bulk_g2s(s_map, io.map, map_bytes, bar);
            bulk_g2s(s_st, g_tile, tile_bytes, bar);
        }
        float act0 = 0.0f, act1 = 0.0f;
        if (tid < A && !cfg.do_reset) {
            float2 a = reinterpret_cast<const float2*>(io.actions)[(size_t)scene * A + tid];
            act0 = a.x; act1 = a.y;
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        first = false;

        if (cfg.do_reset) {
            if (tid < A) phase_reset_slot(v, cfg, tid);
            if (tid == 0) phase_reset_scene(v, cfg);
        } else if (tid == 0) {
            v.hdr(H_EP_STEP) += 1;
        }
        __syncthreads();
        if (tid < A) phase_dynamics(v, cfg, tid, act0, act1);
        __syncthreads();
        if (!cfg.do_reset) {
            for (int idx = tid; idx < A * A; idx += ENV_THREADS) {
                int i = idx / A, j = idx - i * A;
                if (i < j && phase_pair_crash(v, i, j)) { v.crash[i] = 1; v.crash[j] = 1; }
            }
        }
        __syncthreads();
        if (tid < A) phase_outcome(v, cfg, tid);
        __syncthreads();
        if (tid < (int)s_map[M_NSPAWN]) phase_place_free(v, cfg, tid);
        __syncthreads();
        if (tid == 0) s_scene_done = phase_respawn(v, cfg, scene);
        __syncthreads();
        if (tid < A) phase_pose_refresh(v, tid);
        __syncthreads();
        if (tid < A) {
            NeiOut n = phase_neighbours(v, cfg, tid);
            size_t g = (size_t)scene * A + tid;
            if (io.nei_mask) io.nei_mask[g] = n.nei_mask;
            if (io.mf_mask) io.mf_mask[g] = n.mf_mask;
            if (io.nei_reward) io.nei_reward[g] = n.nei_reward;
            if (io.nei_list) {
                uint32_t pk = (uint32_t)(uint8_t)n.list[0] | ((uint32_t)(uint8_t)n.list[1] << 8) |
                              ((uint32_t)(uint8_t)n.list[2] << 16) | ((uint32_t)(uint8_t)n.list[3] << 24);
                reinterpret_cast<uint32_t*>(io.nei_list)[g] = pk;
            }
            io.reward[g] = v.rew[tid];
            io.flags[g] = (uint8_t)v.flags[tid];
            if (io.agent_id) io.agent_id[g] = v.geti(F_ID, tid);
            if (io.lcf) io.lcf[g] = v.f(F_LCF, tid);
            phase_observe_ego(v, cfg, tid);
        } else if (tid == ENV_THREADS - 1) {
            float g = phase_global_reward(v);
            if (io.global_reward) io.global_reward[scene] = g;
            if (io.scene_done) io.scene_done[scene] = (uint8_t)s_scene_done;
        }
        __syncthreads();
        const int n_ray = (int)s_map[M_NRAY];
        for (int idx = tid; idx < A * n_ray; idx += ENV_THREADS) {
            int i = idx / n_ray, k = idx - i * n_ray;
            phase_lidar(v, i, k);
        }
        // persist linger counters were folded into the status word by phase_outcome / phase_respawn
        if (tid < A) v.seti(F_STATUS, tid, v.status(tid) | (v.linger[tid] << 8));
        fence_async_smem();
        __syncthreads();
        // ---- write back: state tile + obs tile through bulk stores ------------------------------------
        float* g_obs = io.obs + (size_t)scene * A * D;
        if (tid == 0) {
            bulk_s2g(g_tile, s_st, tile_bytes);
            if (io.obs_bulk) bulk_s2g(g_obs, s_obs, (uint32_t)(A * D * 4));
            bulk_commit();
        }
        if (!io.obs_bulk) {
            for (int idx = tid; idx < A * D; idx += ENV_THREADS) g_obs[idx] = s_obs[idx];
        }
        if (tid == 0) bulk_wait_read0();
        __syncthreads();
    }
}

size_t env_smem_bytes(int A, int D, int map_words, int tile_words) {
    size_t words = 4 + (size_t)map_words + tile_words + ((A * D + 3) & ~3) + 6 * A + 5 * A + MAX_SPAWN;
    return words * 4 + (size_t)A * A + 16;
}

}  // namespace b2c

// ---- host side: handle + C ABI (include/copo_b200.h) ---------------------------------------------------
using namespace b2c;

struct b2c_env {
    EnvConfig cfg;
    uint32_t* d_map;
    uint32_t* d_state;
    int map_words;
    int tile_words;
    int device;
    int num_sms;
    size_t smem;
};

extern "C" {

int b2c_env_create(const b2c_env_config* c, const uint32_t* map_blob, int map_words, b2c_env** out) {
    if (!c || !map_blob || !out) return b2c_set_error(B2C_ERR_ARG, "b2c_env_create: null argument");
    if (c->num_slots < 1 || c->num_slots > MAX_SLOTS) return b2c_set_error(B2C_ERR_ARG, "num_slots must be in [1, 64]");
    if (map_words < 16 || map_blob[0] != 0xB200C0F0u || (int)map_blob[M_TOTAL] != map_words || (map_words & 3))
        return b2c_set_error(B2C_ERR_ARG, "map blob header is not valid");
    if (c->lcf_std <= 0.0f) return b2c_set_error(B2C_ERR_ARG, "lcf_std must be > 0 (env_wrappers.py:425)");
    int base = (int)map_blob[M_BASE_OBS];
    int D = base + (c->append_lcf ? 1 : 0);
    if ((int)map_blob[M_NSPAWN] > MAX_SPAWN || (int)map_blob[M_NSPAWN] > ENV_THREADS)
        return b2c_set_error(B2C_ERR_ARG, "map has more than 64 spawn places");
    if (EGO_DIM + NAVI_DIM + (int)map_blob[M_NRAY] + (int)map_blob[M_NSIDE] != base)
        return b2c_set_error(B2C_ERR_ARG, "map obs layout does not add up");
    b2c_env* e = new b2c_env();
    EnvConfig& k = e->cfg;
    k.S = c->num_scenes; k.A = c->num_slots; k.AP = (c->num_slots + 3) & ~3; k.D = D;
    k.num_agents = c->num_agents > 0 ? c->num_agents : c->num_slots;
    if (k.num_agents > k.A) { delete e; return b2c_set_error(B2C_ERR_ARG, "num_agents > num_slots"); }
    k.delay_done = c->delay_done; k.horizon = c->horizon; k.agent_horizon = c->agent_horizon;
    k.allow_respawn = c->allow_respawn; k.auto_reset = c->auto_reset; k.append_lcf = c->append_lcf;
    k.lcf_uniform = c->lcf_uniform; k.do_reset = 0; k.new_episode = 0; k.scene_offset = c->scene_offset;
    k.seed = c->seed; k.nei_dist = c->neighbours_distance; k.mf_dist = c->mf_nei_distance;
    k.lcf_mean = c->lcf_mean; k.lcf_std = c->lcf_std; k.force_lcf = c->force_lcf;
    e->map_words = map_words;
    e->tile_words = NUM_FIELDS * k.AP + HEADER_WORDS;
    B2C_CUDA_OR(cudaGetDevice(&e->device), delete e);
    B2C_CUDA_OR(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, e->device), delete e);
    e->smem = env_smem_bytes(k.A, k.D, map_words, e->tile_words);
    B2C_CUDA_OR(cudaFuncSetAttribute(env_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem),
                delete e);
    B2C_CUDA_OR(cudaMalloc(&e->d_map, (size_t)map_words * 4), delete e);
    B2C_CUDA_OR(cudaMalloc(&e->d_state, (size_t)k.S * e->tile_words * 4), delete e);
    B2C_CUDA_OR(cudaMemcpy(e->d_map, map_blob, (size_t)map_words * 4, cudaMemcpyHostToDevice), delete e);
    B2C_CUDA_OR(cudaMemset(e->d_state, 0, (size_t)k.S * e->tile_words * 4), delete e);
    *out = e;
    return B2C_OK;
}

int b2c_env_destroy(b2c_env* e) {
    if (!e) return B2C_OK;
    cudaFree(e->d_map);
    cudaFree(e->d_state);
    delete e;
    return B2C_OK;
}

static int launch_env(b2c_env* e, const float* actions, const b2c_env_io* o, int do_reset, int new_episode,
                      void* stream) {
    if (!e || !o || !o->obs || !o->reward || !o->flags)
        return b2c_set_error(B2C_ERR_ARG, "b2c_env_step: obs, reward and flags outputs are required");
    if (!do_reset && !actions) return b2c_set_error(B2C_ERR_ARG, "b2c_env_step: actions is null");
    EnvConfig cfg = e->cfg;
    cfg.do_reset = do_reset; cfg.new_episode = new_episode;
    EnvIO io;
    io.map = e->d_map; io.state = e->d_state; io.actions = actions; io.obs = o->obs; io.reward = o->reward;
    io.flags = o->flags; io.nei_mask = (unsigned long long*)o->nei_mask; io.mf_mask = (unsigned long long*)o->mf_mask;
    io.nei_reward = o->nei_reward; io.global_reward = o->global_reward; io.nei_list = o->nei_list;
    io.agent_id = o->agent_id; io.lcf = o->lcf; io.scene_done = o->scene_done;
    io.map_words = e->map_words; io.tile_words = e->tile_words;
    io.obs_bulk = (((size_t)cfg.A * cfg.D * 4) % 16 == 0) && (((uintptr_t)o->obs) % 16 == 0);
    int ctas_per_sm = (int)(200 * 1024 / (e->smem + 1024));
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int grid = e->num_sms * ctas_per_sm;
    if (grid > cfg.S) grid = cfg.S;
    env_step_kernel<<<grid, ENV_THREADS, e->smem, (cudaStream_t)stream>>>(cfg, io);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_env_reset(b2c_env* e, const b2c_env_io* out, int new_episode, void* stream) {
    return launch_env(e, nullptr, out, 1, new_episode, stream);
}
int b2c_env_step(b2c_env* e, const float* actions, const b2c_env_io* out, void* stream) {
    return launch_env(e, actions, out, 0, 0, stream);
}
int b2c_env_set_lcf_dist(b2c_env* e, float mean, float std) {
    if (!e) return b2c_set_error(B2C_ERR_ARG, "null env");
    if (!(std > 0.0f)) return b2c_set_error(B2C_ERR_ARG, "set_lcf_dist: std must be > 0 (env_wrappers.py:425)");
    if (!(mean >= -1.0f && mean <= 1.0f))
        return b2c_set_error(B2C_ERR_ARG, "set_lcf_dist: mean must be in [-1, 1] (env_wrappers.py:426)");
    e->cfg.lcf_mean = mean; e->cfg.lcf_std = std;
    return B2C_OK;
}
int b2c_env_set_force_lcf(b2c_env* e, float v) {
    if (!e) return b2c_set_error(B2C_ERR_ARG, "null env");
    e->cfg.force_lcf = v;
    return B2C_OK;
}
int b2c_env_set_num_agents(b2c_env* e, int n) {
    if (!e || n < 1 || n > e->cfg.A) return b2c_set_error(B2C_ERR_ARG, "set_num_agents: out of range");
    e->cfg.num_agents = n;
    return B2C_OK;
}
int b2c_env_obs_dim(const b2c_env* e) { return e ? e->cfg.D : -1; }
int b2c_env_state_words(const b2c_env* e) { return e ? e->tile_words : -1; }
int b2c_env_slots_padded(const b2c_env* e) { return e ? e->cfg.AP : -1; }
int b2c_env_get_state(b2c_env* e, uint32_t* dst_host, void* stream) {
    if (!e || !dst_host) return b2c_set_error(B2C_ERR_ARG, "null argument");
    B2C_CUDA(cudaMemcpyAsync(dst_host, e->d_state, (size_t)e->cfg.S * e->tile_words * 4, cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    B2C_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return B2C_OK;
}
int b2c_env_set_state(b2c_env* e, const uint32_t* src_host, void* stream) {
    if (!e || !src_host) return b2c_set_error(B2C_ERR_ARG, "null argument");
    B2C_CUDA(cudaMemcpyAsync(e->d_state, src_host, (size_t)e->cfg.S * e->tile_words * 4, cudaMemcpyHostToDevice,
                             (cudaStream_t)stream));
    B2C_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return B2C_OK;
}

}  // extern "C"
