// env_step.cu - the scene step (K1 step_state + K2 raycast_obs + K3 neighbour of SURVEY.md 2.2) and its C ABI.
//
// Two launch modes (b2c_env_create picks one; B2C_ENV_SPLIT overrides):
//   * two kernels (scenes of 16 slots and more): env_step_kernel<true> runs the per-slot phases of sim_core.cuh for a
//     group of scenes per CTA (thread = slot), then env_lidar_kernel casts the lasers, one scene per CTA;
//   * one fused kernel (small scenes): env_step_kernel<false>, observation tile and lidar included.
// The map blob and the group's agent-state tiles ([16 fields][AP slots] + 8-word header per scene) are staged into
// shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP); every phase then runs on shared
// memory; the updated state tiles and the observation rows leave through bulk stores.
// Algorithmic HBM bytes per agent-step: 8 (action) + 64 (state in) + 64 (state out) + 4*D (obs) + 4 + 4 + 1
// + 8 = 4*D + 153.
//
// Replaces: MetaDrive env.step (reference call site env_wrappers.py:95), CCEnv._update_distance_map /
// _find_in_range (env_wrappers.py:125-158), LCFEnv.step reward bookkeeping (env_wrappers.py:313-357) and
// LCFEnv._add_lcf (env_wrappers.py:393-418).  Build with -fmad=false (bit-exact spec, oracle/sim.py).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "sim_core.cuh"
#include "b2c_internal.h"

namespace b2c {

#ifndef B2C_LIDAR_OWN
#define B2C_LIDAR_OWN 1
#endif
static constexpr int LIDAR_OWN = B2C_LIDAR_OWN;
static constexpr int LIDAR_RAYS = 72;      // laser count of every shipped map (the specialised kernels assume it)
static constexpr int SPREAD_TAB_WORDS = 160;       // per warp: float4[32] {ox, oy, cc, ss} + int[32] {(k - excl + 4096) << 8 | row}
static constexpr int SPREAD_BITS_WORDS = 72;       // per warp: 32 pairs x up to 72 lasers = 2304 start bits
static constexpr int SPREAD_MAX_RAYS = 72;         // more lasers than this per observer: everything goes through the own-lane loop

// Shared-memory plan of one CTA working on `G` scenes at a time (offsets in bytes, every region 16-byte aligned).
struct SmemPlan {
    int map, st, obs, f, i, need, masks, queue, geom, total;
};
__host__ __device__ inline SmemPlan smem_plan(int G, int A, int D, int map_words, int tile_words, int n_warps,
                                              bool split = false) {
    SmemPlan p;
    int o = 16;                                          // mbarrier
    p.map = o;   o += map_words * 4;
    p.st = o;    o += G * tile_words * 4;
    p.obs = o;   o += split ? 0 : ((G * A * D + 3) & ~3) * 4;      // two-kernel mode: observations go straight to HBM
    p.f = o;     o += ((G * 6 * A + 3) & ~3) * 4;
    p.i = o;     o += ((G * (4 * A + MAX_SPAWN) + 3) & ~3) * 4;
    p.need = o;  o += ((3 * G + 4 + 3) & ~3) * 4;        // need[G], scene_done[G], queue fill (1 or per scene)
    p.masks = o; o += G * 4 * 8;                         // per scene: four slot masks (SceneView::masks)
    p.queue = o; o += split ? 0 : ((G * A * A + 7) & ~7) * 2;    // two-kernel mode: the lidar kernel builds the pair list
    p.geom = o;  o += split ? 0 : n_warps * (SPREAD_TAB_WORDS + SPREAD_BITS_WORDS) * 4;     // lidar_spread's per-warp scratch
    p.total = o;
    return p;
}

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
    uint32_t* obs_split;   // [S][A][kp] bf16 pairs: the policy's tensor-core operand ([hi | lo], mlp_tc.cu), or null
    int kp;                // padded observation width (multiple of 64)
    // two-kernel mode: the state kernel hands these to the lidar kernel
    float* pose;           // [S][rec_words]: x[A], y[A], cos[A], sin[A], {0, part mask lo, hi, 0}, the slots' lidar
                           // broad-phase masks lo[A], hi[A] (boxes a laser of the slot can reach), then the non-laser
                           // observation columns [D - n_ray][A]
    int rec_words;
    int map_words;
    int tile_words;
    int obs_bulk;   // 1 when the obs tiles can leave through a bulk store (16-byte aligned base)
    int group;      // scenes one CTA works on at a time
    int rec_stride; // two-kernel mode: words between two non-laser columns of the record (odd, >= A)
    int* lidar_next; // two-kernel mode: the lidar kernel's scene-group counter, reset here for the launch that follows
    SmemPlan pl;    // shared-memory offsets, computed once on the host (read from the constant bank instead of being
                    // re-derived - the 48-register state kernel rematerialises them at every use)
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
// Programmatic dependent launch (two-kernel mode): a kernel launched with the programmatic-serialisation attribute may
// start while its predecessor in the stream is still running; everything it reads that the predecessor writes comes
// after pdl_wait() (which returns when the predecessor has completed and its writes are visible).  Without the
// attribute both are no-ops.  Rule kept by every kernel here: pdl_trigger() is only called when nothing the NEXT kernel
// reads before its own pdl_wait() can still be written by this kernel or by anything before it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static constexpr int ENV_MAX_THREADS = 256;
static constexpr int ENV_MAX_GROUP = 16;      // scene tag in a queue entry is 4 bits

// two fp32 values -> packed bf16 pairs: hi = bf16(x), lo = bf16(x - hi) (one packed conversion each; the same
// bits as tc::split_rows_kernel produces, tests/test_env_gpu.py::test_env_emits_the_policy_operand)
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
    hi = *reinterpret_cast<uint32_t*>(&h2);
    float r0 = v0 - __uint_as_float(hi << 16);
    float r1 = v1 - __uint_as_float(hi & 0xffff0000u);
    __nv_bfloat162 l2 = __floats2bfloat162_rn(r0, r1);
    lo = *reinterpret_cast<uint32_t*>(&l2);
}

// One warp holds up to 32 (observer, box) pairs, set up one per lane (PairGeom + the observer's tile row).
//   1. Every pair's own lane casts the pair's first LIDAR_OWN lasers: most pairs are far boxes with a window of one or
//      two lasers (Intersection, 40 agents: 36 % one, 28 % two; 3.7 on average), so this needs no distribution at all.
//   2. The remaining lasers (rest = cnt - LIDAR_OWN of the pairs that have more) are spread evenly over the lanes: laser
//      r of the warp (0 <= r < total) belongs to the pair whose [excl, excl + rest) holds r.  The pairs with rest > 0
//      are ranked (ballot); pair `rank` leaves {ox, oy, cc, ss} and {first laser - excl, tile row} in the warp's
//      scratch table and sets bit `excl` of the warp's start bitmap.  excl is strictly increasing over the ranks, so
//      the owner of laser r0 + lane is (start bits before the batch) + popc(start bits of the batch up to my lane) - 1:
//      one broadcast load of a bitmap word and two 16 / 4-byte loads per batch of 32 lasers, no search, no shuffles.
// A window that is empty still gets one laser (harmless: the slab test is exact).
__device__ __forceinline__ int laser_wrap(int k, int n_ray) {     // 0 <= k < 2 * n_ray
    const unsigned a = (unsigned)k, b = (unsigned)(k - n_ray);
    return (int)(a < b ? a : b);
}

__device__ __forceinline__ void lidar_spread(const PairGeom& g, int row, bool live, int n_ray, int D, const float2* ray,
                                             float* lasers, uint32_t* tab, uint32_t* bits) {
    const int lane = threadIdx.x & 31;
    const int cnt = live ? (g.cnt > 0 ? g.cnt : 1) : 0;
    const int own = (n_ray <= SPREAD_MAX_RAYS) ? LIDAR_OWN : n_ray;
    {
        float* dst = lasers + (size_t)row * D;
        for (int q = 0; q < own; ++q) {
            if (n_ray > SPREAD_MAX_RAYS && !__any_sync(0xffffffffu, q < cnt)) break;
            if (q < cnt) {
                const int k = laser_wrap(g.k0 + q, n_ray);
                const float2 rd = ray[k];
                lidar_ray(g.nx1, g.nx2, g.ny1, g.ny2, g.cc, g.ss, rd.x, rd.y, dst + k);
            }
        }
    }
    const int rest = cnt > own ? cnt - own : 0;
    const unsigned has = __ballot_sync(0xffffffffu, rest > 0);
    if (has == 0u) return;
    int incl = rest;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int nb = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += nb;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - rest;
    float4* tab4 = reinterpret_cast<float4*>(tab);
    int* tabi = reinterpret_cast<int*>(tab + 128);
    bits[lane] = 0u; bits[lane + 32] = 0u;
    if (lane < SPREAD_BITS_WORDS - 64) bits[lane + 64] = 0u;
    __syncwarp();
    if (rest > 0) {
        const int rank = __popc(has & ((1u << lane) - 1u));
        tab4[rank] = make_float4(g.ox, g.oy, g.cc, g.ss);
        tabi[rank] = ((g.k0 + own - excl + 4096) << 8) | row;
        atomicOr(bits + (excl >> 5), 1u << (excl & 31));
    }
    __syncwarp();
    const unsigned upto = 0xffffffffu >> (31 - lane);
    int before = 0;
    for (int r0 = 0; r0 < total; r0 += 32) {
        const unsigned starts = bits[r0 >> 5];
        const int owner = before + __popc(starts & upto) - 1;
        before += __popc(starts);
        if (r0 + lane < total) {
            const float4 t = tab4[owner];
            const int code = tabi[owner];
            const int k = laser_wrap((code >> 8) - 4096 + r0 + lane, n_ray);       // first laser of the rest + (r - excl)
            const float2 rd = ray[k];
            lidar_ray(-HALF_L - t.x, HALF_L - t.x, -HALF_W - t.y, HALF_W - t.y, t.z, t.w, rd.x, rd.y,
                      lasers + (size_t)(code & 255) * D + k);
        }
    }
    __syncwarp();                                        // the table and the bitmap are free for the warp's next 32 pairs
}

// SPLIT = false: the whole step in one kernel (observation tile in shared memory, lidar included).
// SPLIT = true: the state half of the two-kernel mode - per-slot phases only, small shared-memory footprint (no
// observation tile) so several times more scenes are resident per SM; ego / navigation features, poses and the
// slots' lidar broad-phase masks go to a scratch record for env_lidar_kernel.
// TA / TD: slots per scene and observation width known at compile time (0 = read from the config at run time).  The
// shapes of BASELINE.json's configurations get their own instantiation: slot loops unroll, the `tid / A` and
// field * AP + slot address arithmetic folds into immediates.
template <bool SPLIT, int TA, int TD>
__global__ void __launch_bounds__(SPLIT ? 128 : ENV_MAX_THREADS, SPLIT ? 10 : 2)      // state kernel: 48 registers, 10 CTAs per SM
env_step_kernel(const __grid_constant__ EnvConfig cfg, const __grid_constant__ EnvIO io) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int A = TA ? TA : cfg.A, AP = TA ? ((TA + 3) & ~3) : cfg.AP, D = TD ? TD : cfg.D, G = io.group;
    const int tid = threadIdx.x, NT = blockDim.x;
    const SmemPlan& pl = io.pl;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* s_map = reinterpret_cast<uint32_t*>(smem_raw + pl.map);
    uint32_t* s_st = reinterpret_cast<uint32_t*>(smem_raw + pl.st);
    float* s_obs = reinterpret_cast<float*>(smem_raw + pl.obs);
    float* s_f = reinterpret_cast<float*>(smem_raw + pl.f);
    int* s_i = reinterpret_cast<int*>(smem_raw + pl.i);
    int* s_need = reinterpret_cast<int*>(smem_raw + pl.need);
    int* s_done = s_need + G;
    int* s_nq = s_need + 2 * G;
    uint16_t* s_queue = reinterpret_cast<uint16_t*>(smem_raw + pl.queue);
    unsigned long long* s_masks = reinterpret_cast<unsigned long long*>(smem_raw + pl.masks);

    int scene_base = 0;                                  // first scene of the group being worked on
    auto view = [&](int sl) {
        SceneView v;
        v.map = s_map; v.st = s_st + sl * io.tile_words;
        v.obs = SPLIT ? io.pose + (size_t)(scene_base + sl) * io.rec_words + 6 * A + 4 : s_obs + (size_t)sl * A * D;
        v.obs_compact = SPLIT ? 1 : 0;
        v.obs_stride = io.rec_stride;
        float* f = s_f + sl * 6 * A;
        v.cs = f; v.sn = f + A; v.rew = f + 2 * A; v.long_last = f + 3 * A; v.loc_s = f + 4 * A; v.loc_l = f + 5 * A;
        int* q = s_i + sl * (4 * A + MAX_SPAWN);
        v.flags = q; v.crash = q + A; v.acted = q + 2 * A; v.linger = q + 3 * A; v.place_free = q + 4 * A;
        v.nqueue = s_nq; v.queue = s_queue;                  // fused mode only
        v.scene_local = SPLIT ? 0 : sl; v.masks = s_masks + 4 * sl;
        v.A = A; v.AP = AP; v.D = D;
        return v;
    };

    const uint32_t tile_bytes = (uint32_t)io.tile_words * 4u;
    const uint32_t map_bytes = (uint32_t)io.map_words * 4u;
    uint32_t parity = 0;
    if (tid == 0) mbar_init(bar, 1);
    for (int k = tid; k < 4 * G; k += NT) s_masks[k] = 0ull;
    // two-kernel mode: the lidar kernel may start its prologue (laser table, barriers) now - it reads the record only
    // after its own pdl_wait(), i.e. when this whole grid is done
    if constexpr (SPLIT) pdl_trigger();
    __syncthreads();
    bool first = true;
    const int sl_a = tid / A, ia = tid - sl_a * A;       // this thread's (scene in group, slot)
    const int n_warps = NT >> 5;
    const int n_groups = (cfg.S + G - 1) / G;

    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int scene0 = grp * G;
        scene_base = scene0;
        const int ng = (cfg.S - scene0 < G) ? cfg.S - scene0 : G;
        uint32_t* g_tiles = io.state + (size_t)scene0 * io.tile_words;
        const bool has_agent = tid < ng * A;
        // ---- stage map (first iteration) + the group's state tiles (1-D TMA bulk copies) -----------------
        if (tid == 0) {
            bulk_wait_read0();                           // the previous group's stores have left shared memory
            mbar_expect_tx(bar, (uint32_t)ng * tile_bytes + (first ? map_bytes : 0u));
            if (first) bulk_g2s(s_map, io.map, map_bytes, bar);
            bulk_g2s(s_st, g_tiles, (uint32_t)ng * tile_bytes, bar);
            *s_nq = 0;
        }
        // the map and the state tiles are already on their way; the actions (and every output buffer) belong to the
        // kernel before this one in the stream until pdl_wait() returns
        if constexpr (SPLIT) {
            if (first) {
                pdl_wait();
                if (blockIdx.x == 0 && tid == 0 && io.lidar_next) *io.lidar_next = 0;     // nobody reads it before this grid is done
            }
        }
        float act0 = 0.0f, act1 = 0.0f;
        if (has_agent && !cfg.do_reset) {
            float2 a = reinterpret_cast<const float2*>(io.actions)[(size_t)scene0 * A + tid];
            act0 = a.x; act1 = a.y;
        }
        if (tid == 0) mbar_wait(bar, parity);            // one thread polls; the others sleep at the barrier instead of
        __syncthreads();                                 // spending issue slots on try_wait loops
        parity ^= 1u;
        first = false;

        SceneView v = view(has_agent ? sl_a : 0);
        // ---- P1: header + dynamics + localisation (thread = slot) ----------------------------------------
        if (has_agent) {
            if (cfg.do_reset) phase_reset_slot(v, cfg, ia);
            if (ia == 0) {
                if (cfg.do_reset) phase_reset_scene(v, cfg); else v.hdr(H_EP_STEP) += 1;
                s_need[sl_a] = cfg.do_reset ? 1 : 0;
                v.masks[0] = 0ull; v.masks[1] = 0ull;
            }
            phase_dynamics(v, cfg, ia, act0, act1);
        }
        __syncthreads();
        // ---- P2: box overlap, balanced ring pairing (thread = slot) --------------------------------------
        if (has_agent && !cfg.do_reset) phase_crash_slot(v, ia);
        __syncthreads();
        // ---- P3: reward / termination (thread = slot) ----------------------------------------------------
        int my_need = 0;
        if (has_agent) {
            phase_outcome(v, cfg, ia);
            if (ia == 0) { v.masks[2] = 0ull; v.masks[3] = 0ull; }      // the overlap phase is done with them
            if (cfg.do_reset || v.status(ia) == ST_EMPTY || v.hdr(H_EP_STEP) >= cfg.horizon) { s_need[sl_a] = 1; my_need = 1; }
        }
        // the barrier doubles as a vote: most steps no scene of the group has anything to spawn or restart, and the two
        // spawn phases with their barriers are skipped
        const int any_need = __syncthreads_or(my_need);
        if (any_need) {
            // ---- P4: spawn-place occupancy, only for scenes that have something to spawn ------------------
            const int n_sp = (int)s_map[M_NSPAWN];
            for (int idx = tid; idx < ng * n_sp; idx += NT) {
                int sl = idx / n_sp, p = idx - sl * n_sp;
                if (s_need[sl]) phase_place_free(view(sl), cfg, p);
            }
            __syncthreads();
            // ---- P5: respawn, sequential per scene; scenes are spread over distinct warps -----------------
            {
                int w = tid >> 5, lane = tid & 31;
                int sl = lane * n_warps + w;
                if (sl < ng) {
                    int need = s_need[sl];
                    s_done[sl] = need ? phase_respawn(view(sl), cfg, scene0 + sl) : 0;
                }
            }
            __syncthreads();
        } else if (tid < ng) {
            s_done[tid] = 0;                                 // read after the barrier below
        }
        // ---- P6a: slot masks (participants / present vehicles) ---------------------------------------------
        if (has_agent) {
            phase_pose_refresh(v, ia);
            phase_masks(v, ia);
        }
        __syncthreads();
        // ---- P6b: neighbours (+ lidar pair queue), per-slot outputs, ego/navi features (thread = slot) ----
        const bool spare = (NT - ng * A) >= ng;          // idle threads take the per-scene reductions
        // maps with many side detectors (Tollgate: 65, a sincos and a division each): when the CTA has more threads than
        // slots - small batches run one scene per CTA, see b2c_env_create - the detectors of a slot are dealt to
        // NT / slots threads after the per-slot phases instead of being walked by the slot's own thread
        // (a shape-specialised kernel knows from its observation width whether its map can have that many: D = 19 + 72
        // lasers + side detectors (+ LCF); the others carry no code for it)
        constexpr bool MAY_SPREAD = SPLIT && (TD == 0 || TD - (EGO_DIM + NAVI_DIM + LIDAR_RAYS + 1) >= 16);
        const int n_side_map = MAY_SPREAD ? (int)s_map[M_NSIDE] : 0;
        const bool side_spread = MAY_SPREAD && n_side_map >= 16 && NT >= 2 * ng * A;
        if (has_agent) {
            NeiOut n = phase_neighbours(v, cfg, ia, io.mf_mask != nullptr, io.nei_list != nullptr);
            if constexpr (SPLIT) {
                uint32_t* pm = reinterpret_cast<uint32_t*>(io.pose + (size_t)(scene0 + sl_a) * io.rec_words + 4 * A + 4);
                pm[ia] = (uint32_t)n.cull_mask; pm[A + ia] = (uint32_t)(n.cull_mask >> 32);
            } else {
                queue_push_mask(v, ia, n.cull_mask);
            }
            size_t g = (size_t)scene0 * A + tid;
            if (io.nei_mask) io.nei_mask[g] = n.nei_mask;
            if (io.mf_mask) io.mf_mask[g] = n.mf_mask;
            if (io.nei_reward) io.nei_reward[g] = n.nei_reward;
            if (io.nei_list) {
                uint32_t pk = (uint32_t)(uint8_t)n.list[0] | ((uint32_t)(uint8_t)n.list[1] << 8) |
                              ((uint32_t)(uint8_t)n.list[2] << 16) | ((uint32_t)(uint8_t)n.list[3] << 24);
                reinterpret_cast<uint32_t*>(io.nei_list)[g] = pk;
            }
            io.reward[g] = v.rew[ia];
            io.flags[g] = (uint8_t)v.flags[ia];
            if (io.agent_id) io.agent_id[g] = v.geti(F_ID, ia);
            const float lcf_now = step_lcf(v, cfg, scene0 + sl_a, ia);
            if (io.lcf) io.lcf[g] = lcf_now;
            phase_observe_ego(v, cfg, ia, lcf_now, side_spread);
            if (!SPLIT) phase_lidar_init(v, ia);
        }
        {
            int sl = spare ? tid - ng * A : ((has_agent && ia == A - 1) ? sl_a : -1);
            if (sl >= 0 && sl < ng) {
                float gr = phase_global_reward(view(sl));
                if (io.global_reward) io.global_reward[scene0 + sl] = gr;
                if (io.scene_done) io.scene_done[scene0 + sl] = (uint8_t)s_done[sl];
            }
        }
        __syncthreads();
        if constexpr (!SPLIT) {
        // ---- P7: lidar.  Each warp takes 32 queued (observer, box) pairs, sets them up one per lane, then
        // spreads the pairs' lasers evenly over its lanes (lidar_spread) ---------------------------------------
        {
            const int nq = *s_nq;
            const int lane = tid & 31, warp = tid >> 5;
            const int n_ray = (int)s_map[M_NRAY];
            const float2* ray2 = reinterpret_cast<const float2*>(s_map + s_map[M_OFF_RAY]);
            uint32_t* s_spread = reinterpret_cast<uint32_t*>(smem_raw + pl.geom) + warp * (SPREAD_TAB_WORDS + SPREAD_BITS_WORDS);
            for (int base = warp * 32; base < nq; base += n_warps * 32) {
                const int e = base + lane;
                PairGeom g;
                g.nx1 = g.nx2 = g.ny1 = g.ny2 = g.cc = g.ss = g.ox = g.oy = 0.0f; g.k0 = 0; g.cnt = 0;
                int row = 0;
                if (e < nq) {
                    int code = s_queue[e];
                    int sl = code >> 12, oi = (code >> 6) & 63, oj = code & 63;
                    lidar_pair_setup(view(sl), oi, oj, g);
                    row = sl * A + oi;                           // < 256: one CTA's slots
                }
                lidar_spread(g, row, e < nq, n_ray, D, ray2, s_obs + EGO_DIM + NAVI_DIM, s_spread,
                             s_spread + SPREAD_TAB_WORDS);
            }
        }
        }
        if constexpr (MAY_SPREAD) {
            if (side_spread) {
                const int n_slots = ng * A, per = NT / n_slots;          // threads per slot
                const int slot = tid % n_slots, part = tid / n_slots;
                if (part < per) {
                    SceneView vs = view(slot / A);
                    for (int kk = part; kk < n_side_map; kk += per) phase_side_item(vs, slot - (slot / A) * A, kk);
                }
            }
        }
        fence_async_smem();
        __syncthreads();
        // ---- write back ------------------------------------------------------------------------------------
        if constexpr (SPLIT) {
            // state tiles through a bulk store; poses and the participant mask to the scratch buffer
            if (tid == 0) {
                bulk_s2g(g_tiles, s_st, (uint32_t)ng * tile_bytes);
                bulk_commit();
            }
            if (has_agent) {
                float* pr = io.pose + (size_t)(scene0 + sl_a) * io.rec_words;
                pr[ia] = v.f(F_X, ia); pr[A + ia] = v.f(F_Y, ia); pr[2 * A + ia] = v.cs[ia]; pr[3 * A + ia] = v.sn[ia];
                if (ia == 0) {
                    uint32_t* pi = reinterpret_cast<uint32_t*>(pr + 4 * A);
                    unsigned long long pm = v.masks[0];
                    pi[0] = 0u; pi[1] = (uint32_t)pm; pi[2] = (uint32_t)(pm >> 32); pi[3] = 0u;
                }
            }
        } else {
        float* g_obs = io.obs + (size_t)scene0 * A * D;
        const uint32_t obs_bytes = (uint32_t)(ng * A * D * 4);
        const bool obs_bulk = io.obs_bulk && (obs_bytes % 16u == 0u);
        if (tid == 0) {
            bulk_s2g(g_tiles, s_st, (uint32_t)ng * tile_bytes);
            if (obs_bulk) bulk_s2g(g_obs, s_obs, obs_bytes);
            bulk_commit();
        }
        if (!obs_bulk) {
            for (int idx = tid; idx < ng * A * D; idx += NT) g_obs[idx] = s_obs[idx];
        }
        if (io.obs_split) {
            // the same observations as the [hi | lo] bf16 operand of the policy's first layer (two columns per item)
            const int half = io.kp >> 1;
            uint32_t* g_sp = io.obs_split + (size_t)scene0 * A * io.kp;
            const int lane = tid & 31, warp = tid >> 5;
            for (int row = warp; row < ng * A; row += n_warps) {
                const float* orow = s_obs + (size_t)row * D;
                uint32_t* srow = g_sp + (size_t)row * io.kp;
                for (int c2 = lane; c2 < half; c2 += 32) {
                    int k = 2 * c2;
                    float v0 = (k < D) ? orow[k] : 0.0f;
                    float v1 = (k + 1 < D) ? orow[k + 1] : 0.0f;
                    __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
                    __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
                    __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
                    srow[c2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                    srow[half + c2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                }
            }
        }
    }
        }
    if (tid == 0) bulk_wait_read0();
}

// ---- lidar half of the two-kernel mode ---------------------------------------------------------------------------
// One CTA works on `G` scenes at a time and draws its next group from a global counter.  Everything a scene needs arrives
// in ONE bulk copy: the record the state kernel left (poses, participant mask, per-slot broad-phase masks, the non-laser
// observation columns).  All threads turn the masks into the (observer, box) pair list in shared memory (every warp takes
// the prefix sum of the mask populations, every thread expands one slice of one observer's mask) and lay out the
// observation tile: whole rows [A][D] - lasers preset to "nothing within range", non-laser columns transposed in from the
// record.  The pair pass lowers the lasers (same pair set-up / laser distribution as the fused kernel: lidar_spread); as
// soon as it is done the next record is fetched, the tile leaves as one bulk store, plus once more as the policy's bf16
// [hi | lo] operand through 16-byte coalesced stores.  Nothing on this path waits on a dependent global load.
struct LidarIO {
    const uint32_t* map;
    const float* pose;
    float* obs;
    uint32_t* obs_split;
    int* next_group;      // groups handed out beyond the first one of every CTA (0 at launch), or null: static striding
    int pair_stride, rec_words, rec_stride, kp, group, S, A, D, n_ray, ray_off;
};

struct LidarPlan {
    int ray, rec, pairs, nq, excl, tab, tile, total;
};
// lidar_spread's scratch: the start bitmap shares the per-warp rows of `excl` (the pair-list bases are dead once the
// list is built); the table lives in the record's non-laser columns of the group's first scene, which the tile
// layout has consumed before the pair pass starts - when they are large enough (and 16-byte aligned), else in a
// region of its own.
__host__ __device__ inline bool lidar_tab_in_record(int A, int D, int n_ray, int rec_stride, int n_warps) {
    return ((6 * A + 4) & 3) == 0 && (D - n_ray) * rec_stride >= n_warps * SPREAD_TAB_WORDS;
}
__host__ __device__ inline LidarPlan lidar_plan(int G, int A, int D, int n_ray, int rec_words, int rec_stride,
                                                int pair_stride, int n_warps) {
    LidarPlan p;
    int o = 16;                                           // mbarrier
    p.ray = o;   o += (n_ray * 8 + 15) & ~15;
    p.rec = o;   o += G * rec_words * 4;                  // rec_words is a multiple of 4
    p.pairs = o; o += G * pair_stride * 2;                // pair_stride is a multiple of 8
    p.nq = o;    o += ((G + 1 + 3) & ~3) * 4;             // pair counts of the group's scenes, then the CTA's next group
    p.excl = o;  o += n_warps * SPREAD_BITS_WORDS * 4;    // per warp: exclusive pair-list bases of the observers (MAX_SLOTS),
                                                          // then the start bitmap of lidar_spread
    p.tab = lidar_tab_in_record(A, D, n_ray, rec_stride, n_warps) ? p.rec + (6 * A + 4) * 4 : o;
    if (p.tab == o) o += n_warps * SPREAD_TAB_WORDS * 4;
    p.tile = o;  o += ((G * A * D + 3) & ~3) * 4;
    p.total = o;
    return p;
}
static_assert(SPREAD_BITS_WORDS >= MAX_SLOTS, "the start bitmap shares the rows of the pair-list bases");

template <int TA, int TD>
__global__ void __launch_bounds__(ENV_MAX_THREADS)
env_lidar_kernel(const __grid_constant__ LidarIO io) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int A = TA ? TA : io.A, D = TD ? TD : io.D, G = io.group;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, n_warps = NT >> 5;
    const int n_ray = TA ? LIDAR_RAYS : io.n_ray;
    const int rec = io.rec_words;
    const LidarPlan pl = lidar_plan(G, A, D, n_ray, rec, io.rec_stride, io.pair_stride, n_warps);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);          // record arrivals
    float2* s_ray = reinterpret_cast<float2*>(smem_raw + pl.ray);
    float* s_rec = reinterpret_cast<float*>(smem_raw + pl.rec);
    uint16_t* s_pairs = reinterpret_cast<uint16_t*>(smem_raw + pl.pairs);
    int* s_nq = reinterpret_cast<int*>(smem_raw + pl.nq);
    int* s_excl = reinterpret_cast<int*>(smem_raw + pl.excl);
    float* s_tile = reinterpret_cast<float*>(smem_raw + pl.tile);     // [G][A][D]
    const int n_groups = (io.S + G - 1) / G;
    const int lid0 = EGO_DIM + NAVI_DIM;
    auto fetch_record = [&](int grp) {                            // thread 0 only
        const int scene0 = grp * G;
        const int ng = (io.S - scene0 < G) ? io.S - scene0 : G;
        mbar_expect_tx(bar, (uint32_t)(ng * rec * 4));
        bulk_g2s(s_rec, io.pose + (size_t)scene0 * rec, (uint32_t)(ng * rec * 4), bar);
    };
    uint32_t parity = 0;
    if (tid == 0) mbar_init(bar, 1);
    {
        const float2* gr = reinterpret_cast<const float2*>(io.map + io.ray_off);     // the map is never rewritten
        for (int k = tid; k < n_ray; k += NT) s_ray[k] = gr[k];
    }
    pdl_wait();                                                   // the state kernel's records are complete and visible
    pdl_trigger();                                                // (so is everything before it: the next kernel may start)
    if (tid == 0 && (int)blockIdx.x < n_groups) fetch_record(blockIdx.x);
    __syncthreads();
    const bool rows_vec = ((D & 3) == 0);                         // rows start 16-byte aligned
    const int n_cell = D >> 2;                                    // float4 cells per row (rows_vec)
    const int cell_lo = (lid0 + 3) >> 2, cell_hi = (lid0 + n_ray) >> 2;      // cells [cell_lo, cell_hi) hold lasers only
    const int AS = io.rec_stride;                                 // words between two non-laser columns of the record
    int parts = NT / A;                                           // threads that share one observer's mask expansion
    parts = parts < 1 ? 1 : (parts > 4 ? 4 : parts);
    const int span = (A + parts - 1) / parts;                     // slots per thread
    // Scene groups are handed out dynamically: a CTA starts on group blockIdx.x and draws the next one from a global
    // counter when it is done with the record of the current one (scenes differ in their pair counts, and the scene
    // count is no multiple of the resident CTAs: static striding left 100 of 1332 CTAs a fourth scene at the C2 shape).
    for (int grp = blockIdx.x; grp < n_groups; grp = s_nq[G]) {
        const int scene0 = grp * G;
        const int ng = (io.S - scene0 < G) ? io.S - scene0 : G;
        if (tid == 0) mbar_wait(bar, parity);                     // one thread polls, the others sleep at the barrier
        __syncthreads();
        parity ^= 1u;
        for (int sl = 0; sl < ng; ++sl) {
            const float* ps = s_rec + sl * rec;
            const uint32_t* hd = reinterpret_cast<const uint32_t*>(ps + 4 * A);
            // ---- pair list: observer `o` owns popc(mask[o]) consecutive entries, ascending box order.  Every warp
            // takes the prefix sum of the populations (lane = observers `lane` and `lane + 32`, A <= 64) and keeps the
            // bases in its own shared-memory row; then every thread expands one slice of one observer's mask (an
            // observer's slots are dealt to `parts` threads), so the bit walks are short and run on all lanes --------
            {
                const uint32_t* cl = hd + 4;
                const uint32_t* ch = cl + A;
                const int o0 = lane, o1 = lane + 32;
                const int c0 = (o0 < A) ? __popc(cl[o0]) + __popc(ch[o0]) : 0;
                const int c1 = (o1 < A) ? __popc(cl[o1]) + __popc(ch[o1]) : 0;
                int i0 = c0, i1 = c1;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int n0 = __shfl_up_sync(0xffffffffu, i0, d), n1 = __shfl_up_sync(0xffffffffu, i1, d);
                    if (lane >= d) { i0 += n0; i1 += n1; }
                }
                const int t0 = __shfl_sync(0xffffffffu, i0, 31), t1 = __shfl_sync(0xffffffffu, i1, 31);
                int* my_excl = s_excl + warp * SPREAD_BITS_WORDS;
                my_excl[o0] = i0 - c0;                            // exclusive bases of observers lane / lane + 32
                my_excl[o1] = t0 + i1 - c1;
                __syncwarp();
                uint16_t* q = s_pairs + (size_t)sl * io.pair_stride;
                for (int item = tid; item < A * parts; item += NT) {
                    const int o = item / parts, b0 = (item - o * parts) * span;
                    const int b1 = (b0 + span < A) ? b0 + span : A;
                    const unsigned long long m = (unsigned long long)cl[o] | ((unsigned long long)ch[o] << 32);
                    const unsigned long long below = (1ull << b0) - 1ull;                     // b0 < A <= 64
                    const unsigned long long upto = (b1 >= 64) ? ~0ull : (1ull << b1) - 1ull;
                    int w = my_excl[o] + __popcll(m & below);
                    BitWalk walk(m & upto & ~below);
                    int j;
                    while (walk.next(j)) q[w++] = (uint16_t)((o << 6) | j);
                }
                if (tid == 0) s_nq[sl] = t0 + t1;
            }
            // ---- observation rows: lasers start at "nothing within range" for participants (0 for empty rows), the
            // other columns come out of the record (stored slot-fastest by the state kernel, odd column stride) -----
            {
                const unsigned long long pm = (unsigned long long)hd[1] | ((unsigned long long)hd[2] << 32);
                float* tile = s_tile + (size_t)sl * A * D;
                const float* eg = ps + 6 * A + 4;
                if (rows_vec) {
                    // whole 16-byte cells: the cells that hold lasers only ...
                    const int n_mid = cell_hi - cell_lo;
                    for (int e = tid; e < A * n_mid; e += NT) {
                        const int i = e / n_mid, q4 = cell_lo + (e - i * n_mid);
                        const float val = ((pm >> i) & 1ull) ? 1.0f : 0.0f;
                        reinterpret_cast<float4*>(tile + (size_t)i * D)[q4] = make_float4(val, val, val, val);
                    }
                    // ... and the cells with columns of the record in them
                    const int n_edge = n_cell - n_mid;
                    for (int e = tid; e < A * n_edge; e += NT) {
                        const int i = e / n_edge;
                        int q4 = e - i * n_edge;
                        q4 = (q4 < cell_lo) ? q4 : q4 - cell_lo + cell_hi;
                        const float val = ((pm >> i) & 1ull) ? 1.0f : 0.0f;
                        float x[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int col = 4 * q4 + u;
                            const int src = (col < lid0) ? col : col - n_ray;
                            x[u] = (col >= lid0 && col < lid0 + n_ray) ? val : eg[src * AS + i];
                        }
                        reinterpret_cast<float4*>(tile + (size_t)i * D)[q4] = make_float4(x[0], x[1], x[2], x[3]);
                    }
                } else {
                    for (int e = tid; e < A * D; e += NT) {
                        const int i = e / D, col = e - i * D;
                        const float val = ((pm >> i) & 1ull) ? 1.0f : 0.0f;
                        const int src = (col < lid0) ? col : col - n_ray;
                        tile[e] = (col >= lid0 && col < lid0 + n_ray) ? val : eg[src * AS + i];
                    }
                }
            }
        }
        __syncthreads();
        uint32_t* my_tab = reinterpret_cast<uint32_t*>(smem_raw + pl.tab) + warp * SPREAD_TAB_WORDS;
        uint32_t* my_bits = reinterpret_cast<uint32_t*>(s_excl + warp * SPREAD_BITS_WORDS);
        for (int sl = 0; sl < ng; ++sl) {
            const float* ps = s_rec + sl * rec;
            const int nq = s_nq[sl];
            const uint16_t* gq = s_pairs + (size_t)sl * io.pair_stride;
            float* tile = s_tile + (size_t)sl * A * D + lid0;
            for (int base = warp * 32; base < nq; base += n_warps * 32) {
                const int e = base + lane;
                PairGeom g;
                g.nx1 = g.nx2 = g.ny1 = g.ny2 = g.cc = g.ss = g.ox = g.oy = 0.0f; g.k0 = 0; g.cnt = 0;
                int oi = 0;
                if (e < nq) {
                    int code = gq[e];
                    oi = (code >> 6) & 63;
                    const int oj = code & 63;
                    lidar_pair_geom(ps[oi], ps[A + oi], ps[2 * A + oi], ps[3 * A + oi], ps[oj], ps[A + oj], ps[2 * A + oj],
                                    ps[3 * A + oj], n_ray, g);
                }
                lidar_spread(g, oi, e < nq, n_ray, D, s_ray, tile, my_tab, my_bits);
            }
        }
        fence_async_smem();
        __syncthreads();
        // the record is dead from here on: the next one travels while this group's rows are stored and converted (the
        // bulk copy only writes the record region, which nothing below reads)
        if (tid == 0) {
            const int nxt = io.next_group ? (int)gridDim.x + atomicAdd(io.next_group, 1) : grp + (int)gridDim.x;
            s_nq[G] = nxt;                                        // read by every thread after the barrier that ends the group
            if (nxt < n_groups) fetch_record(nxt);
        }
        // the group's observation rows are one contiguous block of HBM
        float* g_obs = io.obs + (size_t)scene0 * A * D;
        const int n_rows = ng * A;
        const uint32_t obs_bytes = (uint32_t)(n_rows * D * 4);
        const bool obs_bulk = ((obs_bytes | (uint32_t)(A * D * 4)) & 15u) == 0u &&
                              (reinterpret_cast<uintptr_t>(io.obs) & 15u) == 0u;
        if (obs_bulk) {
            if (tid == 0) {
                bulk_s2g(g_obs, s_tile, obs_bytes);
                bulk_commit();
            }
        } else {
            for (int idx = tid; idx < n_rows * D; idx += NT) g_obs[idx] = s_tile[idx];
        }
        if (io.obs_split) {
            // whole rows as the policy's [hi | lo] bf16 operand
            const int half = io.kp >> 1;                          // 32-bit words per half row (2 columns each)
            uint32_t* g_sp = io.obs_split + (size_t)scene0 * A * io.kp;
            if ((D & 3) == 0 && (half & 3) == 0 && half <= 64) {
                // eight columns per lane (two 16-byte shared loads, one 16-byte store into each half), two rows per warp
                // pass: lanes 0-15 take the even row, lanes 16-31 the odd one (kp <= 128: one pass covers a row)
                const int sub = lane & 15, rsel = lane >> 4;
                const int n_oct = half >> 2;                      // 16-byte cells per half row
                for (int row = 2 * warp + rsel; row < n_rows; row += 2 * n_warps) {
                    if (sub < n_oct) {
                        const float4* trow = reinterpret_cast<const float4*>(s_tile + (size_t)row * D);
                        uint4* srow = reinterpret_cast<uint4*>(g_sp + (size_t)row * io.kp);
                        const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        const float4 x0 = (8 * sub < D) ? trow[2 * sub] : z;
                        const float4 x1 = (8 * sub + 4 < D) ? trow[2 * sub + 1] : z;
                        uint4 hi, lo;
                        split_pair(x0.x, x0.y, hi.x, lo.x);
                        split_pair(x0.z, x0.w, hi.y, lo.y);
                        split_pair(x1.x, x1.y, hi.z, lo.z);
                        split_pair(x1.z, x1.w, hi.w, lo.w);
                        srow[sub] = hi;
                        srow[n_oct + sub] = lo;
                    }
                }
            } else if ((D & 3) == 0 && (half & 1) == 0) {
                // four columns per lane: one 16-byte shared load, one 8-byte store into each half
                for (int row = warp; row < n_rows; row += n_warps) {
                    const float4* trow = reinterpret_cast<const float4*>(s_tile + (size_t)row * D);
                    uint2* srow = reinterpret_cast<uint2*>(g_sp + (size_t)row * io.kp);
#pragma unroll 1
                    for (int q = lane; q < (half >> 1); q += 32) {
                        float4 x = (4 * q < D) ? trow[q] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        uint2 hi, lo;
                        split_pair(x.x, x.y, hi.x, lo.x);
                        split_pair(x.z, x.w, hi.y, lo.y);
                        srow[q] = hi;
                        srow[(half >> 1) + q] = lo;
                    }
                }
            } else {
                for (int row = warp; row < n_rows; row += n_warps) {
                    const float* trow = s_tile + (size_t)row * D;
                    uint32_t* srow = g_sp + (size_t)row * io.kp;
                    for (int c2 = lane; c2 < half; c2 += 32) {
                        const int k = 2 * c2;
                        const float v0 = (k < D) ? trow[k] : 0.0f;
                        const float v1 = (k + 1 < D) ? trow[k + 1] : 0.0f;
                        __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
                        __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
                        __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
                        srow[c2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                        srow[half + c2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0) bulk_wait_read0();                          // the tile has left shared memory
    }
}

}  // namespace b2c

// ---- host side: handle + C ABI (include/copo_b200.h) ---------------------------------------------------
using namespace b2c;

typedef void (*b2c_step_fn)(const EnvConfig, const EnvIO);
typedef void (*b2c_lidar_fn)(const LidarIO);

// Shape-specialised instantiations (slots x observation width) for the launch mode each shape runs by default:
// two-kernel mode from 16 slots up, the fused kernel below.  Everything else - other shapes, a mode forced through
// B2C_ENV_SPLIT, maps with another laser count, B2C_ENV_GENERIC=1 - runs the generic kernels (all sizes at run time).
#define B2C_SPLIT_SHAPES(X) X(40, 92) X(40, 91) X(40, 157) X(40, 156) X(30, 92) X(30, 91) X(20, 97) X(20, 96)
#define B2C_FUSED_SHAPES(X) X(10, 92) X(10, 91)
static b2c_step_fn pick_step_kernel(bool split, int A, int D, bool generic) {
    if (!generic) {
#define X(a, d) if (split && A == a && D == d) return env_step_kernel<true, a, d>;
        B2C_SPLIT_SHAPES(X)
#undef X
#define X(a, d) if (!split && A == a && D == d) return env_step_kernel<false, a, d>;
        B2C_FUSED_SHAPES(X)
#undef X
    }
    return split ? env_step_kernel<true, 0, 0> : env_step_kernel<false, 0, 0>;
}
static b2c_lidar_fn pick_lidar_kernel(int A, int D, bool generic) {
    if (!generic) {
#define X(a, d) if (A == a && D == d) return env_lidar_kernel<a, d>;
        B2C_SPLIT_SHAPES(X)
#undef X
    }
    return env_lidar_kernel<0, 0>;
}

struct b2c_env {
    EnvConfig cfg;
    b2c_step_fn step_fn;
    b2c_lidar_fn lidar_fn;
    int group;
    int threads;
    int split;               // 1: two-kernel mode (state kernel + lidar kernel)
    int specialised;         // 1: the shape has its own kernel instantiation
    int pdl;                 // two-kernel mode, programmatic dependent launch (B2C_ENV_PDL): bit 0 = the lidar kernel starts
                             // while the state kernel drains, bit 1 = the state kernel starts while its predecessor drains
    int lidar_group, lidar_threads, lidar_ctas, ray_off;
    size_t lidar_smem;
    float* d_pose;
    int* d_next;             // the lidar kernel's scene-group counter
    int pair_stride, n_ray, rec_words, rec_stride;
    uint32_t* d_map;
    uint32_t* d_state;
    int map_words;
    int tile_words;
    int device;
    int num_sms;
    size_t smem;
};

// Launch with (or without) the programmatic-serialisation attribute: the kernel may begin while the one before it in
// the stream drains (see pdl_wait / pdl_trigger).
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*fn)(KArgs...), int grid, int threads, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3((unsigned)threads); lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&lc, fn, args...);
}

extern "C" {

int b2c_env_create(const b2c_env_config* c, const uint32_t* map_blob, int map_words, b2c_env** out) {
    if (!c || !map_blob || !out) return b2c_set_error(B2C_ERR_ARG, "b2c_env_create: null argument");
    if (c->num_slots < 1 || c->num_slots > MAX_SLOTS) return b2c_set_error(B2C_ERR_ARG, "num_slots must be in [1, 64]");
    if (map_words < 16 || map_blob[0] != 0xB200C0F0u || (int)map_blob[M_TOTAL] != map_words || (map_words & 3))
        return b2c_set_error(B2C_ERR_ARG, "map blob header is not valid");
    if (c->lcf_std <= 0.0f) return b2c_set_error(B2C_ERR_ARG, "lcf_std must be > 0 (env_wrappers.py:425)");
    int base = (int)map_blob[M_BASE_OBS];
    int D = base + (c->append_lcf ? 1 : 0);
    if ((int)map_blob[M_NSPAWN] > MAX_SPAWN)
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
    // Launch shapes (measured on B200, profiles/r01_sweep.txt: small CTAs hide the phase barriers best).
    //   fused mode: one thread per slot in the per-slot phases, as many scenes per CTA as keep four CTAs per SM
    //   two-kernel mode (default from 16 slots up; B2C_ENV_SPLIT=0/1 overrides): the state kernel has no observation
    //   tile, so it packs its 128 threads with slots (3 scenes of 40) at 7 CTAs per SM; the lidar kernel takes one
    //   scene per CTA at 8+ CTAs per SM (profiles/r01_e_split_sweep.txt)
    e->split = (c->num_slots >= 16) ? 1 : 0;       // small scenes (parking lot, 10 slots) are faster fused
    if (const char* sp = getenv("B2C_ENV_SPLIT")) e->split = atoi(sp) ? 1 : 0;
    e->threads = 128;
    if (const char* t = getenv("B2C_ENV_THREADS")) e->threads = atoi(t);
    if (e->threads < 32 || e->threads > (e->split ? 128 : ENV_MAX_THREADS) || (e->threads & 31)) e->threads = 128;
    while (e->threads < k.A) e->threads += 32;
    int fit = e->threads / k.A;
    if (fit > ENV_MAX_GROUP) fit = ENV_MAX_GROUP;
    if (fit > k.S) fit = k.S;
    if (fit < 1) { delete e; return b2c_set_error(B2C_ERR_ARG, "num_slots does not fit one CTA"); }
    e->group = fit;
    const int smem_cap = e->split ? 36 * 1024 : 56 * 1024;
    while (e->group > 1 && smem_plan(e->group, k.A, k.D, map_words, e->tile_words, e->threads / 32, e->split).total > smem_cap)
        e->group -= 1;
    // Small batches of a map with many side detectors (Tollgate at the C4 shape: 1024 scenes per GPU = 342 groups of 3, a
    // third of a wave, every CTA walking 65 detectors per slot on the slot's own thread): one scene per CTA instead, its
    // 128 threads then share each slot's detectors three ways (env_step_kernel, side_spread)
    if (e->split && (int)map_blob[M_NSIDE] >= 16 && e->group > 1 &&
        (k.S + e->group - 1) / e->group <= 3 * e->num_sms && e->threads >= 2 * k.A)
        e->group = 1;
    if (const char* g = getenv("B2C_ENV_GROUP")) {
        int gg = atoi(g);
        if (gg >= 1 && gg <= fit) e->group = gg;
    }
    e->smem = (size_t)smem_plan(e->group, k.A, k.D, map_words, e->tile_words, e->threads / 32, e->split).total;
    e->n_ray = (int)map_blob[M_NRAY];
    e->d_pose = nullptr; e->d_next = nullptr; e->rec_words = 0; e->pair_stride = 0; e->rec_stride = 0;
    const bool generic = getenv("B2C_ENV_GENERIC") != nullptr || e->n_ray != LIDAR_RAYS;
    e->step_fn = pick_step_kernel(e->split != 0, k.A, k.D, generic);
    e->lidar_fn = pick_lidar_kernel(k.A, k.D, generic);
    e->specialised = (e->step_fn != (e->split ? env_step_kernel<true, 0, 0> : env_step_kernel<false, 0, 0>)) ? 1 : 0;
    // measured on B200 (profiles/r02_b_env_pdl.md, C2 shape, back-to-back steps): none 0.1022 ms, lidar early 0.1136,
    // state early 0.0989, both 0.1089 - the state kernel gains from loading its tiles while its predecessor drains, the
    // lidar kernel's early CTAs only take shared memory away from the state kernel's last wave; so only bit 1 is on
    e->pdl = 2;
    if (const char* t = getenv("B2C_ENV_PDL")) e->pdl = atoi(t) & 3;
    if (e->split) {
        B2C_CUDA_OR(cudaFuncSetAttribute(e->step_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem),
                    delete e);
        e->lidar_threads = 128;
        if (const char* t = getenv("B2C_LIDAR_THREADS")) e->lidar_threads = atoi(t);
        if (e->lidar_threads < 32 || e->lidar_threads > ENV_MAX_THREADS || (e->lidar_threads & 31)) e->lidar_threads = 128;
        e->lidar_ctas = 9;                          // resident CTAs per SM (shared memory); each walks several scenes, so
        if (const char* t = getenv("B2C_LIDAR_CTAS")) e->lidar_ctas = atoi(t) > 0 ? atoi(t) : 9;   // the prologue is paid once
        e->ray_off = (int)map_blob[M_OFF_RAY];
        e->lidar_group = 1;
        if (const char* g = getenv("B2C_LIDAR_GROUP")) e->lidar_group = atoi(g) > 0 ? atoi(g) : 1;
        e->pair_stride = (k.A * k.A + 7) & ~7;
        e->rec_stride = k.A | 1;
        e->rec_words = (6 * k.A + 4 + (k.D - e->n_ray) * e->rec_stride + 3) & ~3;
        auto lsm = [&](int g) {
            return (size_t)lidar_plan(g, k.A, k.D, e->n_ray, e->rec_words, e->rec_stride, e->pair_stride,
                                      e->lidar_threads / 32).total;
        };
        while (e->lidar_group > 1 && lsm(e->lidar_group) > 100 * 1024) e->lidar_group -= 1;
        if (e->lidar_group > k.S) e->lidar_group = k.S;
        e->lidar_smem = lsm(e->lidar_group);
        B2C_CUDA_OR(cudaFuncSetAttribute(e->lidar_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->lidar_smem),
                    delete e);
        B2C_CUDA_OR(cudaMalloc(&e->d_pose, (size_t)k.S * e->rec_words * 4), delete e);
        B2C_CUDA_OR(cudaMalloc(&e->d_next, sizeof(int)), delete e);
        B2C_CUDA_OR(cudaMemset(e->d_next, 0, sizeof(int)), delete e);
    } else {
        B2C_CUDA_OR(cudaFuncSetAttribute(e->step_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem),
                    delete e);
    }
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
    cudaFree(e->d_pose);
    cudaFree(e->d_next);
    delete e;
    return B2C_OK;
}

static void launch_lidar(b2c_env* e, float* obs, uint32_t* obs_split, int kp, cudaStream_t st, int first = 0,
                         int count = -1) {
    EnvConfig cfg = e->cfg;
    if (count >= 0) cfg.S = count;
    LidarIO li;
    li.next_group = getenv("B2C_LIDAR_STATIC") ? nullptr : e->d_next;
    li.map = e->d_map; li.pose = e->d_pose ? e->d_pose + (size_t)first * e->rec_words : nullptr; li.obs = obs; li.obs_split = obs_split;
    li.pair_stride = e->pair_stride; li.rec_words = e->rec_words; li.n_ray = e->n_ray; li.ray_off = e->ray_off;
    li.rec_stride = e->rec_stride;
    li.kp = kp; li.group = e->lidar_group; li.S = cfg.S; li.A = cfg.A; li.D = cfg.D;
    // one CTA per group of scenes (the hardware scheduler balances the uneven pair counts); very large batches loop
    int lg = (cfg.S + e->lidar_group - 1) / e->lidar_group;
    int lgrid = e->num_sms * e->lidar_ctas;
    if (lgrid > lg) lgrid = lg;
    launch_pdl(e->lidar_fn, lgrid, e->lidar_threads, e->lidar_smem, st, (e->pdl & 1) != 0, li);
}

static int launch_env(b2c_env* e, const float* actions, const b2c_env_io* o, int do_reset, int new_episode,
                      void* stream, int first = 0, int count = -1) {
    if (!e || !o || !o->obs || !o->reward || !o->flags)
        return b2c_set_error(B2C_ERR_ARG, "b2c_env_step: obs, reward and flags outputs are required");
    if (!do_reset && !actions) return b2c_set_error(B2C_ERR_ARG, "b2c_env_step: actions is null");
    if (count < 0) count = e->cfg.S;
    if (first < 0 || count < 1 || first + count > e->cfg.S)
        return b2c_set_error(B2C_ERR_ARG, "b2c_env_step_scenes: scene range out of bounds");
    EnvConfig cfg = e->cfg;
    cfg.do_reset = do_reset; cfg.new_episode = new_episode;
    cfg.S = count; cfg.scene_offset += first;            // the kernels index scenes from the start of the range
    EnvIO io;
    io.map = e->d_map; io.state = e->d_state + (size_t)first * e->tile_words; io.actions = actions; io.obs = o->obs;
    io.reward = o->reward;
    io.flags = o->flags; io.nei_mask = (unsigned long long*)o->nei_mask; io.mf_mask = (unsigned long long*)o->mf_mask;
    io.nei_reward = o->nei_reward; io.global_reward = o->global_reward; io.nei_list = o->nei_list;
    io.agent_id = o->agent_id; io.lcf = o->lcf; io.scene_done = o->scene_done;
    io.obs_split = (uint32_t*)o->obs_split; io.kp = (cfg.D + 63) / 64 * 64;
    io.map_words = e->map_words; io.tile_words = e->tile_words;
    io.obs_bulk = ((((size_t)cfg.A * cfg.D * 4) % 16 == 0) && (((uintptr_t)o->obs) % 16 == 0)) ? 1 : 0;
    io.group = e->group;
    io.pose = e->d_pose ? e->d_pose + (size_t)first * e->rec_words : nullptr; io.rec_words = e->rec_words;
    io.rec_stride = e->rec_stride; io.lidar_next = e->d_next;
    io.pl = smem_plan(e->group, cfg.A, cfg.D, e->map_words, e->tile_words, e->threads / 32, e->split != 0);
    int ctas_per_sm = (int)(227 * 1024 / (e->smem + 1024));
    int by_threads = 2048 / e->threads;
    if (ctas_per_sm > by_threads) ctas_per_sm = by_threads;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int n_groups = (cfg.S + e->group - 1) / e->group;
    int grid = e->num_sms * ctas_per_sm;
    if (grid > n_groups) grid = n_groups;
    cudaStream_t st = (cudaStream_t)stream;
    if (!e->split) {
        e->step_fn<<<grid, e->threads, e->smem, st>>>(cfg, io);
    } else {
        launch_pdl(e->step_fn, grid, e->threads, e->smem, st, (e->pdl & 2) != 0, cfg, io);
        B2C_CUDA(cudaGetLastError());
        launch_lidar(e, o->obs, io.obs_split, io.kp, st, first, count);
    }
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}

int b2c_env_reset(b2c_env* e, const b2c_env_io* out, int new_episode, void* stream) {
    return launch_env(e, nullptr, out, 1, new_episode, stream);
}
int b2c_env_step(b2c_env* e, const float* actions, const b2c_env_io* out, void* stream) {
    return launch_env(e, actions, out, 0, 0, stream);
}
int b2c_env_step_scenes(b2c_env* e, const float* actions, const b2c_env_io* out, int first_scene, int num_scenes,
                        void* stream) {
    return launch_env(e, actions, out, 0, 0, stream, first_scene, num_scenes);
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
int b2c_env_relaunch_lidar(b2c_env* e, const b2c_env_io* o, void* stream) {
    if (!e || !o || !o->obs) return b2c_set_error(B2C_ERR_ARG, "b2c_env_relaunch_lidar: null argument");
    if (!e->split) return b2c_set_error(B2C_ERR_STATE, "b2c_env_relaunch_lidar: the env runs the fused kernel");
    B2C_CUDA(cudaMemsetAsync(e->d_next, 0, sizeof(int), (cudaStream_t)stream));     // the state kernel's job in a full step
    launch_lidar(e, o->obs, (uint32_t*)o->obs_split, (e->cfg.D + 63) / 64 * 64, (cudaStream_t)stream);
    B2C_CUDA(cudaGetLastError());
    return B2C_OK;
}
int b2c_env_obs_dim(const b2c_env* e) { return e ? e->cfg.D : -1; }
int b2c_env_kernels_per_step(const b2c_env* e) { return e ? (e->split ? 2 : 1) : -1; }
int b2c_env_shape_specialised(const b2c_env* e) { return e ? e->specialised : -1; }
int b2c_env_obs_split_width(const b2c_env* e) { return e ? 2 * ((e->cfg.D + 63) / 64 * 64) : -1; }
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
