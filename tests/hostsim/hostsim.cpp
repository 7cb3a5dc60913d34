// TEST-ONLY: host build of the device phases in copo_b200/csrc/sim_core.cuh.
// Lets the kernel's step logic be compared with oracle/sim.py without a GPU (items run sequentially where
// the kernel runs them on threads).  Never loaded by the copo_b200 package - there is no CPU fallback.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC hostsim.cpp -o _hostsim.so
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../copo_b200/csrc/sim_core.cuh"

using namespace b2c;

struct HostIO {
    float* obs; float* reward; uint8_t* flags; unsigned long long* nei_mask; unsigned long long* mf_mask;
    float* nei_reward; float* global_reward; int8_t* nei_list; int32_t* agent_id; float* lcf; uint8_t* scene_done;
};

extern "C" int hostsim_step(const EnvConfig* cfgp, const uint32_t* map, int map_words, uint32_t* state,
                            const float* actions, const HostIO* io) {
    EnvConfig cfg = *cfgp;
    const int A = cfg.A, AP = cfg.AP, D = cfg.D;
    const int tile_words = NUM_FIELDS * AP + HEADER_WORDS;
    std::vector<float> fbuf(6 * A), obs((size_t)A * D);
    std::vector<int> ibuf(5 * A + MAX_SPAWN);
    std::vector<uint8_t> cand((size_t)A * A);
    for (int scene = 0; scene < cfg.S; ++scene) {
        SceneView v;
        v.map = map; v.st = state + (size_t)scene * tile_words; v.obs = obs.data();
        float* s_f = fbuf.data(); int* s_i = ibuf.data();
        v.cs = s_f; v.sn = s_f + A; v.rew = s_f + 2 * A; v.long_last = s_f + 3 * A; v.loc_s = s_f + 4 * A;
        v.loc_l = s_f + 5 * A;
        v.flags = s_i; v.crash = s_i + A; v.acted = s_i + 2 * A; v.linger = s_i + 3 * A; v.ncand = s_i + 4 * A;
        v.cand = cand.data(); v.place_free = s_i + 5 * A; v.A = A; v.AP = AP; v.D = D;
        if (cfg.do_reset) {
            for (int i = 0; i < A; ++i) phase_reset_slot(v, cfg, i);
            phase_reset_scene(v, cfg);
        } else {
            v.hdr(H_EP_STEP) += 1;
        }
        for (int i = 0; i < A; ++i) {
            float a0 = 0.f, a1 = 0.f;
            if (!cfg.do_reset) { a0 = actions[((size_t)scene * A + i) * 2]; a1 = actions[((size_t)scene * A + i) * 2 + 1]; }
            phase_dynamics(v, cfg, i, a0, a1);
        }
        if (!cfg.do_reset)
            for (int i = 0; i < A; ++i)
                for (int j = i + 1; j < A; ++j)
                    if (phase_pair_crash(v, i, j)) { v.crash[i] = 1; v.crash[j] = 1; }
        for (int i = 0; i < A; ++i) phase_outcome(v, cfg, i);
        for (int p = 0; p < (int)map[M_NSPAWN]; ++p) phase_place_free(v, cfg, p);
        int sd = phase_respawn(v, cfg, scene);
        for (int i = 0; i < A; ++i) phase_pose_refresh(v, i);
        for (int i = 0; i < A; ++i) {
            NeiOut n = phase_neighbours(v, cfg, i);
            size_t g = (size_t)scene * A + i;
            io->nei_mask[g] = n.nei_mask; io->mf_mask[g] = n.mf_mask; io->nei_reward[g] = n.nei_reward;
            for (int k = 0; k < NEI_K; ++k) io->nei_list[g * NEI_K + k] = n.list[k];
            io->reward[g] = v.rew[i]; io->flags[g] = (uint8_t)v.flags[i];
            io->agent_id[g] = v.geti(F_ID, i); io->lcf[g] = v.f(F_LCF, i);
            phase_observe_ego(v, cfg, i);
        }
        io->global_reward[scene] = phase_global_reward(v);
        io->scene_done[scene] = (uint8_t)sd;
        const int n_ray = (int)map[M_NRAY];
        for (int i = 0; i < A; ++i)
            for (int k = 0; k < n_ray; ++k) phase_lidar(v, i, k);
        for (int i = 0; i < A; ++i) v.seti(F_STATUS, i, v.status(i) | (v.linger[i] << 8));
        memcpy(io->obs + (size_t)scene * A * D, obs.data(), sizeof(float) * A * D);
    }
    return 0;
}
