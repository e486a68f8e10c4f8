// hk_game.cuh — device-resident immutable description of one discrete race game (track + karts + parameters).
#pragma once
#include <cuda_runtime.h>
#include "../../include/hk_abi.h"

#define HK_MAX_SECTIONS 64
#define HK_MAX_ENV_KARTS 16
#define HK_MAX_TYPES 24

namespace hk {

struct DevGame {
    int n_sections, n_karts, n_env_karts;
    int vmax;                                   // (int)GetMaxSpeed()  KartDiscreteGame.cs:329
    int n_cand;                                 // velocity levels x 4 lanes (generation-order index space)
    int pad_[2];
    hk_game_params p;
    hk_kart karts[HK_MAX_KARTS];                // DiscreteGameState.kartAgents[i].m_Kart constants
    hk_kart env_karts[HK_MAX_ENV_KARTS];        // envController.Agents[player].m_Kart constants
    hk_section sections[HK_MAX_SECTIONS];
    uint32_t cdf[HK_MAX_ACTIONS + 1][HK_MAX_ACTIONS];   // rollout-policy index distribution per legal-move count
    // ---- transition tables (built on the device by build_tables_kernel at hk_game_create) -------------------------------
    // The time update of applyAction is a pure function of (section geometry, lane, velocity bucket, action) because the
    // reference evaluates computeTOC with tyre wear 0 (quirk B.6-3), so it — and with it the static part of the rollout
    // policy's sort order — is tabulated per geometry TYPE; tyre load per (type, lane pair, max_velocity) likewise.
    int tables_ok;                              // 0: more than HK_MAX_TYPES geometries -> kernels use the direct path only
    int n_types, nv;                            // geometry types; velocity levels (n_cand = 4 nv)
    int off_dt, off_order, off_load, off_radius, off_lmask, table_bytes;   // byte offsets into `tables`
    int off_od, pad2_;                          // u64[T][4][nv][3][nc]: per rank of the policy order (time update << 8 | generation index)
    unsigned char type_of[HK_MAX_SECTIONS];     // section -> type
    unsigned char rep_section[HK_MAX_TYPES];    // a section of that type
    unsigned char sec_flags[HK_MAX_SECTIONS];   // bit0 straight(s), bit1 straight(s) != straight(s+1), bits 2-3 optimalLaneSign + 1
    float radius_tab[HK_MAX_TYPES * 16];        // radiusOfLane per (type, lane, target lane): read four times per ply, so it travels with this
                                                // struct into shared memory (filled by build_tables_kernel)
    const unsigned char* tables;                // device blob: dt int32[T][4][nv][nc] | order u8[T][4][nv][3][nc] | load f32[T][16][nv] | radius f32[T][16]
                                                //              | lmask u64[T][4][nv][3][4][nv]: ranks of the moves into lane l1 with velocity level <= j
};
static_assert(sizeof(DevGame) % 4 == 0, "DevGame is copied word-wise");

void policy_cdf_host(int cnt, uint32_t* cdf);

struct ThreadCtx;
// hk_mcts_search_batch on device pointers (all of them for n_roots roots; d_eps / d_vals / d_nnodes may be null): chunks of roots, tree
// slabs in the calling thread's scratch slot 12, everything enqueued on `s`.  d_status[r] = 1 where upNext() == -1 was reached.
int mcts_search_device(const hk_game* g, const hk_game_state* d_roots, int n_roots, int iterations, int rollouts_per_leaf, uint64_t seed,
                       hk_game_state* d_best, int* d_nbest, int* d_eps, double* d_vals, int* d_nnodes, int* d_status, ThreadCtx* c,
                       cudaStream_t s);
// hk_mcts_forest_search on device pointers (d_fresh may be null = all fresh; a negative entry skips that tree), enqueued on `s`.
// max_plies: upper bound on the plies of a playout from any root of the forest (sizes the fast path's records; 0 = HK_MAX_PLIES).
int mcts_seq_search_device(hk_mcts_forest* f, const hk_game_state* d_roots, const int* d_fresh, int iterations, uint64_t seed,
                           hk_game_state* d_best, int* d_nbest, int* d_nnodes, int* d_status, cudaStream_t s, bool clear_best = true,
                           int max_plies = 0);
const hk_game_params& game_params_of(const hk_game* g);
int game_karts_of(const hk_game* g);

}  // namespace hk
