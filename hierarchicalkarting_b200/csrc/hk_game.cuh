// hk_game.cuh — device-resident immutable description of one discrete race game (track + karts + parameters).
#pragma once
#include "../../include/hk_abi.h"

#define HK_MAX_SECTIONS 64
#define HK_MAX_ENV_KARTS 16

namespace hk {

struct DevGame {
    int n_sections, n_karts, n_env_karts;
    int vmax;                                   // (int)GetMaxSpeed()  KartDiscreteGame.cs:329
    int n_cand;                                 // velocity levels x 4 lanes (generation-order index space)
    int pad_[3];
    hk_game_params p;
    hk_kart karts[HK_MAX_KARTS];                // DiscreteGameState.kartAgents[i].m_Kart constants
    hk_kart env_karts[HK_MAX_ENV_KARTS];        // envController.Agents[player].m_Kart constants
    hk_section sections[HK_MAX_SECTIONS];
    uint32_t cdf[HK_MAX_ACTIONS + 1][HK_MAX_ACTIONS];   // rollout-policy index distribution per legal-move count
};
static_assert(sizeof(DevGame) % 4 == 0, "DevGame is copied word-wise");

void policy_cdf_host(int cnt, uint32_t* cdf);

}  // namespace hk
