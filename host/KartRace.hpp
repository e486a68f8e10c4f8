// KartRace.hpp — C++ host-side mirror of the headless race loop over the C-ABI (include/hk_abi.h, hk_race_*): what
// RacingEnvController.FixedUpdate + HierarchicalKartAgent.FixedUpdate / SolveLQR / OnTriggerEnter do per physics step
// (Assets/Karting/Scripts/RacingEnvController.cs:239-321, AI/HierarchicalKartAgent.cs:145-166,319-347,611-662,699-1224) for many
// independent 2-kart races, with Unity's PhysX kart replaced by the kinematic model the planners assume
// (AI/MPC/KartMPCDynamics.cs:55-70).  Everything runs in libhk_b200 (CUDA); this header only owns the handle and the arrays.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>
#include "../include/hk_abi.h"

namespace KartGame { namespace AI { namespace Race {

inline void hk_check(int status)
{
    if (status == HK_OK) return;
    std::string msg = hk_last_error();
    if (status == HK_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error("hk_b200 status " + std::to_string(status) + ": " + msg);
}

struct Checkpoint {            // one RacingEnvController.Sections[i]: DiscretePositionTracker fields + collider positions
    hk_section section;
    double trigger[2], forward[2], lane[4][2];
};

class HeadlessRaces {
public:
    HeadlessRaces(const std::vector<Checkpoint>& track, const hk_race_params& params) : params_(params), n_sections_((int)track.size())
    {
        std::vector<hk_section> sec;
        std::vector<double> trig, fwd, lane;
        for (const Checkpoint& c : track) {
            sec.push_back(c.section);
            trig.insert(trig.end(), c.trigger, c.trigger + 2);
            fwd.insert(fwd.end(), c.forward, c.forward + 2);
            for (int l = 0; l < 4; ++l) lane.insert(lane.end(), c.lane[l], c.lane[l] + 2);
        }
        hk_check(hk_track_create(sec.data(), trig.data(), fwd.data(), lane.data(), n_sections_, &track_));
    }
    ~HeadlessRaces() { hk_track_destroy(track_); }
    HeadlessRaces(const HeadlessRaces&) = delete;
    HeadlessRaces& operator=(const HeadlessRaces&) = delete;

    // karts / plans: [n_races][2], updated in place; returns the number of LQNG solves that hit a zero pivot
    long long run(std::vector<hk_race_kart>& karts, std::vector<hk_race_plan>& plans, int first_step, int n_steps, std::vector<double>* u_last = nullptr)
    {
        if (karts.size() != plans.size() || karts.size() % 2) throw std::invalid_argument("karts / plans must hold 2 entries per race");
        const int n_races = (int)karts.size() / 2;
        if (u_last) u_last->assign((size_t)n_races * 4, 0.0);
        int64_t bad = 0;
        hk_check(hk_race_run(track_, &params_, n_races, first_step, n_steps, karts.data(), plans.data(), u_last ? u_last->data() : nullptr, &bad));
        return bad;
    }
    // the same loop in HighLevelMode.MCTS (params.highModeMcts = 1): every planEvery steps every agent replans by planWithMCTS + the
    // waypoint hand-off on the GPU (HierarchicalKartAgent.cs:180-283, 331-353, 366-402); `game` is the hk_game of this track
    long long runMcts(const hk_game* game, int iterations, int rolloutsPerLeaf, unsigned long long seed, std::vector<hk_race_kart>& karts,
                      std::vector<hk_race_plan>& plans, int first_step, int n_steps)
    {
        if (karts.size() != plans.size() || karts.size() % 2) throw std::invalid_argument("karts / plans must hold 2 entries per race");
        int64_t bad = 0;
        hk_check(hk_race_run_mcts(track_, &params_, game, iterations, rolloutsPerLeaf, seed, (int)karts.size() / 2, first_step, n_steps, karts.data(),
                                  plans.data(), nullptr, &bad));
        return bad;
    }
    void planFixed(const std::vector<hk_race_kart>& karts, std::vector<hk_race_plan>& plans)       // HierarchicalKartAgent.cs:145-166
    {
        hk_check(hk_race_plan_fixed(track_, &params_, (int)karts.size(), karts.data(), plans.data()));
    }
    int sections() const { return n_sections_; }
    const hk_race_params& params() const { return params_; }

private:
    hk_race_params params_;
    int n_sections_;
    hk_track* track_ = nullptr;
};

}}}  // namespace KartGame::AI::Race
