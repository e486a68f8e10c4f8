// test_host.cpp — exercises the C++ host mirror (host/KartLQR.hpp, host/KartMCTS.hpp) through libhk_b200.so.
// Reads one 2-kart problem + expected u0 (written by tests/test_host_cpp.py from the oracle) and checks 1e-9 parity;
// then runs a short tree search.  Needs a B200 (the library has no CPU path); prints "HOST_OK" on success.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "KartLQR.hpp"
#include "KartMCTS.hpp"
#include "KartRace.hpp"

using namespace KartGame::AI;

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: test_host problem.txt\n"); return 2; }
    FILE* f = std::fopen(argv[1], "r");
    if (!f) return 2;
    double dt, x0[2][4], tgt[2][4], tw[2][4], cw[2], aw[2][2], otgt[2][4], otw[2][3], want[2];
    if (std::fscanf(f, "%lf", &dt) != 1) return 2;
    for (int i = 0; i < 2; ++i) {
        for (double* p : {x0[i], tgt[i], tw[i]}) for (int k = 0; k < 4; ++k) if (std::fscanf(f, "%lf", &p[k]) != 1) return 2;
        if (std::fscanf(f, "%lf %lf %lf", &cw[i], &aw[i][0], &aw[i][1]) != 3) return 2;
        for (int k = 0; k < 4; ++k) if (std::fscanf(f, "%lf", &otgt[i][k]) != 1) return 2;
        for (int k = 0; k < 3; ++k) if (std::fscanf(f, "%lf", &otw[i][k]) != 1) return 2;
    }
    if (std::fscanf(f, "%lf %lf", &want[0], &want[1]) != 2) return 2;
    std::fclose(f);

    LQR::LinearizedBicycle d0(dt, LQR::Vector(x0[0], x0[0] + 4)), d1(dt, LQR::Vector(x0[1], x0[1] + 4));
    LQR::KartLQRDynamics* dyn[2] = {&d0, &d1};
    std::vector<std::unique_ptr<LQR::LQRCheckpointReachAvoidCost>> costs;
    for (int i = 0; i < 2; ++i) {
        LQR::LQRCheckpointReachAvoidCost::Weights w{{0, tw[i][0]}, {1, tw[i][1]}, {2, tw[i][2]}, {3, tw[i][3]}};
        LQR::LQRCheckpointReachAvoidCost::Weights ow{{0, otw[i][0]}, {1, otw[i][1]}, {2, otw[i][2]}};
        costs.emplace_back(new LQR::LQRCheckpointReachAvoidCost(LQR::Vector(tgt[i], tgt[i] + 4), w, cw[i], dyn[i],
                                                               {LQR::Vector(otgt[i], otgt[i] + 4)}, {ow}, {{0, {aw[i][0]}}, {1, {aw[i][1]}}},
                                                               {{0, {0}}, {1, {1}}}, {dyn[1 - i]}));
    }
    auto u = LQR::KartLQR::solveFeedbackLQR({&d0, &d1}, {costs[0].get(), costs[1].get()}, {d0.initial(), d1.initial()}, 3);
    const double scale = std::fmax(std::fabs(want[0]), std::fabs(want[1]));
    for (int k = 0; k < 2; ++k)
        if (std::fabs(u[k] - want[k]) > 1e-9 * scale) { std::printf("LQR mismatch %d: %.17g vs %.17g\n", k, u[k], want[k]); return 1; }
    try {                                                              // dimension mismatch -> invalid_argument (ArgumentException)
        LQR::KartLQR::solveFeedbackLQR({&d0, &d1}, {costs[0].get()}, {d0.initial(), d1.initial()}, 3);
        std::printf("expected invalid_argument\n"); return 1;
    } catch (const std::invalid_argument&) {}

    // a tiny oval: 4 straights + 4 left curves, 2 karts
    std::vector<hk_section> secs;
    for (int i = 0; i < 8; ++i) secs.push_back(i % 2 ? hk_section{15, 10, 10, 45, 1, 1} : hk_section{0, 10, 10, 0, 0, 4});
    std::vector<hk_kart> karts(2, hk_kart{7, 16, 15, 10, 2, 0.5f, 0.001f});
    hk_game_params gp{2, 100, 2, 8, 3, 0.1f, 0.75f, 6000};
    auto tables = std::make_shared<MCTS::GameTables>(secs, karts, std::vector<hk_kart>{}, gp);
    hk_game_state root{};
    root.n_karts = 2; root.initialSection = 0; root.lastCompletedSection = 0; root.finalSection = 8;
    for (int i = 0; i < 2; ++i) root.karts[i] = hk_kart_state{0, i, 0, 0, 0, 2, 2 + i, 2500, 0, 0};
    MCTS::KartMCTSNode node(MCTS::DiscreteGameState(tables, root));
    MCTS::KartMCTS::rolloutsPerChild = 512;
    MCTS::KartMCTS::constructSearchTree(&node, 0.2, /*parallel=*/true);                       // leaf-parallel processLeaf, tree on the host
    auto best = MCTS::KartMCTS::getBestStatesSequence(&node);
    if (node.numEpisodes <= 0 || node.children.empty() || best.empty()) { std::printf("MCTS produced no plan\n"); return 1; }
    {   // what HierarchicalKartAgent calls: the sequential search (parallel == false), device-resident, then continued on the same root
        MCTS::KartMCTSNode seq(MCTS::DiscreteGameState(tables, root));
        MCTS::KartMCTS::iterationsPerSecond = 1000.0;
        MCTS::KartMCTS::constructSearchTree(&seq, 0.15);                                        // 150 iterations
        if (seq.numEpisodes != 150 || seq.children.empty()) { std::printf("sequential MCTS: root has %d episodes\n", seq.numEpisodes); return 1; }
        int sum = 0;
        for (auto& kv : seq.children) sum += kv.second->numEpisodes;
        if (sum != 150) { std::printf("sequential MCTS: children hold %d of 150 episodes\n", sum); return 1; }
        auto chain = MCTS::KartMCTS::getBestStatesSequence(&seq);
        if (chain.size() != 8 || seq.deviceBestStates.size() != 8) { std::printf("sequential MCTS: best states %zu / %zu, expected the chain to depth 8\n", chain.size(), seq.deviceBestStates.size()); return 1; }
        for (auto& st : chain)
            for (int i = 0; i < st.s.n_karts; ++i)
                if (st.s.karts[i].section != st.s.lastCompletedSection) { std::printf("sequential MCTS: best state with a kart behind\n"); return 1; }
        const int nodesBefore = seq.childrenAsRoot;
        MCTS::KartMCTS::constructSearchTree(&seq, 0.1);                                         // constructSearchTree(root): 100 more
        if (seq.numEpisodes != 250 || seq.childrenAsRoot <= nodesBefore) { std::printf("continued MCTS: %d episodes, %d nodes\n", seq.numEpisodes, seq.childrenAsRoot); return 1; }
    }
    {   // the same search for several roots at once on the device
        std::vector<hk_game_state> roots(5, root);
        for (int r = 0; r < 5; ++r) { roots[r].initialSection = roots[r].lastCompletedSection = r; roots[r].finalSection = r + 8; for (int i = 0; i < 2; ++i) roots[r].karts[i].section = r; }
        auto plans = MCTS::KartMCTS::searchBatch(tables, roots, 40, 64, 12345ull);
        int withPlan = 0;
        for (auto& p : plans) withPlan += !p.empty();
        if (plans.size() != 5 || withPlan == 0) { std::printf("batched tree search produced no plan\n"); return 1; }
    }
    // headless races on a square ring: 4 straights of 20 m joined by 90-degree corners, 8 checkpoints, 3 races x 2 karts
    {
        std::vector<Race::Checkpoint> ring;
        const double cx[8] = {10, 10, 5, -5, -10, -10, -5, 5}, cz[8] = {-5, 5, 10, 10, 5, -5, -10, -10};
        const double fx[8] = {0, 0, -1, -1, 0, 0, 1, 1}, fz[8] = {1, 1, 0, 0, -1, -1, 0, 0};
        for (int i = 0; i < 8; ++i) {
            Race::Checkpoint c{};
            c.section = hk_section{0, 10, 10, 0, 0, 3};
            c.trigger[0] = cx[i]; c.trigger[1] = cz[i]; c.forward[0] = fx[i]; c.forward[1] = fz[i];
            const double off[4] = {-3.5, -1.25, 1.25, 3.5};
            for (int l = 0; l < 4; ++l) { c.lane[l][0] = cx[i] + off[l] * fz[i]; c.lane[l][1] = cz[i] - off[l] * fx[i]; }
            ring.push_back(c);
        }
        hk_race_params rp{};
        rp.dt = (double)0.02f; rp.accel = 7; rp.braking = 16; rp.coastingDrag = 5; rp.topSpeed = 15; rp.gateHalfWidth = 10;
        rp.maxLaneChanges = 3; rp.goalSection = 64; rp.highModeMcts = 0; rp.velocityBucketSize = 2; rp.treeSearchDepth = 8;
        rp.planEvery = 100; rp.horizon = 3;
        Race::HeadlessRaces races(ring, rp);
        std::vector<hk_race_kart> rk(6);
        std::vector<hk_race_plan> pl(6);
        for (int i = 0; i < 6; ++i) {
            rk[i] = hk_race_kart{};
            rk[i].x = 10 + (i % 2 ? 1.25 : -1.25); rk[i].z = -4.0 + 0.3 * (i / 2); rk[i].v = 1.0; rk[i].h = 1.5707963267948966;
            rk[i].steer = 3.25f; rk[i].section = 0; rk[i].lane = 2 + i % 2; rk[i].active = 1;
            pl[i] = hk_race_plan{};
        }
        std::vector<double> ulast;
        const long long bad = races.run(rk, pl, 0, 400, &ulast);
        for (int i = 0; i < 6; ++i)
            if (bad != 0 || rk[i].section < 2 || !std::isfinite(rk[i].x) || rk[i].v < 0 || rk[i].v > 15) {
                std::printf("race: kart %d section %d v %.3f bad %lld\n", i, rk[i].section, rk[i].v, bad);
                return 1;
            }
    }
    {   // the MCTS-mode loop on the same ring (8 sections), GPU-resident planning every 100 steps
        std::vector<Race::Checkpoint> ring;
        const double cx[8] = {10, 10, 5, -5, -10, -10, -5, 5}, cz[8] = {-5, 5, 10, 10, 5, -5, -10, -10};
        const double fx[8] = {0, 0, -1, -1, 0, 0, 1, 1}, fz[8] = {1, 1, 0, 0, -1, -1, 0, 0};
        std::vector<hk_section> gsecs;
        for (int i = 0; i < 8; ++i) {
            Race::Checkpoint c{};
            c.section = hk_section{0, 10, 10, 0, 0, 3};
            gsecs.push_back(c.section);
            c.trigger[0] = cx[i]; c.trigger[1] = cz[i]; c.forward[0] = fx[i]; c.forward[1] = fz[i];
            const double off[4] = {-3.5, -1.25, 1.25, 3.5};
            for (int l = 0; l < 4; ++l) { c.lane[l][0] = cx[i] + off[l] * fz[i]; c.lane[l][1] = cz[i] - off[l] * fx[i]; }
            ring.push_back(c);
        }
        hk_race_params rp{};
        rp.dt = (double)0.02f; rp.accel = 7; rp.braking = 16; rp.coastingDrag = 5; rp.topSpeed = 15; rp.gateHalfWidth = 10;
        rp.maxLaneChanges = 3; rp.goalSection = 64; rp.highModeMcts = 1; rp.velocityBucketSize = 2; rp.treeSearchDepth = 8;
        rp.planEvery = 100; rp.horizon = 3;
        Race::HeadlessRaces races(ring, rp);
        auto gtables = std::make_shared<MCTS::GameTables>(gsecs, karts, std::vector<hk_kart>{}, gp);
        std::vector<hk_race_kart> rk(4);
        std::vector<hk_race_plan> pl(4);
        for (int i = 0; i < 4; ++i) {
            rk[i] = hk_race_kart{};
            rk[i].x = 10 + (i % 2 ? 1.25 : -1.25); rk[i].z = -4.0 + 0.3 * (i / 2); rk[i].v = 1.0; rk[i].h = 1.5707963267948966;
            rk[i].steer = 3.25f; rk[i].section = 0; rk[i].lane = 2 + i % 2; rk[i].active = 1;
            pl[i] = hk_race_plan{};
        }
        const long long bad = races.runMcts(gtables->handle(), 24, 16, 7ull, rk, pl, 0, 250);
        for (int i = 0; i < 4; ++i)
            if (bad != 0 || rk[i].section < 2 || !std::isfinite(rk[i].x)) { std::printf("mcts race: kart %d section %d bad %lld\n", i, rk[i].section, bad); return 1; }
    }
    std::printf("HOST_OK u0=(%.12g, %.12g) episodes=%d plan=%zu\n", u[0], u[1], node.numEpisodes, best.size());
    return 0;
}
