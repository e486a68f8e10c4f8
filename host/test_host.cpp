// test_host.cpp — exercises the C++ host mirror (host/KartLQR.hpp, host/KartMCTS.hpp) through libhk_b200.so.
// Reads one 2-kart problem + expected u0 (written by tests/test_host_cpp.py from the oracle) and checks 1e-9 parity;
// then runs a short tree search.  Needs a B200 (the library has no CPU path); prints "HOST_OK" on success.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "KartLQR.hpp"
#include "KartMCTS.hpp"

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
    MCTS::KartMCTS::constructSearchTree(&node, 0.2);
    auto best = MCTS::KartMCTS::getBestStatesSequence(&node);
    if (node.numEpisodes <= 0 || node.children.empty() || best.empty()) { std::printf("MCTS produced no plan\n"); return 1; }
    std::printf("HOST_OK u0=(%.12g, %.12g) episodes=%d plan=%zu\n", u[0], u[1], node.numEpisodes, best.size());
    return 0;
}
