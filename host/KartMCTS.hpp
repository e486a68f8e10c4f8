// KartMCTS.hpp — C++ host-side mirror of the reference's C# namespace KartGame.AI.MCTS over the C-ABI (include/hk_abi.h).
//   DiscreteKartAction / DiscreteKartState / DiscreteGameState   Assets/Karting/Scripts/AI/MCTS/KartDiscreteGame.cs:14-34,174-447
//   KartMCTSNode / KartMCTS                                       Assets/Karting/Scripts/AI/MCTS/KartMCTS.cs:18-38,45-290
// The game itself (transitions, legal moves, scores, rollouts) is evaluated by libhk_b200 on the GPU; the search tree is
// host-side as in the reference.  One tree iteration = expand every legal child of the selected leaf and play
// `rolloutsPerChild` rollouts from each in ONE launch — the reference's own leaf-parallel processLeaf (:124-159) with R > 1.
#pragma once
#include <chrono>
#include <cmath>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>
#include "../include/hk_abi.h"

namespace KartGame { namespace AI { namespace MCTS {

inline void hk_check(int status)
{
    if (status == HK_OK) return;
    std::string msg = hk_last_error();
    if (status == HK_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    if (status == HK_ERR_NO_UPNEXT) throw std::out_of_range(msg);            // ArgumentOutOfRangeException at KartDiscreteGame.cs:326
    throw std::runtime_error("hk_b200 status " + std::to_string(status) + ": " + msg);
}

using DiscreteKartAction = hk_action;
using DiscreteKartState = hk_kart_state;
struct ActionLess {
    bool operator()(const hk_action& a, const hk_action& b) const
    {
        return std::tie(a.min_velocity, a.max_velocity, a.lane) < std::tie(b.min_velocity, b.max_velocity, b.lane);
    }
};

// envController + kartAgents + gameParams of a DiscreteGameState, as an immutable device-resident handle
class GameTables {
public:
    GameTables(const std::vector<hk_section>& sections, const std::vector<hk_kart>& karts, const std::vector<hk_kart>& envKarts,
               const hk_game_params& params) : params_(params)
    {
        hk_check(hk_game_create(sections.data(), (int)sections.size(), karts.data(), (int)karts.size(),
                                envKarts.empty() ? nullptr : envKarts.data(), (int)envKarts.size(), &params, &g_));
    }
    ~GameTables() { hk_game_destroy(g_); }
    GameTables(const GameTables&) = delete;
    GameTables& operator=(const GameTables&) = delete;
    hk_game* handle() const { return g_; }
    const hk_game_params& params() const { return params_; }
private:
    hk_game* g_ = nullptr;
    hk_game_params params_;
};

class DiscreteGameState {                                   // KartDiscreteGame.cs:174-447
public:
    std::shared_ptr<GameTables> tables;
    hk_game_state s;
    DiscreteGameState(std::shared_ptr<GameTables> t, const hk_game_state& st) : tables(std::move(t)), s(st) {}

    int upNext() { query(); return upnext_; }                                        // :188
    std::pair<bool, std::vector<float>> isOver()                                     // :251
    {
        query();
        if (upnext_ < 0) throw std::out_of_range("upNext() == -1");
        return {over_ != 0, scores_};
    }
    std::vector<DiscreteKartAction> nextMoves()                                      // :322 (generation order)
    {
        query();
        if (upnext_ < 0) throw std::out_of_range("upNext() == -1");
        return moves_gen_;
    }
    DiscreteGameState makeMove(const DiscreteKartAction& a)                          // :420
    {
        hk_game_state out[2];
        hk_check(hk_game_replay_batch(tables->handle(), 1, 1, &s, &a, out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
        return DiscreteGameState(tables, out[1]);
    }
private:
    void query()
    {
        if (queried_) return;
        int up = 0, over = 0, ns = 0, nm = 0, idx[HK_MAX_ACTIONS];
        float sc[2 * HK_MAX_KARTS];
        hk_action mv[HK_MAX_ACTIONS];
        hk_check(hk_game_replay_batch(tables->handle(), 1, 0, &s, nullptr, nullptr, &up, &over, &ns, sc, &nm, mv, idx));
        upnext_ = up; over_ = over;
        scores_.assign(sc, sc + ns);
        std::map<int, hk_action> byGen;
        for (int k = 0; k < nm; ++k) byGen[idx[k]] = mv[k];
        for (auto& kv : byGen) moves_gen_.push_back(kv.second);
        queried_ = true;
    }
    bool queried_ = false;
    int upnext_ = -1, over_ = 0;
    std::vector<float> scores_;
    std::vector<DiscreteKartAction> moves_gen_;
};

class KartMCTSNode {                                        // KartMCTS.cs:18-38
public:
    DiscreteGameState state;
    KartMCTSNode* parent;
    std::map<DiscreteKartAction, std::unique_ptr<KartMCTSNode>, ActionLess> children;
    float totalValue = 0.0f;
    int numEpisodes = 0;
    int childrenAsRoot = 0;
    std::string createdBy;
    KartMCTSNode(DiscreteGameState st, KartMCTSNode* p = nullptr, std::string by = "") : state(std::move(st)), parent(p), createdBy(std::move(by)) {}
};

class KartMCTS {
public:
    static inline long long rolloutsPerChild = 4096;
    static inline std::mt19937_64 random{std::random_device{}()};

    static KartMCTSNode* constructSearchTree(KartMCTSNode* root, double T = 0.09, bool /*parallel*/ = false)   // :80-106
    {
        double total = 0.0;
        unsigned long long it = 0;
        const unsigned long long seed = random();
        while (total < T) {
            auto t0 = std::chrono::steady_clock::now();
            KartMCTSNode* leaf = findLeaf(root);
            root->childrenAsRoot += processLeaf(leaf, seed, it++ * (unsigned long long)rolloutsPerChild * HK_MAX_ACTIONS);
            total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        return root;
    }

    // constructSearchTree + getBestStatesSequence for many roots in one GPU launch (hk_mcts_search_batch, one thread block per
    // tree): what planWithMCTS (HierarchicalKartAgent.cs:194-283) does per agent, for every agent of many races at once
    static std::vector<std::vector<DiscreteGameState>> searchBatch(const std::shared_ptr<GameTables>& tables, const std::vector<hk_game_state>& roots,
                                                                   int iterations, int rolloutsPerLeaf, unsigned long long seed)
    {
        const int n = (int)roots.size();
        std::vector<hk_game_state> best((size_t)n * HK_MCTS_MAX_SEQ);
        std::vector<int32_t> nBest(n);
        hk_check(hk_mcts_search_batch(tables->handle(), roots.data(), n, iterations, rolloutsPerLeaf, seed, best.data(), nBest.data(), nullptr,
                                      nullptr, nullptr));
        std::vector<std::vector<DiscreteGameState>> out(n);
        for (int r = 0; r < n; ++r)
            for (int k = 0; k < nBest[r]; ++k) out[r].push_back(DiscreteGameState(tables, best[(size_t)r * HK_MCTS_MAX_SEQ + k]));
        return out;
    }

    static std::vector<DiscreteGameState> getBestStatesSequence(KartMCTSNode* node)                          // :108-122
    {
        std::vector<DiscreteGameState> best;
        try {
            while (!node->children.empty()) {
                node = node->children[upperConfidenceStrategy(node)].get();
                bool all = true;
                for (int i = 0; i < node->state.s.n_karts; ++i) all &= node->state.s.karts[i].section == node->state.s.lastCompletedSection;
                if (all) best.push_back(node->state);
            }
        } catch (const std::domain_error&) {}                                                                 // DivideByZeroException
        return best;
    }

    static DiscreteKartAction upperConfidenceStrategy(KartMCTSNode* node)                                    // :167-192
    {
        std::uniform_int_distribution<size_t> pick(0, node->children.size() - 1);
        auto it = node->children.begin();
        std::advance(it, pick(random));
        DiscreteKartAction best = it->first;
        float best_uct = UCTWeight(it->second.get());
        for (auto& kv : node->children) {
            const float w = UCTWeight(kv.second.get());
            if (w > best_uct) { best_uct = w; best = kv.first; }
        }
        return best;
    }

private:
    static float UCTWeight(KartMCTSNode* n)                                                                   // :162-165 (integer division, no sqrt term)
    {
        if (n->numEpisodes == 0) throw std::domain_error("division by zero");
        return (n->totalValue / (float)n->numEpisodes) + std::sqrt(1.0f) * std::log((float)(n->parent->numEpisodes / n->numEpisodes));
    }
    static KartMCTSNode* findLeaf(KartMCTSNode* root)                                                         // :194-201
    {
        while (!root->children.empty() && root->children.size() == root->state.nextMoves().size())
            root = root->children[upperConfidenceStrategy(root)].get();
        return root;
    }
    static void backpropagate(KartMCTSNode* node, const std::vector<float>& result, int count)               // :280-289
    {
        for (; node; node = node->parent) { node->totalValue += result[node->state.upNext()] * count; node->numEpisodes += count; }
    }
    static int processLeaf(KartMCTSNode* node, unsigned long long seed, unsigned long long offset)           // GPU form of :124-159
    {
        auto over = node->state.isOver();
        if (over.first) { backpropagate(node, over.second, 1); return 0; }
        auto moves = node->state.nextMoves();
        int created = 0;
        std::vector<KartMCTSNode*> kids;
        std::vector<hk_game_state> leaves;
        for (auto& mv : moves) {
            auto& slot = node->children[mv];
            if (!slot) { slot.reset(new KartMCTSNode(node->state.makeMove(mv), node)); ++created; }
            kids.push_back(slot.get());
            leaves.push_back(slot->state.s);
        }
        const int n = (int)kids.size(), K = node->state.s.n_karts;
        std::vector<int64_t> visit((size_t)n * HK_MAX_ACTIONS), nan((size_t)n * HK_MAX_ACTIONS), plies(n);
        std::vector<double> reward((size_t)n * HK_MAX_ACTIONS * HK_MAX_KARTS);
        hk_check(hk_mcts_rollouts_multi(node->state.tables->handle(), leaves.data(), n, rolloutsPerChild, seed, offset, visit.data(),
                                        reward.data(), nan.data(), plies.data()));
        for (int j = 0; j < n; ++j) {
            long long cnt = 0;
            double sum[HK_MAX_KARTS] = {0, 0, 0, 0};
            for (int a = 0; a < HK_MAX_ACTIONS; ++a) {
                cnt += visit[(size_t)j * HK_MAX_ACTIONS + a] - nan[(size_t)j * HK_MAX_ACTIONS + a];
                for (int k = 0; k < K; ++k) sum[k] += reward[((size_t)j * HK_MAX_ACTIONS + a) * HK_MAX_KARTS + k];
            }
            if (cnt == 0) { auto o = kids[j]->state.isOver(); if (o.first) backpropagate(kids[j], o.second, (int)rolloutsPerChild); continue; }
            for (KartMCTSNode* nd = kids[j]; nd; nd = nd->parent) { nd->totalValue += (float)sum[nd->state.upNext()]; nd->numEpisodes += (int)cnt; }
        }
        return created;
    }
};

}}}  // namespace KartGame::AI::MCTS
