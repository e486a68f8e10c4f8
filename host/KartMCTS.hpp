// KartMCTS.hpp — C++ host-side mirror of the reference's C# namespace KartGame.AI.MCTS over the C-ABI (include/hk_abi.h).
//   DiscreteKartAction / DiscreteKartState / DiscreteGameState   Assets/Karting/Scripts/AI/MCTS/KartDiscreteGame.cs:14-34,174-447
//   KartMCTSNode / KartMCTS                                       Assets/Karting/Scripts/AI/MCTS/KartMCTS.cs:18-38,45-290
// The game itself (transitions, legal moves, scores, playouts) is evaluated by libhk_b200 on the GPU.  constructSearchTree keeps the
// reference's two modes: parallel == false (what HierarchicalKartAgent calls, HierarchicalKartAgent.cs:250,271) is the sequential
// search — findLeaf, ONE playout whose every state becomes a node, backpropagate — run on the device (hk_mcts_forest_search) and
// handed back as a KartMCTSNode graph that a later constructSearchTree(root) continues; parallel == true is the reference's
// leaf-parallel processLeaf (:124-159) with `rolloutsPerChild` playouts per child, tree on the host.
#pragma once
#include <algorithm>
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

struct ForestHandle {                                       // device-resident tree behind a root built by the sequential mode
    hk_mcts_forest* f = nullptr;
    int maxNodes = 0, iterations = 0;
    ~ForestHandle() { hk_mcts_forest_destroy(f); }
};

class KartMCTSNode {                                        // KartMCTS.cs:18-38
public:
    DiscreteGameState state;
    KartMCTSNode* parent;
    std::map<DiscreteKartAction, std::unique_ptr<KartMCTSNode>, ActionLess> children;
    std::vector<DiscreteKartAction> insertionOrder;         // the Dictionary's enumeration order (Keys.ElementAt / foreach, :169-177)
    float totalValue = 0.0f;
    int numEpisodes = 0;
    int childrenAsRoot = 0;
    std::string createdBy;
    std::shared_ptr<ForestHandle> forest;                   // root of a device-resident tree only
    std::vector<hk_game_state> deviceBestStates;            // getBestStatesSequence as the device walked it (Philox picks)
    KartMCTSNode(DiscreteGameState st, KartMCTSNode* p = nullptr, std::string by = "") : state(std::move(st)), parent(p), createdBy(std::move(by)) {}
    KartMCTSNode* addChild(const DiscreteKartAction& a, std::unique_ptr<KartMCTSNode> c)
    {
        KartMCTSNode* raw = c.get();
        children[a] = std::move(c);
        insertionOrder.push_back(a);
        return raw;
    }
};

class KartMCTS {
public:
    static inline long long rolloutsPerChild = 4096;
    static inline double iterationsPerSecond = 600.0;       // sequential mode: how the wall-clock budget T (:55) maps to an iteration count
    static inline int reserveSearches = 3;                  // a new device tree is sized for this many calls (CyclesRootProcessed < 3, HKA:265)
    static inline std::mt19937_64 random{std::random_device{}()};

    // The sequential search on the device: a new tree if `root` has none (constructSearchTree(state), :50-78), else the root's tree is
    // continued (constructSearchTree(root), :80-106).  The node graph below `root` is rebuilt from the device's node records.
    static KartMCTSNode* constructSequential(KartMCTSNode* root, int iterations)
    {
        const hk_game_state& rs = root->state.s;
        int plies = 0;
        for (int k = 0; k < rs.n_karts; ++k) plies += std::max(0, rs.finalSection - rs.karts[k].section);
        plies = std::max(plies, 1);
        hk_game_state best[HK_MCTS_MAX_SEQ];
        int32_t nBest = 0, nNodes = 0, status = 0, fresh = 0;
        if (!root->forest) {
            if (!root->children.empty()) throw std::invalid_argument("constructSearchTree(root): root was not built by the sequential search");
            auto h = std::make_shared<ForestHandle>();
            h->maxNodes = 1 + reserveSearches * iterations * plies;
            hk_check(hk_mcts_forest_create(root->state.tables->handle(), 1, h->maxNodes, &h->f));
            root->forest = h;
            fresh = 1;
        }
        hk_check(hk_mcts_forest_search(root->forest->f, &rs, &fresh, iterations, random(), best, &nBest, &nNodes, &status));
        if (status == 2) throw std::domain_error("division by zero");                                        // DivideByZeroException in findLeaf
        root->forest->iterations += iterations;
        std::vector<hk_mcts_node> rec((size_t)nNodes);
        hk_check(hk_mcts_forest_nodes(root->forest->f, 0, rec.data(), nNodes, &nNodes));
        // states are not stored on the device: replay every node's action path in one batched call
        std::vector<int> parent((size_t)nNodes, -1), depth((size_t)nNodes, 0);
        for (int i = 0; i < nNodes; ++i)
            for (int c = rec[i].first_child; c >= 0; c = rec[c].next_sibling) { parent[c] = i; depth[c] = depth[i] + 1; }
        int maxDepth = 0;
        for (int d : depth) maxDepth = std::max(maxDepth, d);
        const hk_game_params& gp = root->state.tables->params();
        auto actionOf = [&](int gi) { const int v = 6 + (gi >> 2) * gp.velocityBucketSize; return hk_action{v, std::min(v + gp.velocityBucketSize, vmaxOf(root)), (gi & 3) + 1}; };
        std::vector<hk_game_state> roots((size_t)nNodes, rs), states((size_t)nNodes * (maxDepth + 1));
        std::vector<hk_action> acts((size_t)nNodes * std::max(maxDepth, 1), hk_action{6, 6 + gp.velocityBucketSize, 1});
        for (int i = 1; i < nNodes; ++i)
            for (int n = i; parent[n] >= 0; n = parent[n]) acts[(size_t)i * maxDepth + depth[n] - 1] = actionOf(rec[n].gen);
        if (maxDepth > 0)
            hk_check(hk_game_replay_batch(root->state.tables->handle(), nNodes, maxDepth, roots.data(), acts.data(), states.data(), nullptr, nullptr,
                                          nullptr, nullptr, nullptr, nullptr, nullptr));
        root->children.clear(); root->insertionOrder.clear();
        root->totalValue = rec[0].totalValue; root->numEpisodes = rec[0].numEpisodes; root->childrenAsRoot = nNodes - 1;
        std::vector<KartMCTSNode*> node((size_t)nNodes, nullptr);
        node[0] = root;
        for (int i = 0; i < nNodes; ++i)                                  // creation order: a parent precedes its children
            for (int c = rec[i].first_child; c >= 0; c = rec[c].next_sibling) {
                auto ch = std::make_unique<KartMCTSNode>(DiscreteGameState(root->state.tables, states[(size_t)c * (maxDepth + 1) + depth[c]]), node[i]);
                ch->totalValue = rec[c].totalValue; ch->numEpisodes = rec[c].numEpisodes;
                node[c] = node[i]->addChild(actionOf(rec[c].gen), std::move(ch));
            }
        root->deviceBestStates.assign(best, best + nBest);
        return root;
    }

    static KartMCTSNode* constructSearchTree(KartMCTSNode* root, double T = 0.09, bool parallel = false)   // :50-78, :80-106
    {
        if (!parallel) return constructSequential(root, std::max(1, (int)(T * iterationsPerSecond)));
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
        std::uniform_int_distribution<size_t> pick(0, node->insertionOrder.size() - 1);
        DiscreteKartAction best = node->insertionOrder[pick(random)];                                        // Keys.ElementAt(random.Next(Count))
        float best_uct = UCTWeight(node->children[best].get());
        for (auto& key : node->insertionOrder) {                                                             // foreach (var item in node.children)
            const float w = UCTWeight(node->children[key].get());
            if (w > best_uct) { best_uct = w; best = key; }
        }
        return best;
    }

private:
    static int vmaxOf(KartMCTSNode*) { return 15; }                                                          // (int)GetMaxSpeed() of the shipped karts
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
            auto it = node->children.find(mv);
            KartMCTSNode* kid = it != node->children.end() ? it->second.get()
                                                           : (++created, node->addChild(mv, std::make_unique<KartMCTSNode>(node->state.makeMove(mv), node)));
            kids.push_back(kid);
            leaves.push_back(kid->state.s);
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
