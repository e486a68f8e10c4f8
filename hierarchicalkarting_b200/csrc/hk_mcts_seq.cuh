// hk_mcts_seq.cuh — the reference's tree search AS ITS CALLERS RUN IT, on the device, one THREAD per tree (included by hk_game.cu).
//
// HierarchicalKartAgent never passes `parallel` (HierarchicalKartAgent.cs:250,271), so what the reference executes is the sequential
// branch of KartMCTS.constructSearchTree (Assets/Karting/Scripts/AI/MCTS/KartMCTS.cs:50-78 / :80-106, lines :61-66 / :91-96):
//   leaf = findLeaf(root)            :194-201   descend while the node has a child for EVERY legal move, by upperConfidenceStrategy
//   simulate(leaf)                   :238-278   one playout of the biased policy; EVERY state on the way becomes a tree node (:271-276)
//   backpropagate(terminal, scores)  :280-289   every node from the terminal one to the root adds scores[its own upNext()], numEpisodes++
// and afterwards getBestStatesSequence (:108-122).  The tree is therefore a trie of playouts (the root of a 20-move game never has all
// its children, so findLeaf nearly always returns the root), and its best-states walk reaches the terminal depth.
//
// Mapping.  A search is inherently sequential (iteration k+1 reads the statistics iteration k wrote), so the parallel axis is the tree:
// BASELINE config 5 has 32,768 agents replanning at once.  One thread owns one tree: it carries the game state down the tree (nodes
// store no state: 32 bytes each), plays table-driven plies (fast_legal / fast_move, shared with the rollout kernel) and touches global
// memory only for the nodes on its path.  A warp per tree was the first candidate; it leaves 31 lanes idle in every table-driven ply
// and its issue rate, not memory latency, bounds the launch (DESIGN.md §5.2 has the arithmetic and the measurement).
//
// Random sources (as include/hk_abi.h documents them; the CPU checker in tests/ draws the same): tree r has key = seed + r.  Policy index of iteration `it`
// (counted over the life of the tree), ply p of the playout: word 0 of Philox4x32-10(key, (it, 0, p, 0)) through the closed-form
// distribution g.cdf; initial pick of upperConfidenceStrategy (:169): word 0 of Philox4x32-10(key ^ 0x9E3779B97F4A7C15, (picks, 0, 0, 0))
// modulo the child count.  totalValue is float32 and is updated in the reference's order, so trees are BIT-EQUAL to the oracle's.

struct SeqTree {                                  // per-tree header; persists between calls like HierarchicalKartAgent.currentRoot
    hk_game_state root;
    unsigned long long key, picks, iters;
    int n_nodes, status, root_upnext, pad_;
    signed char root_cnt[HK_MAX_KARTS];           // prepared policy-ordered legal lists of karts that start from a non-action bucket
    unsigned char root_order[HK_MAX_KARTS][HK_MAX_ACTIONS];   // (quirk B.6-1: the root's (0, bucket)); -1 = not prepared
};

static_assert(sizeof(hk_mcts_node) == 32, "hk_mcts_node is one 32-byte sector");

__device__ __forceinline__ float seq_uct(const float* __restrict__ logtab, int n_log, int parent_n, float total, int n)   // :162-165
{
    const int ratio = parent_n / n;                                   // integer division (quirk B.6-7); caller has checked n != 0
    const float lg = ratio <= 0 ? -INFINITY : (ratio < n_log ? __ldg(&logtab[ratio]) : (float)log((double)(float)ratio));
    return total / (float)n + 1.0f * lg;
}

// upperConfidenceStrategy (:167-192) over the insertion-ordered child list of `nd`: node index of the chosen child (its record in
// `out`), -2 when a child without episodes makes UCTWeight divide by zero (DivideByZeroException).
__device__ int seq_ucs(const hk_mcts_node* __restrict__ nodes, const hk_mcts_node& nd, int n_children, unsigned long long key,
                       unsigned long long& picks, const float* __restrict__ logtab, int n_log, hk_mcts_node& out)
{
    const int index = (int)(philox_first(key ^ 0x9E3779B97F4A7C15ull, picks, 0u) % (unsigned)n_children);
    picks += 1;
    // one pass: remember the initial pick's weight when the walk reaches it; the loop's strict `>` starts from that weight, so
    // children before it compete with it later — collect the weights first (n_children <= 36)
    float w[HK_MAX_ACTIONS];
    int idx[HK_MAX_ACTIONS];
    int c = nd.first_child;
    for (int j = 0; j < n_children; ++j) {
        const hk_mcts_node ch = nodes[c];
        if (ch.numEpisodes == 0) return -2;
        w[j] = seq_uct(logtab, n_log, nd.numEpisodes, ch.totalValue, ch.numEpisodes);
        idx[j] = c;
        c = ch.next_sibling;
    }
    int best = index;
    float best_w = w[index];
    for (int j = 0; j < n_children; ++j)
        if (w[j] > best_w) { best_w = w[j]; best = j; }
    out = nodes[idx[best]];
    return idx[best];
}

// applies the action with generation index gi for kart np to the carried state (makeMove :420-446)
__device__ __forceinline__ void seq_apply(const DevGame& g, const Tables& tb, hk_game_state& st, int np, int gi, int& lcs_idx)
{
    const int lvl = (g.tables_ok && st.karts[np].player == 0) ? velocity_level(g, st.karts[np].min_velocity, st.karts[np].max_velocity) : -1;
    if (lvl >= 0) {
        fast_move(g, tb, st, np, lvl, gi, st.karts[np].section % g.n_sections, lcs_idx);
    } else {
        make_move(g, st, np, action_of(g, gi));
        lcs_idx = st.lastCompletedSection % g.n_sections;
    }
}

constexpr int SEQ_MAX_PATH = HK_MAX_PLIES + 1;

__global__ void __launch_bounds__(64) seq_search_kernel(const DevGame* __restrict__ gg, SeqTree* __restrict__ trees, hk_mcts_node* __restrict__ slabs,
                                                        int max_nodes, int n_trees, int iterations, unsigned long long seed, int tree_base,
                                                        const hk_game_state* __restrict__ roots, const int* __restrict__ fresh,
                                                        const float* __restrict__ logtab, int n_log,
                                                        hk_game_state* __restrict__ best_out, int* __restrict__ n_best_out,
                                                        int* __restrict__ n_nodes_out, int* __restrict__ status_out, int active_lanes)
{
    __shared__ DevGame g;
    {
        const int* src = reinterpret_cast<const int*>(gg);
        int* dst = reinterpret_cast<int*>(&g);
        for (int i = threadIdx.x; i < (int)(sizeof(DevGame) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    // Only the first `active_lanes` lanes of a warp own a tree (32 in production; the experiment with fewer is recorded at the launch site).
    const int wl = threadIdx.x & 31;
    if (wl >= active_lanes) return;
    const int t = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * active_lanes + wl;
    if (t >= n_trees) return;
    if (fresh && fresh[t] < 0) {                                       // this tree does not search in this call (its outputs stay untouched)
        if (status_out) status_out[t] = 0;
        return;
    }
    const Tables tb(g);
    SeqTree& tr = trees[t];
    hk_mcts_node* nodes = slabs + (size_t)t * max_nodes;
    int status = 0;
    if (!fresh || fresh[t]) {                                          // new KartMCTSNode(state) :52
        tr.root = roots[t];
        tr.key = seed + (unsigned long long)(tree_base + t);
        tr.picks = 0; tr.iters = 0; tr.n_nodes = 1; tr.status = 0;
        const int up = up_next(tr.root);
        tr.root_upnext = up;
        hk_mcts_node r;
        r.child_mask = 0ull; r.totalValue = 0.0f; r.numEpisodes = 0; r.first_child = -1; r.last_child = -1; r.next_sibling = -1;
        r.gen = 255; r.n_legal = 255; r.upnext = (signed char)up; r.pad_ = 0;
        nodes[0] = r;
        for (int k = 0; k < HK_MAX_KARTS; ++k) {
            int cnt = -1;
            if (k < tr.root.n_karts) {
                const int lvl = (g.tables_ok && tr.root.karts[k].player == 0) ? velocity_level(g, tr.root.karts[k].min_velocity, tr.root.karts[k].max_velocity) : -1;
                if (lvl < 0) {
                    unsigned long long keys[HK_MAX_ACTIONS];
                    cnt = legal_moves(g, tr.root, k, keys);
                    for (int c = 0; c < g.n_cand; ++c)
                        if (keys[c] != ~0ull) {                        // keys are distinct: rank = number of smaller keys
                            int rank = 0;
                            for (int j = 0; j < g.n_cand; ++j) rank += keys[j] < keys[c];
                            tr.root_order[k][rank] = (unsigned char)c;
                        }
                }
            }
            tr.root_cnt[k] = (signed char)cnt;
        }
    } else {
        status = tr.status;
    }
    const unsigned long long key = tr.key;
    unsigned long long picks = tr.picks, iters = tr.iters;
    int n_nodes = tr.n_nodes;
    const hk_game_state root = tr.root;
    const int root_lcs_idx = root.lastCompletedSection % g.n_sections;
    int path[SEQ_MAX_PATH];

    for (int it = 0; it < iterations && status == 0; ++it, ++iters) {
        hk_game_state st = root;
        int lcs_idx = root_lcs_idx;
        unsigned moved = 0;
        int node = 0, depth = 0;
        hk_mcts_node nd = nodes[0];
        path[0] = 0;
        // ---- findLeaf :194-201 ----------------------------------------------------------------------------------------------------
        for (;;) {
            if (nd.first_child < 0) break;                             // children.Count > 0
            const int nch = __popcll(nd.child_mask);
            if (nch != (int)nd.n_legal) break;                         // children.Count == state.nextMoves().Count
            hk_mcts_node ch;
            const int c = seq_ucs(nodes, nd, nch, key, picks, logtab, n_log, ch);
            if (c < 0) { status = 2; break; }
            const int np = nd.upnext;
            seq_apply(g, tb, st, np, ch.gen, lcs_idx);
            moved |= 1u << np;
            node = c; nd = ch;
            if (depth + 1 < SEQ_MAX_PATH) path[++depth] = c; else { status = 3; break; }
        }
        if (status) break;
        // ---- simulate :238-278 --------------------------------------------------------------------------------------------------------
        float scores[2 * HK_MAX_KARTS];
        int n_scores = 0;
        bool dirty = false;                                            // nd differs from nodes[node]
        for (unsigned ply = 0;; ++ply) {
            const int np = nd.upnext;
            if (np < 0) { status = 1; break; }                         // ArgumentOutOfRangeException at KartDiscreteGame.cs:326
            const int lvl = (g.tables_ok && st.karts[np].player == 0) ? velocity_level(g, st.karts[np].min_velocity, st.karts[np].max_velocity) : -1;
            int cnt, gi, sidx = 0;
            unsigned long long mask = 0ull;
            const unsigned char* ord = nullptr;
            unsigned long long keys[HK_MAX_ACTIONS];
            int kind;
            if (lvl >= 0) { kind = 1; cnt = fast_legal(g, tb, st, np, lvl, mask, ord, sidx, lcs_idx); }
            else if (!((moved >> np) & 1u) && tr.root_cnt[np] >= 0) { kind = 0; cnt = tr.root_cnt[np]; }
            else { kind = 2; cnt = legal_moves(g, st, np, keys); }
            if (is_over(g, st, cnt, np, scores, n_scores)) break;      // :243-249
            if (nd.n_legal == 255) { nd.n_legal = (unsigned char)cnt; dirty = true; }
            const int index = policy_index(g, cnt, philox_first(key, iters, ply));   // :266-269
            if (kind == 1) gi = __ldg(&ord[nth_set_bit(mask, index)]);
            else if (kind == 0) gi = tr.root_order[np][index];
            else gi = select_kth(keys, g.n_cand, index);
            // leaf.children.ContainsKey(move) :271
            int child;
            hk_mcts_node ch;
            if ((nd.child_mask >> gi) & 1ull) {
                if (dirty) { nodes[node] = nd; dirty = false; }
                child = nd.first_child;
                for (;;) {
                    ch = nodes[child];
                    if (ch.gen == gi) break;
                    child = ch.next_sibling;
                }
                if (kind == 1) fast_move(g, tb, st, np, lvl, gi, sidx, lcs_idx);
                else { make_move(g, st, np, action_of(g, gi)); lcs_idx = st.lastCompletedSection % g.n_sections; }
            } else {                                                   // new KartMCTSNode(state.makeMove(move), leaf) :273
                if (n_nodes >= max_nodes) { status = 3; break; }
                child = n_nodes++;
                if (nd.last_child >= 0) nodes[nd.last_child].next_sibling = child; else nd.first_child = child;
                nd.last_child = child;
                nd.child_mask |= 1ull << gi;
                nodes[node] = nd; dirty = false;
                if (kind == 1) fast_move(g, tb, st, np, lvl, gi, sidx, lcs_idx);
                else { make_move(g, st, np, action_of(g, gi)); lcs_idx = st.lastCompletedSection % g.n_sections; }
                ch.child_mask = 0ull; ch.totalValue = 0.0f; ch.numEpisodes = 0; ch.first_child = -1; ch.last_child = -1; ch.next_sibling = -1;
                ch.gen = (unsigned char)gi; ch.n_legal = 255; ch.upnext = (signed char)up_next(st); ch.pad_ = 0;
                nodes[child] = ch;
            }
            moved |= 1u << np;
            node = child; nd = ch;
            if (depth + 1 < SEQ_MAX_PATH) path[++depth] = child; else { status = 3; break; }
        }
        if (status) break;
        if (dirty) nodes[node] = nd;
        // ---- backpropagate :280-289 ---------------------------------------------------------------------------------------------------
        for (int d = depth; d >= 0; --d) {
            hk_mcts_node* p = &nodes[path[d]];
            const int up = p->upnext;
            float tot = p->totalValue;
            if (up >= 0 && up < n_scores) tot += scores[up];
            p->totalValue = tot;
            p->numEpisodes += 1;
        }
    }
    tr.picks = picks; tr.iters = iters; tr.n_nodes = n_nodes; tr.status = status;

    // ---- getBestStatesSequence :108-122 (it consumes picks like any other upperConfidenceStrategy call; the counter persists) ----------
    int nb = 0;
    if (best_out && (status == 0 || status == 3)) {
        hk_game_state st = root;
        int lcs_idx = root_lcs_idx;
        hk_mcts_node nd = nodes[0];
        while (nd.first_child >= 0) {
            hk_mcts_node ch;
            const int c = seq_ucs(nodes, nd, __popcll(nd.child_mask), key, picks, logtab, n_log, ch);
            if (c < 0) break;                                          // catch (DivideByZeroException) { } :120
            seq_apply(g, tb, st, nd.upnext, ch.gen, lcs_idx);
            nd = ch;
            bool all = true;
            for (int i = 0; i < st.n_karts; ++i) all &= st.karts[i].section == st.lastCompletedSection;
            if (all) { if (nb < HK_MCTS_MAX_SEQ) best_out[(size_t)t * HK_MCTS_MAX_SEQ + nb] = st; ++nb; }
        }
        tr.picks = picks;
    }
    if (n_best_out) n_best_out[t] = nb < HK_MCTS_MAX_SEQ ? nb : HK_MCTS_MAX_SEQ;
    if (n_nodes_out) n_nodes_out[t] = n_nodes;
    if (status_out) status_out[t] = status;
}
