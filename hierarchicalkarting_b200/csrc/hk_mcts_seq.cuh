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
    unsigned long long iters0;                    // iters when the current call started (the fast path's playouts are numbered from it)
    int n_nodes, status, root_upnext;
    int aux_ok;                                   // the tree's prefix table (see seq_insert_kernel) describes its top levels
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

// One ply of simulate (:243-269) on the carried state: the legal moves of kart np in policy order, isOver, the index draw, the move.
// Returns false when the state is terminal (scores filled), true after applying the chosen move (cnt = nextMoves().Count, gi = the move).
__device__ __forceinline__ bool seq_ply(const DevGame& g, const Tables& tb, const SeqTree& tr, hk_game_state& st, int& lcs_idx, unsigned& moved, int np,
                                        unsigned long long key, unsigned long long iters, unsigned ply, float* scores, int& n_scores, int& cnt, int& gi)
{
    const int lvl = (g.tables_ok && st.karts[np].player == 0) ? velocity_level(g, st.karts[np].min_velocity, st.karts[np].max_velocity) : -1;
    int sidx = 0, kind;
    unsigned long long mask = 0ull;
    const unsigned char* ord = nullptr;
    unsigned long long keys[HK_MAX_ACTIONS];
    if (lvl >= 0) { kind = 1; cnt = fast_legal(g, tb, st, np, lvl, mask, ord, sidx, lcs_idx); }
    else if (!((moved >> np) & 1u) && tr.root_cnt[np] >= 0) { kind = 0; cnt = tr.root_cnt[np]; }
    else { kind = 2; cnt = legal_moves(g, st, np, keys); }
    if (is_over(g, st, cnt, np, scores, n_scores)) return false;       // :243-249
    const int index = policy_index(g, cnt, philox_first(key, iters, ply));   // :266-269
    if (kind == 1) gi = __ldg(&ord[nth_set_bit(mask, index)]);
    else if (kind == 0) gi = tr.root_order[np][index];
    else gi = select_kth(keys, g.n_cand, index);
    if (kind == 1) fast_move(g, tb, st, np, lvl, gi, sidx, lcs_idx);
    else { make_move(g, st, np, action_of(g, gi)); lcs_idx = st.lastCompletedSection % g.n_sections; }
    moved |= 1u << np;
    return true;
}

// leaf.children.ContainsKey(move) ? leaf.children[move] : new KartMCTSNode(...) (:271-276) on the node records.  `nd` is the register copy of
// nodes[node] (dirty: differs from memory); cnt = nextMoves().Count of the node's state, up_after = upNext() of the state after the move.
// Returns the child (its record in nd, node updated), or -1 when the slab is full.
__device__ __forceinline__ int seq_descend(hk_mcts_node* __restrict__ nodes, int& node, hk_mcts_node& nd, bool& dirty, int cnt, int gi, int up_after,
                                           int& n_nodes, int max_nodes)
{
    if (nd.n_legal == 255) { nd.n_legal = (unsigned char)cnt; dirty = true; }
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
    } else {                                                           // new KartMCTSNode(state.makeMove(move), leaf) :273
        if (n_nodes >= max_nodes) return -1;
        child = n_nodes++;
        if (nd.last_child >= 0) nodes[nd.last_child].next_sibling = child; else nd.first_child = child;
        nd.last_child = child;
        nd.child_mask |= 1ull << gi;
        nodes[node] = nd; dirty = false;
        ch.child_mask = 0ull; ch.totalValue = 0.0f; ch.numEpisodes = 0; ch.first_child = -1; ch.last_child = -1; ch.next_sibling = -1;
        ch.gen = (unsigned char)gi; ch.n_legal = 255; ch.upnext = (signed char)up_after; ch.pad_ = 0;
        nodes[child] = ch;
    }
    node = child; nd = ch;
    return child;
}

__device__ __forceinline__ void seq_backprop(hk_mcts_node* __restrict__ nodes, const int* path, int depth, const float* scores, int n_scores)   // :280-289
{
    for (int d = depth; d >= 0; --d) {
        hk_mcts_node* p = &nodes[path[d]];
        const int up = p->upnext;
        float tot = p->totalValue;
        if (up >= 0 && up < n_scores) tot += scores[up];
        p->totalValue = tot;
        p->numEpisodes += 1;
    }
}

#define SEQ_LOAD_GAME()                                                                                      \
    __shared__ DevGame g;                                                                                    \
    {                                                                                                        \
        const int* src = reinterpret_cast<const int*>(gg);                                                   \
        int* dst = reinterpret_cast<int*>(&g);                                                               \
        for (int i_ = threadIdx.x; i_ < (int)(sizeof(DevGame) / 4); i_ += blockDim.x) dst[i_] = src[i_];     \
    }                                                                                                        \
    __syncthreads();

// The general sequential kernel, one thread per tree, in three optional phases: INIT (new KartMCTSNode(state) for fresh trees), `iterations`
// (or remaining[t], if given) full iterations — findLeaf by upperConfidenceStrategy, simulate, backpropagate —, BEST (getBestStatesSequence
// and the outputs).  The fast path below (seq_playouts_kernel + seq_insert_kernel) takes the iterations whose findLeaf returns the root;
// this kernel takes the rest, and is the whole search when the fast path is switched off.
__global__ void __launch_bounds__(64) seq_search_kernel(const DevGame* __restrict__ gg, SeqTree* __restrict__ trees, hk_mcts_node* __restrict__ slabs,
                                                        int max_nodes, int n_trees, int iterations, const int* __restrict__ remaining,
                                                        unsigned long long seed, int tree_base,
                                                        const hk_game_state* __restrict__ roots, const int* __restrict__ fresh,
                                                        const float* __restrict__ logtab, int n_log,
                                                        hk_game_state* __restrict__ best_out, int* __restrict__ n_best_out,
                                                        int* __restrict__ n_nodes_out, int* __restrict__ status_out, int active_lanes,
                                                        bool do_init, bool do_best)
{
    SEQ_LOAD_GAME();
    // Only the first `active_lanes` lanes of a warp own a tree (32 in production; the experiment with fewer is recorded at the launch site).
    const int wl = threadIdx.x & 31;
    if (wl >= active_lanes) return;
    const int t = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * active_lanes + wl;
    if (t >= n_trees) return;
    if (fresh && fresh[t] < 0) {                                       // this tree does not search in this call (its outputs stay untouched)
        if (status_out && do_best) status_out[t] = 0;
        return;
    }
    const Tables tb(g);
    SeqTree& tr = trees[t];
    hk_mcts_node* nodes = slabs + (size_t)t * max_nodes;
    int status = 0;
    if (do_init && (!fresh || fresh[t])) {                             // new KartMCTSNode(state) :52
        tr.root = roots[t];
        tr.key = seed + (unsigned long long)(tree_base + t);
        tr.picks = 0; tr.iters = 0; tr.n_nodes = 1; tr.status = 0; tr.aux_ok = 0;
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
    if (do_init) tr.iters0 = tr.iters;
    const unsigned long long key = tr.key;
    unsigned long long picks = tr.picks, iters = tr.iters;
    int n_nodes = tr.n_nodes;
    const hk_game_state root = tr.root;
    const int root_lcs_idx = root.lastCompletedSection % g.n_sections;
    int path[SEQ_MAX_PATH];
    const int my_iterations = remaining ? remaining[t] : iterations;

    for (int it = 0; it < my_iterations && status == 0; ++it, ++iters) {
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
            int cnt, gi;
            if (!seq_ply(g, tb, tr, st, lcs_idx, moved, np, key, iters, ply, scores, n_scores, cnt, gi)) break;
            if (seq_descend(nodes, node, nd, dirty, cnt, gi, up_next(st), n_nodes, max_nodes) < 0) { status = 3; break; }
            if (depth + 1 < SEQ_MAX_PATH) path[++depth] = node; else { status = 3; break; }
        }
        if (status) break;
        if (dirty) nodes[node] = nd;
        seq_backprop(nodes, path, depth, scores, n_scores);
    }
    tr.picks = picks; tr.iters = iters; tr.n_nodes = n_nodes; tr.status = status;
    if (my_iterations > 0) tr.aux_ok = 0;                              // nodes were created without the prefix table
    if (!do_best) return;

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

// ---- fast path: the playouts of many iterations at once, then their insertion in iteration order -------------------------------------------
// As long as the root does not have a child for every legal move, findLeaf (:194-201) returns the root, and the playout of iteration `it` —
// its moves and terminal scores — is a function of the root state and the Philox draws (it, ply) only: NOT of the tree, which merely records
// it.  So the playouts of a whole chunk of iterations are computed in parallel, one thread per (tree, iteration) (seq_playouts_kernel: the
// machine is full — 32,768 trees alone are only 7 warps per SM), and one thread per tree then replays what the sequential loop would have
// done to the node records, in order: look up / create the child of every ply, set nextMoves().Count where a node is processed for the first
// time, backpropagate (seq_insert_kernel: no game arithmetic, ~30 instructions per ply).  The moment a root IS fully expanded (few legal
// moves), the tree leaves the fast path: remaining[t] counts the iterations the general kernel above still has to run for it.  The node
// records, counters and float32 sums come out bit-equal to the sequential kernel's (tests/test_mcts_seq_gpu.py runs both).
// Measured and rejected (round 2, 32,768 trees x 512 iterations, insertion 14.0 ms): (a) the insertion as ONE flat loop per lane (one node load
// per trip; hop / descend / create / next iteration as a per-lane state machine, so that the 32 trees of a warp do not wait for each other at
// every ply) — 19.9 ms: the lanes' different states serialise, 6x the warp-instructions (ncu: 56 k against 9.6 k per tree and chunk) for a
// third of the long-scoreboard stalls; (b) the search of a planning event on its own stream beside the LQNG steps that pass before its result
// lands (hk_race_run_planned, apply_delay 45) — 78.6-86 ms against 77.2 ms per 200 steps, at the lowest stream priority too.
//   record of a playout: word 0 = plies | n_scores << 8 | error << 16; words 1..8 = score bits; word 9 + p = gi | nextMoves().Count << 8 |
//   (upNext() after the move & 0xff) << 16; the head is padded to 12 words and the record to a multiple of 4 (16-byte aligned, read as uint4)
constexpr int SEQ_REC_HEAD = 12;                                    // head, 8 score words, 3 unused: the ply words start 16-byte aligned
static_assert(SEQ_REC_HEAD >= 1 + 2 * HK_MAX_KARTS && SEQ_REC_HEAD % 4 == 0, "record head");
__host__ __device__ constexpr size_t seq_rec_words(int cap) { return (size_t)SEQ_REC_HEAD + (size_t)((cap + 3) & ~3); }   // records are read as uint4

// Once both karts hold action velocity buckets (after their first moves) the playout continues in the packed, register-resident form of the
// rollout kernel (rollout_packed2<true>, which also leaves the record words): 10.5 -> 7.4 ms per 32,768 trees x 512 iterations.  Resident blocks
// per SM 3 / 4 / 5 / 6 / 8 (109 / 119 / 96 / 80 / 64 registers): 7.7 / 7.7 / 7.1 / 6.8 / 7.8 ms.
// NKMAX 2: games of one or two karts; 4: three or four karts with their own packed form (rollout_packedN) — 13.1 -> 11.9 ms per 32,768 Duos
// trees x 256 iterations at 5 blocks per SM (12.3 ms at 4): the select chains and the 4-way upNext keep a ply ~1.7x as expensive as with two karts.
template <int NKMAX>
__global__ void __launch_bounds__(128, NKMAX == 2 ? 6 : 5) seq_playouts_kernel(const DevGame* __restrict__ gg, const SeqTree* __restrict__ trees, int n_trees,
                                                           const int* __restrict__ fresh, const int* __restrict__ remaining, int chunk, int base,
                                                           int cap, unsigned* __restrict__ recs)
{
    SEQ_LOAD_GAME();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_trees * chunk) return;
    const int t = (int)(idx / chunk), i = (int)(idx % chunk);
    if ((fresh && fresh[t] < 0) || remaining[t] >= 0) return;          // not searching / already on the general path
    const Tables tb(g);
    const SeqTree& tr = trees[t];
    if (tr.status) return;
    unsigned* rec = recs + (size_t)idx * seq_rec_words(cap);
    hk_game_state st = tr.root;
    int lcs_idx = st.lastCompletedSection % g.n_sections;
    unsigned moved = 0;
    const unsigned long long it = tr.iters0 + (unsigned long long)(base + i);   // not tr.iters: the insertion of the previous chunk may still be running
    float scores[2 * HK_MAX_KARTS];
    int n_scores = 0, np = tr.root_upnext, err = 0;
    unsigned ply = 0;
    for (;; ++ply) {
        if (np < 0) { err = 1; break; }
        if (NKMAX == 2 && g.tables_ok && st.n_karts == 2) {            // both karts at action buckets: the rest runs packed in registers
            K2 k0, k1;
            if (pack_k2(g, st.karts[0], k0) && pack_k2(g, st.karts[1], k1)) {
                int first_gi = -1;
                const int r = rollout_packed2<true>(g, tb, st, k0, k1, tr.key, it, scores, n_scores, first_gi, (int)ply, rec + SEQ_REC_HEAD, cap);
                if (r < 0) err = r == -1 ? 1 : 3; else ply = (unsigned)r;
                break;
            }
        }
        if (NKMAX == 4 && g.tables_ok && st.n_karts >= 3) {            // every kart at an action bucket
            K2 k0, k1, k2, k3 = K2{0, 0, 0, 0};
            if (pack_k2(g, st.karts[0], k0) && pack_k2(g, st.karts[1], k1) && pack_k2(g, st.karts[2], k2) &&
                (st.n_karts == 3 || pack_k2(g, st.karts[3], k3))) {
                const int r = st.n_karts == 3 ? rollout_packedN<3, true>(g, tb, st, k0, k1, k2, k3, tr.key, it, scores, n_scores, (int)ply, rec + SEQ_REC_HEAD, cap)
                                              : rollout_packedN<4, true>(g, tb, st, k0, k1, k2, k3, tr.key, it, scores, n_scores, (int)ply, rec + SEQ_REC_HEAD, cap);
                if (r < 0) err = r == -1 ? 1 : 3; else ply = (unsigned)r;
                break;
            }
        }
        int cnt, gi;
        if (!seq_ply(g, tb, tr, st, lcs_idx, moved, np, tr.key, it, ply, scores, n_scores, cnt, gi)) break;
        if ((int)ply >= cap) { err = 3; break; }
        np = up_next(st);
        rec[SEQ_REC_HEAD + ply] = (unsigned)gi | ((unsigned)cnt << 8) | (((unsigned)np & 0xffu) << 16);
    }
    rec[0] = ply | ((unsigned)n_scores << 8) | ((unsigned)err << 16);
    for (int k = 0; k < 2 * HK_MAX_KARTS; ++k) rec[1 + k] = k < n_scores ? __float_as_uint(scores[k]) : 0u;
}

// Prefix table (auxiliary, per tree, not part of the tree): aux[(gi_0)], aux[nc + gi_0 nc + gi_1], aux[nc + nc^2 + (gi_0 nc + gi_1) nc + gi_2] = index of
// the node reached from the root by the moves gi_0 (, gi_1 (, gi_2)), 0 = not there.  The child lists of the top levels are the long ones (3.6 /
// 3.7 / 3.1 sibling hops per visit at depth 0 / 1 / 2 of a 512-iteration tree, 1.7 and less below; up to 8-12 for the unluckiest of a warp's 32
// lanes, and the warp waits for that one), and every hop is a dependent load.  With the table the three entries are loaded at once — their addresses
// depend on the playout record only — then the three nodes at once: two dependent stages instead of ~10 (~24 for the warp).  The table is kept by
// this kernel alone: aux_ok is cleared when anything else creates nodes in the tree (the general kernel, the checked insertion), and the
// lists are walked again from then on.
__global__ void seq_aux_clear_kernel(SeqTree* __restrict__ trees, int* __restrict__ aux, long long aux_stride, int n_trees, const int* __restrict__ fresh)
{
    const int t = blockIdx.y;                                         // one row of blocks per tree: kept trees leave at once
    if (fresh && fresh[t] <= 0) return;                               // kept tree (or not searching): its table stays
    int* a = aux + (size_t)t * aux_stride;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < aux_stride; e += (long long)gridDim.x * blockDim.x) a[e] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) trees[t].aux_ok = 1;
}

__global__ void __launch_bounds__(64) seq_insert_kernel(SeqTree* __restrict__ trees, hk_mcts_node* __restrict__ slabs, int max_nodes, int n_trees,
                                                        const int* __restrict__ fresh, int* __restrict__ remaining, int chunk, int count, int todo_after,
                                                        int cap, const unsigned* __restrict__ recs, int* __restrict__ aux_all, long long aux_stride,
                                                        int aux_levels, int nc)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trees) return;
    if ((fresh && fresh[t] < 0) || remaining[t] >= 0) return;
    SeqTree& tr = trees[t];
    hk_mcts_node* nodes = slabs + (size_t)t * max_nodes;
    int* aux = aux_all ? aux_all + (size_t)t * aux_stride : nullptr;
    bool aux_ok = aux && tr.aux_ok;
    int status = tr.status, n_nodes = tr.n_nodes;
    unsigned long long iters = tr.iters;
    int path[SEQ_MAX_PATH];
    int i = 0;
    // What the loop waits for is memory, so everything whose address does not depend on the tree is fetched ahead: the first 16 words of the
    // NEXT record (head, scores, ply words 0..3: four 16-byte loads) while this one is inserted, the ply words of this record four at a time
    // one group ahead, and the next record's table entries as soon as this iteration is through its own top plies (the only place where
    // the table is written).  The root's record stays in registers (its memory copy is kept current).
    const int rq = (int)(seq_rec_words(cap) / 4);                      // uint4 per record
    const uint4* rq0 = reinterpret_cast<const uint4*>(recs) + (size_t)t * chunk * rq;
    uint4 nA = make_uint4(0, 0, 0, 0), nB = nA, nC = nA, nP = nA;
    if (count > 0) { nA = rq0[0]; nB = rq0[1]; nC = rq0[2]; nP = rq0[3]; }
    hk_mcts_node root = nodes[0];
    int npre0 = 0, npre1 = 0, npre2 = 0;
    bool npre_valid = false;
    for (; i < count && status == 0; ++i, ++iters) {
        if (root.first_child >= 0 && __popcll(root.child_mask) == (int)root.n_legal) break;   // findLeaf would descend: the general kernel takes over
        const uint4* rcur = rq0 + (size_t)i * rq;
        const uint4 A = nA, B = nB, C = nC, P = nP;
        if (i + 1 < count) { const uint4* rn = rcur + rq; nA = rn[0]; nB = rn[1]; nC = rn[2]; nP = rn[3]; }
        const unsigned head = A.x;
        const int len = head & 0xff, n_scores = (head >> 8) & 0xff, err = head >> 16;
        if (err) { status = err; break; }
        const float s0 = __uint_as_float(A.y), s1 = __uint_as_float(A.z), s2 = __uint_as_float(A.w), s3 = __uint_as_float(B.x),
                    s4 = __uint_as_float(B.y), s5 = __uint_as_float(B.z), s6 = __uint_as_float(B.w), s7 = __uint_as_float(C.x);
        static_assert(2 * HK_MAX_KARTS == 8, "eight score words");
        int node = 0;
        if (n_nodes + len <= max_nodes) {
            // The terminal scores are known before the descent, so backpropagate (:280-289) is folded into it: a node on the path gets its
            // totalValue += scores[upNext()] and numEpisodes += 1 while its record is in registers — a walked node is written back once, a
            // new node is written once with its first episode already in — instead of a second pass that loads every node of the path
            // again.  Every node still sees its additions in iteration order.  (Not when the slab might fill up in this iteration: the
            // sequential loop then stops before backpropagating.)
            auto credit = [&](hk_mcts_node& x) {
                const int up = x.upnext;
                float sc = s0;                                         // a select chain: an indexed array would live in local memory
                sc = up == 1 ? s1 : sc; sc = up == 2 ? s2 : sc; sc = up == 3 ? s3 : sc; sc = up == 4 ? s4 : sc;
                sc = up == 5 ? s5 : sc; sc = up == 6 ? s6 : sc; sc = up == 7 ? s7 : sc;
                if (up >= 0 && up < n_scores) x.totalValue += sc;
                x.numEpisodes += 1;
            };
            // prefix table: entries of the first plies (already here when the previous iteration fetched them), then their nodes
            int slot0 = -1, slot1 = -1, slot2 = -1, pre0 = 0, pre1 = 0, pre2 = 0;
            hk_mcts_node pn0, pn1, pn2;
            if (aux_ok) {
                const int g0 = P.x & 0xff, g1 = P.y & 0xff, g2 = P.z & 0xff;      // words past len: unused
                if (len > 0) slot0 = g0;
                if (len > 1 && aux_levels > 1) slot1 = nc + g0 * nc + g1;
                if (len > 2 && aux_levels > 2) slot2 = nc + nc * nc + (g0 * nc + g1) * nc + g2;
                if (npre_valid) { pre0 = npre0; pre1 = npre1; pre2 = npre2; }
                else {
                    if (slot0 >= 0) pre0 = aux[slot0];
                    if (slot1 >= 0) pre1 = aux[slot1];
                    if (slot2 >= 0) pre2 = aux[slot2];
                }
                if (pre0 > 0) pn0 = nodes[pre0];
                if (pre1 > 0) pn1 = nodes[pre1];
                if (pre2 > 0) pn2 = nodes[pre2];
            }
            npre_valid = false;
            auto fetch_next_entries = [&]() {                          // after this iteration's last table write
                if (!aux_ok || i + 1 >= count) return;
                const int nlen = nA.x & 0xff;
                const int h0 = min((int)(nP.x & 0xff), nc - 1), h1 = min((int)(nP.y & 0xff), nc - 1), h2 = min((int)(nP.z & 0xff), nc - 1);
                npre0 = nlen > 0 ? aux[h0] : 0;
                npre1 = (nlen > 1 && aux_levels > 1) ? aux[nc + h0 * nc + h1] : 0;
                npre2 = (nlen > 2 && aux_levels > 2) ? aux[nc + nc * nc + (h0 * nc + h1) * nc + h2] : 0;
                npre_valid = true;
            };
            hk_mcts_node nd = root;
            credit(nd);
            if (len == 0) { nodes[0] = nd; root = nd; fetch_next_entries(); continue; }   // isOver() at the root: it collects the episode
            uint4 cur4 = P, next4 = P;
            if (len > 4) next4 = rcur[4];
            for (int ply = 0; ply < len; ++ply) {
                if ((ply & 3) == 0 && ply > 0) {
                    cur4 = next4;
                    if (ply + 4 < len) next4 = rcur[3 + (ply >> 2) + 1];
                }
                const int q = ply & 3;
                const unsigned w = q == 0 ? cur4.x : q == 1 ? cur4.y : q == 2 ? cur4.z : cur4.w;
                const int gi = w & 0xff;
                if (nd.n_legal == 255) nd.n_legal = (unsigned char)((w >> 8) & 0xff);
                const int slot = ply == 0 ? slot0 : ply == 1 ? slot1 : ply == 2 ? slot2 : -1;
                int child;
                hk_mcts_node ch;
                if ((nd.child_mask >> gi) & 1ull) {
                    nodes[node] = nd;                                  // statistics (and possibly nextMoves().Count) changed
                    if (node == 0) root = nd;
                    const int pre = ply == 0 ? pre0 : ply == 1 ? pre1 : ply == 2 ? pre2 : 0;
                    if (slot >= 0 && pre > 0) {
                        child = pre;
                        ch = ply == 0 ? pn0 : ply == 1 ? pn1 : pn2;
                    } else {
                        child = nd.first_child;
                        for (;;) {
                            ch = nodes[child];
                            if (ch.gen == gi) break;
                            child = ch.next_sibling;
                        }
                    }
                } else {
                    child = n_nodes++;
                    if (nd.last_child >= 0) nodes[nd.last_child].next_sibling = child; else nd.first_child = child;
                    nd.last_child = child;
                    nd.child_mask |= 1ull << gi;
                    nodes[node] = nd;
                    if (node == 0) root = nd;
                    ch.child_mask = 0ull; ch.totalValue = 0.0f; ch.numEpisodes = 0; ch.first_child = -1; ch.last_child = -1; ch.next_sibling = -1;
                    ch.gen = (unsigned char)gi; ch.n_legal = 255; ch.upnext = (signed char)((w >> 16) & 0xff); ch.pad_ = 0;
                    if (slot >= 0) aux[slot] = child;
                }
                credit(ch);
                node = child; nd = ch;
                if (ply == 2 || ply + 1 == len) { if (ply <= 2) fetch_next_entries(); }   // once: at ply 2, or at the last ply of a shorter playout
            }
            nodes[node] = nd;
            continue;
        }
        aux_ok = false;                                                // the checked form below does not keep the table
        npre_valid = false;
        const unsigned* rec = reinterpret_cast<const unsigned*>(rcur);
        const float scores[2 * HK_MAX_KARTS] = {s0, s1, s2, s3, s4, s5, s6, s7};
        hk_mcts_node nd = root;
        int depth = 0;
        bool dirty = false;
        path[0] = 0;
        for (int ply = 0; ply < len; ++ply) {
            const unsigned w = rec[SEQ_REC_HEAD + ply];
            if (seq_descend(nodes, node, nd, dirty, (w >> 8) & 0xff, w & 0xff, (int)(signed char)((w >> 16) & 0xff), n_nodes, max_nodes) < 0) { status = 3; break; }
            path[++depth] = node;                                      // len <= cap <= HK_MAX_PLIES
        }
        if (status) break;
        if (dirty) nodes[node] = nd;
        seq_backprop(nodes, path, depth, scores, n_scores);
        root = nodes[0];
    }
    tr.iters = iters; tr.n_nodes = n_nodes; tr.status = status;
    if (aux && !aux_ok) tr.aux_ok = 0;
    if (status == 0 && i < count) remaining[t] = (count - i) + todo_after;            // the rest of this call's iterations, sequentially
}

// root children of every tree in insertion order (generation index, numEpisodes, totalValue; -1 / 0 / 0 past the last child):
// what hk_mcts_search_seq_batch reports beside the best states
__global__ void seq_root_children_kernel(const hk_mcts_node* __restrict__ slabs, int max_nodes, int n_trees, int* __restrict__ gen,
                                         int* __restrict__ episodes, float* __restrict__ values)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_trees) return;
    const hk_mcts_node* nodes = slabs + (size_t)t * max_nodes;
    int j = 0;
    for (int c = nodes[0].first_child; c >= 0 && j < HK_MAX_ACTIONS; ++j) {
        const hk_mcts_node ch = nodes[c];
        gen[(size_t)t * HK_MAX_ACTIONS + j] = ch.gen;
        episodes[(size_t)t * HK_MAX_ACTIONS + j] = ch.numEpisodes;
        values[(size_t)t * HK_MAX_ACTIONS + j] = ch.totalValue;
        c = ch.next_sibling;
    }
    for (; j < HK_MAX_ACTIONS; ++j) {
        gen[(size_t)t * HK_MAX_ACTIONS + j] = -1;
        episodes[(size_t)t * HK_MAX_ACTIONS + j] = 0;
        values[(size_t)t * HK_MAX_ACTIONS + j] = 0.0f;
    }
}
