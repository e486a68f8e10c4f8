/*
 * hk_oracle_mcts.c — CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the reference's tree search AS ITS CALLERS RUN IT — constructSearchTree with parallel == false
 * (HierarchicalKartAgent.cs:250,271 never pass `parallel`):
 *   KartMCTSNode                                  Assets/Karting/Scripts/AI/MCTS/KartMCTS.cs:18-38
 *   KartMCTS.constructSearchTree (state / root)   ...:50-78, 80-106   (the !parallel branch :61-66, :91-96)
 *   KartMCTS.findLeaf                             ...:194-201
 *   KartMCTS.UCTWeight / upperConfidenceStrategy  ...:162-165, 167-192
 *   KartMCTS.simulate                             ...:238-278  (EVERY state of the playout becomes a tree node, :271-276)
 *   KartMCTS.backpropagate                        ...:280-289  (from the terminal node of the playout)
 *   KartMCTS.getBestStatesSequence                ...:108-122
 * over the game primitives of hk_oracle_game.c.  The reference's loop is wall-clock budgeted (:55); here the budget is an
 * iteration count.  A tree object survives between calls like HierarchicalKartAgent.currentRoot (:265-283).
 *
 * Random sources.  mode 0 (bit-reproducible, shared with the CUDA library): the policy index of iteration `it`, ply `p` of the
 * playout is word 0 of Philox4x32-10(key, counter = (it, 0, p, 0)) through the closed-form CDF of hk_oracle_policy_cdf; the
 * random initial pick of upperConfidenceStrategy (:169) is word 0 of Philox4x32-10(key ^ 0x9E3779B97F4A7C15, counter = (number of
 * picks so far, 0, 0, 0)) modulo the child count.  `it` and the pick counter run on across calls on the same tree.
 * mode 1 (the reference's own procedures, for distribution tests): random.Next(n) = floor(u * n), the truncated Gaussian of
 * NextGaussian :218-236 with polar Box-Muller N(0,1), all from one xorshift64* state.
 *
 * PARITY UNPINNED by the reference (it has no tests and cannot be compiled here).  A second, independently written restatement
 * (oracle/np_mcts_seq.py, Python objects and dictionaries as in the C# text) is compared with this file in tests/.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "hk_oracle.h"

typedef struct {
    hk_game_state state;        /* KartMCTSNode.state */
    int parent;                 /* KartMCTSNode.parent (-1 = null) */
    int gen;                    /* generation index of the action that created the node (the Dictionary key) */
    float totalValue;           /* :23 */
    int numEpisodes;            /* :24 */
    int n_children, cap_children;
    int* children;              /* Dictionary<DiscreteKartAction, KartMCTSNode> in insertion order (what Keys / foreach enumerate) */
} onode;

struct hk_oracle_tree {
    const hk_oracle_game* g;
    onode* nodes;
    int n_nodes, cap_nodes;
    uint64_t key;               /* optional: the Philox key the owner of the tree searches it with */
    uint64_t picks;             /* upperConfidenceStrategy calls so far */
    uint64_t iters;             /* search iterations so far */
    long long childrenAsRoot;   /* root.childrenAsRoot (:64) */
    uint32_t cdfs[HK_MAX_ACTIONS + 1][HK_MAX_ACTIONS];
};

static int new_node(hk_oracle_tree* t, const hk_game_state* st, int parent, int gen)
{
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes = t->cap_nodes ? 2 * t->cap_nodes : 256;
        t->nodes = (onode*)realloc(t->nodes, sizeof(onode) * (size_t)t->cap_nodes);
    }
    onode* n = &t->nodes[t->n_nodes];
    memset(n, 0, sizeof(*n));
    n->state = *st; n->parent = parent; n->gen = gen;
    return t->n_nodes++;
}

hk_oracle_tree* hk_oracle_tree_create(const hk_oracle_game* g, const hk_game_state* root)
{
    hk_oracle_tree* t = (hk_oracle_tree*)calloc(1, sizeof(*t));
    t->g = g;
    for (int c = 1; c <= HK_MAX_ACTIONS; ++c) hk_oracle_policy_cdf(c, t->cdfs[c]);
    new_node(t, root, -1, -1);                                        /* new KartMCTSNode(state) :52 */
    return t;
}

void hk_oracle_tree_destroy(hk_oracle_tree* t)
{
    if (!t) return;
    for (int i = 0; i < t->n_nodes; ++i) free(t->nodes[i].children);
    free(t->nodes); free(t);
}

static double u01(uint64_t* s)
{
    uint64_t x = *s; x ^= x >> 12; x ^= x << 25; x ^= x >> 27; *s = x;
    return (double)((x * 0x2545F4914F6CDD1DULL) >> 11) * (1.0 / 9007199254740992.0);
}

/* UCTWeight :162-165.  Returns 0 and sets *dz when the integer division divides by zero (DivideByZeroException). */
static float uct_weight(const hk_oracle_tree* t, int node, int* dz)
{
    const onode* n = &t->nodes[node];
    const onode* p = &t->nodes[n->parent];
    if (n->numEpisodes == 0) { *dz = 1; return 0.0f; }
    int ratio = p->numEpisodes / n->numEpisodes;                      /* int / int */
    float lg = (float)log((double)(float)ratio);                      /* Mathf.Log(float) = (float)Math.Log((double)f) */
    return (n->totalValue / (float)n->numEpisodes) + 1.0f * lg;       /* Mathf.Sqrt(1.0f) == 1 */
}

/* upperConfidenceStrategy :167-192: position (in insertion order) of the chosen child, -2 on DivideByZeroException */
static int ucs(hk_oracle_tree* t, int node, int mode, uint64_t key, uint64_t* rng)
{
    const onode* n = &t->nodes[node];
    int index;
    if (mode == 0) {
        uint32_t r[4];
        hk_oracle_philox4x32_10(key ^ 0x9E3779B97F4A7C15ULL, (uint32_t)t->picks, (uint32_t)(t->picks >> 32), 0u, 0u, r);
        index = (int)(r[0] % (uint32_t)n->n_children);
    } else {
        index = (int)(u01(rng) * n->n_children);                      /* random.Next(children.Count) :169 */
    }
    t->picks += 1;
    int dz = 0;
    int best = index;
    float best_uct = uct_weight(t, n->children[best], &dz);
    if (dz) return -2;
    for (int j = 0; j < n->n_children; ++j) {                         /* foreach (var item in node.children) :177 */
        float w = uct_weight(t, n->children[j], &dz);
        if (dz) return -2;
        if (w > best_uct) { best_uct = w; best = j; }
    }
    return best;
}

/* one iteration of the while loop :55-73 with parallel == false.  0 ok, -1 upNext() == -1, -2 DivideByZero in findLeaf */
static int iterate(hk_oracle_tree* t, int mode, uint64_t key, uint64_t* rng)
{
    const hk_oracle_game* g = t->g;
    /* findLeaf :194-201 */
    int leaf = 0;
    for (;;) {
        const onode* n = &t->nodes[leaf];
        if (n->n_children == 0) break;
        int cnt = hk_oracle_next_moves(g, &n->state, 0, 0);
        if (cnt < 0) return -1;
        if (n->n_children != cnt) break;
        int j = ucs(t, leaf, mode, key, rng);
        if (j < 0) return j;
        leaf = t->nodes[leaf].children[j];
    }
    /* simulate :238-278 */
    float scores[2 * HK_MAX_KARTS];
    int n_scores = 0, ply = 0, new_states = 0;
    for (;;) {
        hk_game_state st = t->nodes[leaf].state;
        int over = hk_oracle_is_over(g, &st, scores, &n_scores);     /* :243 */
        if (over < 0) return -1;
        if (over) break;                                              /* :246-249 */
        hk_action mv[HK_MAX_ACTIONS]; int gi[HK_MAX_ACTIONS];
        int cnt = hk_oracle_policy_moves(g, &st, mv, gi);             /* :256 */
        int index;
        if (mode == 0) {
            uint32_t r[4];
            hk_oracle_philox4x32_10(key, (uint32_t)t->iters, (uint32_t)(t->iters >> 32), (uint32_t)ply, 0u, r);
            index = hk_oracle_policy_index(cnt, t->cdfs[cnt], r[0]);
        } else {
            index = hk_oracle_reference_policy_index(cnt, rng);       /* :266-269 */
        }
        int child = -1;
        for (int j = 0; j < t->nodes[leaf].n_children; ++j)          /* leaf.children.ContainsKey(move) :271 */
            if (t->nodes[t->nodes[leaf].children[j]].gen == gi[index]) { child = t->nodes[leaf].children[j]; break; }
        if (child < 0) {
            hk_game_state ns = hk_oracle_make_move(g, &st, mv[index], 0);
            child = new_node(t, &ns, leaf, gi[index]);                /* :273 */
            onode* l = &t->nodes[leaf];
            if (l->n_children == l->cap_children) {
                l->cap_children = l->cap_children ? 2 * l->cap_children : 4;
                l->children = (int*)realloc(l->children, sizeof(int) * (size_t)l->cap_children);
            }
            l->children[l->n_children++] = child;
            new_states += 1;                                          /* :274 */
        }
        leaf = child;                                                 /* :276 */
        ++ply;
    }
    t->childrenAsRoot += new_states;                                  /* :64 */
    /* backpropagate :280-289 */
    for (int node = leaf; node >= 0; node = t->nodes[node].parent) {
        int up = hk_oracle_up_next(g, &t->nodes[node].state);
        if (up >= 0 && up < n_scores) t->nodes[node].totalValue += scores[up];   /* result[-1] / past the end would throw */
        t->nodes[node].numEpisodes += 1;
    }
    t->iters += 1;
    return 0;
}

int hk_oracle_tree_search(hk_oracle_tree* t, int iterations, int mode, uint64_t key, uint64_t* rng_state)
{
    for (int i = 0; i < iterations; ++i) {
        int rc = iterate(t, mode, key, rng_state);
        if (rc) return rc;
    }
    return 0;
}

/* getBestStatesSequence :108-122; returns the number of states written (at most max_out) */
int hk_oracle_tree_best_states(hk_oracle_tree* t, int mode, uint64_t key, uint64_t* rng_state, hk_game_state* out, int max_out)
{
    int node = 0, nb = 0;
    while (t->nodes[node].n_children > 0) {
        int j = ucs(t, node, mode, key, rng_state);
        if (j < 0) break;                                             /* catch (DivideByZeroException) { } :120 */
        node = t->nodes[node].children[j];
        const hk_game_state* s = &t->nodes[node].state;
        int all = 1;
        for (int i = 0; i < s->n_karts; ++i) all &= s->karts[i].section == s->lastCompletedSection;
        if (all) { if (nb < max_out) out[nb] = *s; ++nb; }
    }
    return nb < max_out ? nb : max_out;
}

int hk_oracle_tree_size(const hk_oracle_tree* t) { return t->n_nodes; }
/* a tree keeps the key it was started with (HierarchicalKartAgent.currentRoot continues its own random streams) */
void hk_oracle_tree_set_key(hk_oracle_tree* t, uint64_t key) { t->key = key; }
uint64_t hk_oracle_tree_key(const hk_oracle_tree* t) { return t->key; }
long long hk_oracle_tree_children_as_root(const hk_oracle_tree* t) { return t->childrenAsRoot; }

/* nodes in creation order; any output may be NULL.  first_child / next_sibling give the insertion-ordered child lists. */
void hk_oracle_tree_dump(const hk_oracle_tree* t, int32_t* parent, int32_t* gen, float* totalValue, int32_t* numEpisodes,
                         int32_t* n_children, int32_t* first_child, int32_t* next_sibling, hk_game_state* states)
{
    for (int i = 0; i < t->n_nodes; ++i) {
        const onode* n = &t->nodes[i];
        if (parent) parent[i] = n->parent;
        if (gen) gen[i] = n->gen;
        if (totalValue) totalValue[i] = n->totalValue;
        if (numEpisodes) numEpisodes[i] = n->numEpisodes;
        if (n_children) n_children[i] = n->n_children;
        if (first_child) first_child[i] = n->n_children ? n->children[0] : -1;
        if (states) states[i] = n->state;
    }
    if (next_sibling) {
        for (int i = 0; i < t->n_nodes; ++i) next_sibling[i] = -1;
        for (int i = 0; i < t->n_nodes; ++i)
            for (int j = 0; j + 1 < t->nodes[i].n_children; ++j) next_sibling[t->nodes[i].children[j]] = t->nodes[i].children[j + 1];
    }
}

/* n independent searches (tree r: key = seed + r, xorshift state rng_states[r] in mode 1) + getBestStatesSequence, OpenMP over trees.
 * Outputs as hk_mcts_search_seq_batch (include/hk_abi.h): root children in insertion order, -1 / 0 past the end. */
int hk_oracle_tree_search_batch(const hk_oracle_game* g, const hk_game_state* roots, int n, int iterations, int mode, uint64_t seed,
                                const uint64_t* rng_states, hk_game_state* best, int32_t* n_best, int max_seq, int32_t* root_gen,
                                int32_t* root_episodes, float* root_values, int32_t* n_nodes, int threads)
{
    int err = 0;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads > 0 ? threads : 1)
    for (int r = 0; r < n; ++r) {
        hk_oracle_tree* t = hk_oracle_tree_create(g, &roots[r]);
        uint64_t rng = rng_states ? rng_states[r] : 88172645463325252ULL;
        if (!rng) rng = 1;
        int rc = hk_oracle_tree_search(t, iterations, mode, seed + (uint64_t)r, &rng);
        if (rc) {
#pragma omp atomic write
            err = rc;
        }
        n_best[r] = hk_oracle_tree_best_states(t, mode, seed + (uint64_t)r, &rng, best + (size_t)r * max_seq, max_seq);
        if (n_nodes) n_nodes[r] = t->n_nodes;
        for (int j = 0; j < HK_MAX_ACTIONS; ++j) {
            const int in = j < t->nodes[0].n_children;
            const onode* c = in ? &t->nodes[t->nodes[0].children[j]] : 0;
            if (root_gen) root_gen[(size_t)r * HK_MAX_ACTIONS + j] = in ? c->gen : -1;
            if (root_episodes) root_episodes[(size_t)r * HK_MAX_ACTIONS + j] = in ? c->numEpisodes : 0;
            if (root_values) root_values[(size_t)r * HK_MAX_ACTIONS + j] = in ? c->totalValue : 0.0f;
        }
        hk_oracle_tree_destroy(t);
    }
    return err;
}
