/*
 * hk_abi.h — C-ABI of libhk_b200.so, the B200 (sm_100a) planning kernels behind HierarchicalKarting's
 * C# planners.  This is what the reference-side P/Invoke stubs bind (INTEGRATION.md shows them).
 *
 * The reference has NO native boundary today; the drop-in boundary is the public static C# API
 *   KartGame.AI.LQR.KartLQR.solveFeedbackLQR            Assets/Karting/Scripts/AI/LQR/KartLQR.cs:17
 *   KartGame.AI.MCTS.KartMCTS.constructSearchTree (x2)   Assets/Karting/Scripts/AI/MCTS/KartMCTS.cs:50,80
 *   KartGame.AI.MCTS.DiscreteGameState.{upNext,isOver,nextMoves,makeMove}
 *                                                        Assets/Karting/Scripts/AI/MCTS/KartDiscreteGame.cs:188,251,322,420
 * and every entry point below names the reference code whose body it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every buffer; nothing is retained after a call returns
 *     except the immutable hk_game handle;
 *   - every call returns HK_OK (0) or a negative hk_status; hk_last_error() gives the per-thread message;
 *     no exceptions, no callbacks;
 *   - re-entrant: the reference calls the LQNG solver from the Unity main thread and the MCTS from one
 *     background thread per agent (HierarchicalKartAgent.cs:246-283), so each host thread gets its own
 *     CUDA stream and scratch buffers lazily;
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns HK_ERR_NO_DEVICE.
 *
 * Dimensions (KartLQRDynamics.cs:27-28, MPC/KartMPC.cs:13-20): per kart state (x,z,v,h) = 4, control (a,w) = 2;
 * N players -> n = 4N joint states, m = 2N joint controls; T = horizon+1 backward steps (KartLQR.cs:64).
 */
#ifndef HK_ABI_H
#define HK_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HK_ABI_VERSION 1
#define HK_XDIM 4            /* LinearizedBicycle.xDim  KartLQRDynamics.cs:27 */
#define HK_UDIM 2            /* LinearizedBicycle.uDim  KartLQRDynamics.cs:28 */
#define HK_MAX_PLAYERS 4     /* HierarchicalKartAgent.cs:709-725 keeps N in 1..4 */
#define HK_MAX_HORIZON 31    /* reference uses 3 (HierarchicalKartAgent.cs:1201) */
#define HK_MAX_KARTS 4
#define HK_MAX_ACTIONS 36    /* velocity buckets 6..14 step bucket(>=1) x 4 lanes  KartDiscreteGame.cs:329-340 */
#define HK_MAX_PLIES 64      /* karts x treeSearchDepth upper bound for the trace buffers */

typedef enum hk_status {
    HK_OK = 0,
    HK_ERR_INVALID_ARGUMENT = -1,   /* MathNet would throw ArgumentException on a dimension mismatch */
    HK_ERR_NO_DEVICE = -2,          /* no CUDA device / wrong architecture: there is no CPU path */
    HK_ERR_CUDA = -3,
    HK_ERR_OUT_OF_MEMORY = -4,
    HK_ERR_NO_UPNEXT = -5           /* upNext() == -1: C# would throw ArgumentOutOfRangeException at KartDiscreteGame.cs:326 */
} hk_status;

/* ---- library ------------------------------------------------------------------------------------------------ */
int         hk_abi_version(void);
int         hk_init(int device);                 /* binds the calling process to one CUDA device (one process per GPU) */
void        hk_shutdown(void);
const char* hk_last_error(void);                 /* thread-local */
int         hk_device_count(void);
long long   hk_kernel_launch_count(void);        /* kernels launched by this library so far (all threads); for benchmarks */

/* Measurement aid (bench.py): the FP64 roofline denominator of THIS device, now.  Runs a DMMA m8n8k4 f64 stream (32 warps per SM) back to
 * back for at least `seconds` and reports its flop rate, the SM clock it actually ran at (clock64 cycles / event time of the last launch;
 * may be NULL) and the time it ran (may be NULL).  A short run gives the burst peak (boost clocks), a run of a second or more the
 * sustained one (clocks settled under FP64 power). */
int         hk_probe_fp64_peak(double seconds, double* tflops, double* sm_mhz_effective, double* seconds_run);

/* ---- LQNG: batched feedback linear-quadratic Nash game ------------------------------------------------------ */
/*
 * Replaces the body of KartLQR.solveFeedbackLQR (KartLQR.cs:17-128) for `batch` independent problems.
 * Host pointers, problem-major ("record per problem") layout; T = horizon+1; n = 4N; m = 2N.
 * If time_varying == 0 the [T] dimension of A,B,Q,q,R is absent (reference behaviour: time invariant).
 *   A     [batch][T?][N][4][4]   per-player A_i = KartLQRDynamics.getA()   (joint A = blockdiag, KartLQR.cs:33-37)
 *   B     [batch][T?][N][4][2]   per-player B_i = KartLQRDynamics.getB()   (joint B_i = zero-padded, KartLQR.cs:41-52)
 *   Q     [batch][T?][N][n][n]   KartLQRCosts.getQMatrix() of player i, row-major
 *   q     [batch][T?][N][n]      KartLQRCosts.getQVec()
 *   R     [batch][T?][N][2][2]   KartLQRCosts.getRMatrix()
 *   x0    [batch][n]             concatenated initial states (KartLQR.cs:54-60)
 * Outputs (any may be NULL except u0):
 *   u0    [batch][m]             -P x0 - alpha at t = 0 for every player; the reference returns u0[0..1] (KartLQR.cs:121-127)
 *   P     [batch][T][m][n]       feedback gains of every backward step t (index t = the loop variable of KartLQR.cs:64)
 *   alpha [batch][T][m]          feedback offsets
 *   traj  [batch][T+1][n]        closed-loop rollout x_{t+1} = A_t x_t + sum_k B_k,t u_k,t, u_t = -P_t x_t - alpha_t, x_0 = x0
 *   status[batch]                0 ok, 1 = a pivot of the coupled m x m system was exactly zero (MathNet yields inf/NaN silently)
 */
int hk_lqng_solve_batch(int batch, int n_players, int horizon, int time_varying,
                        const double* A, const double* B, const double* Q, const double* q, const double* R,
                        const double* x0,
                        double* u0, double* P, double* alpha, double* traj, int* status);

/* Same with DEVICE pointers in the same layout, asynchronous on `cuda_stream` (a cudaStream_t; NULL = the calling
 * thread's own stream).  Used by bench.py for the HBM-resident number and by pipelines that keep problems on the GPU. */
int hk_lqng_solve_batch_device(int batch, int n_players, int horizon, int time_varying,
                               const double* dA, const double* dB, const double* dQ, const double* dq, const double* dR,
                               const double* dx0,
                               double* du0, double* dP, double* dalpha, double* dtraj, int* dstatus,
                               void* cuda_stream);

/* What the C# shim of solveFeedbackLQR calls: batch of one, host pointers (KartLQR.cs:17). */
int hk_lqng_solve_one(int n_players, int horizon,
                      const double* A, const double* B, const double* Q, const double* q, const double* R,
                      const double* x0, double* u0 /* [m] */);

/*
 * Device-side problem assembly + solve: builds A,B from LinearizedBicycle
 * (KartLQRDynamics.cs:40-62) and Q,q,R from LQRCheckpointReachAvoidCost (KartLQRCosts.cs:57-140) on the GPU, so
 * only the compact per-player description crosses PCIe.  Per problem and per player i (private ordering
 * [self, its others...]; the cost blocks are used in that private order on the JOINT state, exactly as
 * KartLQRCosts.cs:62-94 + KartLQR.cs:62 do — quirk Q3 of SURVEY.md A.3):
 *   x0      [batch][N][4]      initial states (joint order)
 *   target  [batch][N][4]      targetState of player i
 *   tw      [batch][N][4]      targetWeights[x,z,v,h]
 *   cw      [batch][N]         controlWeight
 *   aw      [batch][N][N-1][2] avoidWeights[x],[z] of the k-th entry of player i's avoidDynamics list
 *   otgt    [batch][N][N-1][4] opponentTargetStates
 *   otw     [batch][N][N-1][3] opponentTargetWeights[x,z,v]
 *   dt                          (double)Time.fixedDeltaTime  HierarchicalKartAgent.cs:707
 */
int hk_lqng_assemble_solve_batch(int batch, int n_players, int horizon, double dt,
                                 const double* x0, const double* target, const double* tw, const double* cw,
                                 const double* aw, const double* otgt, const double* otw,
                                 double* u0, int* status);

/* The same call with the seven arrays interleaved per problem — one record of 13 N + 9 N (N-1) doubles (352 bytes for N = 2):
 *   records [batch][ x0 N x 4 | target N x 4 | tw N x 4 | cw N | aw N x (N-1) x 2 | otgt N x (N-1) x 4 | otw N x (N-1) x 3 ]
 * (what a C# caller fills as an array of blittable structs).  One host-to-device copy per chunk of the batch instead of seven, and one TMA
 * bulk copy per problem in the 2-kart kernel: the end-to-end form of the call when the host is the bottleneck. */
int hk_lqng_assemble_solve_packed(int batch, int n_players, int horizon, double dt, const double* records, double* u0, int* status);

/* ---- discrete race game + leaf-parallel rollouts ------------------------------------------------------------ */
typedef struct hk_section {      /* DiscretePositionTracker.cs:35-40 */
    float   insideR;             /* trackInsideRadius (0 => straight, :197) */
    float   length;              /* trackLength */
    float   width;               /* trackWidth */
    float   turnDeg;             /* turnDegrees */
    int32_t leftTurn;            /* bool */
    int32_t optimalLane;         /* 1..4 */
} hk_section;

typedef struct hk_kart {         /* ArcadeKart.Stats fields read by the game (ArcadeKart.cs:210,517-547; KartDiscreteGame.cs:70-116,168) */
    float accel, braking, topSpeed, reverseSpeed, maxGs, minGs, tireWearFactor;
} hk_kart;

typedef struct hk_game_params {  /* DiscreteGameParams HierarchicalKartAgent.cs:35-49 + RacingEnvController fields the game reads */
    int32_t velocityBucketSize, timePrecision, sectionWindow, treeSearchDepth;
    int32_t maxLaneChanges;              /* RacingEnvController.MaxLaneChanges :112 */
    float   collisionWindow;             /* unused by the reference game (collision filter is `false &&`, KartDiscreteGame.cs:398) */
    float   teamScoreRewardMultiplier;   /* RacingEnvController.cs:89 */
    int32_t maxEpisodeSteps;             /* RacingEnvController.cs:126 */
} hk_game_params;

typedef struct hk_kart_state {   /* DiscreteKartState KartDiscreteGame.cs:21-33; `name` is the kart's index in the list */
    int32_t player;              /* index into envController.Agents used by applyAction (:129); the reference always passes 0 */
    int32_t team, section, timeAtSection, min_velocity, max_velocity, lane, tireAge, laneChanges, infeasible;
} hk_kart_state;

typedef struct hk_action { int32_t min_velocity, max_velocity, lane; } hk_action;   /* DiscreteKartAction :14-19 */

typedef struct hk_game_state {   /* DiscreteGameState :174-182 (value part) */
    int32_t       n_karts;
    int32_t       initialSection, lastCompletedSection, finalSection;
    hk_kart_state karts[HK_MAX_KARTS];
} hk_game_state;

typedef struct hk_game hk_game;  /* immutable after creation => shareable between caller threads */

/* karts[n_karts]: constants of DiscreteGameState.kartAgents (used by nextMoves, :326-357);
 * env_karts[n_env_karts]: constants of envController.Agents (used by applyAction through hk_kart_state.player, :129);
 * env_karts == NULL => same table as karts. */
int  hk_game_create(const hk_section* sections, int n_sections, const hk_kart* karts, int n_karts,
                    const hk_kart* env_karts, int n_env_karts, const hk_game_params* params, hk_game** out);
void hk_game_destroy(hk_game* g);

/*
 * Exact-parity entry: for each of `batch` root states, play a fixed action sequence on the GPU with
 * DiscreteGameState.makeMove (:420-446) and report after every move the full state, upNext() (:188-243),
 * isOver() (:251-317; scores[] holds the first 2*HK_MAX_KARTS entries of the reference's score list, n_scores its
 * length) and the legal move set of nextMoves() (:322-415) in the rollout policy's sort order (KartMCTS.cs:256).
 *   actions     [batch][len]
 *   states_out  [batch][len+1]            state before move k (k = 0 is the root) ... after the last move
 *   upnext_out  [batch][len+1]
 *   over_out    [batch][len+1], n_scores_out [batch][len+1], scores_out [batch][len+1][2*HK_MAX_KARTS]
 *   n_moves_out [batch][len+1], moves_out [batch][len+1][HK_MAX_ACTIONS]  (sorted; generation-order index = vi*4+lane-1
 *                                         is stored in moves_index_out [batch][len+1][HK_MAX_ACTIONS])
 * Any output pointer may be NULL.  Moves after a terminal state are still applied (makeMove does not check).
 */
int hk_game_replay_batch(const hk_game* g, int batch, int len, const hk_game_state* roots, const hk_action* actions,
                         hk_game_state* states_out, int32_t* upnext_out, int32_t* over_out, int32_t* n_scores_out,
                         float* scores_out, int32_t* n_moves_out, hk_action* moves_out, int32_t* moves_index_out);

/*
 * Leaf-parallel rollouts (the GPU form of KartMCTS.processLeaf :124-159 + simulate :238-278): n_rollouts independent
 * playouts of the biased random policy from `leaf`, reduced on the GPU per FIRST action (generation-order index
 * a = vi*4 + lane-1, vi = (min_velocity-6)/bucket).  Rollout r uses Philox4x32-10 with key = seed and counter =
 * (rollout_offset + r, ply) — disjoint counter ranges make multi-GPU shards independent and reproducible.
 *   visit      [HK_MAX_ACTIONS]               rollouts whose first action was a
 *   reward_sum [HK_MAX_ACTIONS][HK_MAX_KARTS] sum over those rollouts of the terminal score list entry k (what
 *                                             backpropagate indexes with upNext(), KartMCTS.cs:284); NaN scores
 *                                             (max==min, KartDiscreteGame.cs:307) are counted in nan_count instead
 *   nan_count  [HK_MAX_ACTIONS]
 *   plies_sum  total number of plies played (for throughput accounting), may be NULL
 * If the leaf is already terminal every output is zero and the call returns HK_OK.
 */
int hk_mcts_rollouts(const hk_game* g, const hk_game_state* leaf, int64_t n_rollouts, uint64_t seed,
                     uint64_t rollout_offset, int64_t* visit, double* reward_sum, int64_t* nan_count,
                     int64_t* plies_sum);

/* Same rollouts, but every rollout also reports what it did, so that an oracle can replay it move by move:
 *   n_plies_out [n_rollouts]; actions_out [n_rollouts][HK_MAX_PLIES]; choice_out [n_rollouts][HK_MAX_PLIES] (index into
 *   the sorted legal list); n_scores_out [n_rollouts]; scores_out [n_rollouts][2*HK_MAX_KARTS]. */
int hk_mcts_rollouts_trace(const hk_game* g, const hk_game_state* leaf, int64_t n_rollouts, uint64_t seed,
                           uint64_t rollout_offset, int32_t* n_plies_out, hk_action* actions_out, int32_t* choice_out,
                           int32_t* n_scores_out, float* scores_out);

/* Many leaves in one launch (tree-parallel batches / many races): leaves[n_leaves], rollouts_per_leaf each;
 * outputs carry a leading [n_leaves] dimension.  Rollout ids are leaf_index * rollouts_per_leaf + r + rollout_offset. */
int hk_mcts_rollouts_multi(const hk_game* g, const hk_game_state* leaves, int n_leaves, int64_t rollouts_per_leaf,
                           uint64_t seed, uint64_t rollout_offset, int64_t* visit, double* reward_sum,
                           int64_t* nan_count, int64_t* plies_sum);

#define HK_MCTS_MAX_SEQ 16
/*
 * Batched tree search, LEAF-PARALLEL form: KartMCTS.constructSearchTree (KartMCTS.cs:50-106) followed by getBestStatesSequence (:108-122) for
 * n_roots independent root states in one launch, one thread block per tree — what planWithMCTS (HierarchicalKartAgent.cs:
 * 194-283) does per agent on a background thread, for every agent of many races at once.  Each of `iterations` iterations
 * walks the tree with upperConfidenceStrategy (:167-192) to a node without children (findLeaf :194-201), creates all its
 * children and plays rollouts_per_leaf rollouts through each (the reference's own leaf-parallel variant, processLeaf
 * :124-159), then backpropagates (:280-289).  Root r uses Philox key seed + r for its rollouts (ids it * R * HK_MAX_ACTIONS +
 * child * R + k, as hk_mcts_rollouts_multi numbers them) and counter stream (seed + r) ^ 0x9E3779B97F4A7C15 for the random
 * initial pick of upperConfidenceStrategy, so a host mirror driven by the same streams builds the same tree.
 * Terminal nodes: this mode is a HYBRID of the reference's two — the expansion is processLeaf's, but a terminal leaf (and a terminal
 * child, R times) backpropagates its terminal scores as the sequential simulate() path does (:246-249, :280-289), whereas the reference's
 * processLeaf returns without backpropagating (:126-129).  For what the reference's callers execute, use hk_mcts_forest_search.
 *   best_states   [n_roots][HK_MCTS_MAX_SEQ]  the states of getBestStatesSequence, n_best [n_roots] of them
 *   root_episodes [n_roots][HK_MAX_ACTIONS]   numEpisodes of the root's children in nextMoves() order (may be NULL)
 *   root_values   [n_roots][HK_MAX_ACTIONS]   their totalValue (may be NULL);  n_nodes [n_roots] tree sizes (may be NULL)
 */
int hk_mcts_search_batch(const hk_game* g, const hk_game_state* roots, int n_roots, int iterations, int rollouts_per_leaf,
                         uint64_t seed, hk_game_state* best_states, int32_t* n_best, int32_t* root_episodes,
                         double* root_values, int32_t* n_nodes);

/*
 * The tree search AS THE REFERENCE'S CALLERS RUN IT: KartMCTS.constructSearchTree with parallel == false (KartMCTS.cs:50-78 / :80-106 —
 * HierarchicalKartAgent.cs:250,271 never pass `parallel`): per iteration findLeaf (:194-201), ONE playout of simulate (:238-278) in which
 * every state becomes a tree node (:271-276), backpropagate from the playout's terminal node (:280-289); then getBestStatesSequence
 * (:108-122).  The reference's wall-clock budget T (:55) becomes an iteration count.  One GPU thread owns one tree, so the call pays
 * off for many trees at once (every agent of many races); a single tree runs at the latency of one thread.
 * Trees live in a device-resident forest and survive between calls, as HierarchicalKartAgent.currentRoot does (:265-283): tree r is
 * (re)started from roots[r] where fresh[r] > 0 (constructSearchTree(state), fresh == NULL: all), continued where fresh[r] == 0
 * (constructSearchTree(root); roots[r] is ignored) and left alone where fresh[r] < 0 (its outputs are not written).  Random sources: tree r started by a call with `seed` has key = seed + r for life;
 * policy index of iteration `it` (counted over the life of the tree), ply p of its playout = word 0 of Philox4x32-10(key, counter
 * (it, 0, p, 0)) through hk_policy_cdf; the random initial pick of upperConfidenceStrategy (:169) = word 0 of
 * Philox4x32-10(key ^ 0x9E3779B97F4A7C15, counter (picks so far, 0, 0, 0)) modulo the child count.  totalValue is float32, updated in
 * the reference's order.
 *   best_states [n_trees][HK_MCTS_MAX_SEQ], n_best [n_trees]   getBestStatesSequence of every tree after this call
 *   n_nodes     [n_trees]  tree sizes (may be NULL)
 *   status      [n_trees]  (may be NULL) 0 ok; 1 upNext() == -1 was reached (ArgumentOutOfRangeException, KartDiscreteGame.cs:326);
 *                          2 UCTWeight divided by zero inside findLeaf (DivideByZeroException, KartMCTS.cs:164); 3 max_nodes_per_tree
 *                          reached (the search of that tree stopped early).  Status 1 makes the call return HK_ERR_NO_UPNEXT.
 * max_nodes_per_tree must cover 1 + (iterations over the life of a tree) x (plies of a playout <= sum over karts of finalSection - section).
 */
typedef struct hk_mcts_node {    /* one KartMCTSNode (KartMCTS.cs:18-38) of a device-resident tree; the state is not stored (replay the path) */
    uint64_t child_mask;         /* generation indices (vi*4 + lane-1) of the actions in `children` */
    float    totalValue;         /* :23 */
    int32_t  numEpisodes;        /* :24 */
    int32_t  first_child, last_child, next_sibling;   /* children in insertion order = the Dictionary's enumeration order; -1 = none */
    uint8_t  gen;                /* generation index of the action that created this node (255: the root) */
    uint8_t  n_legal;            /* state.nextMoves().Count (255: never evaluated, i.e. a terminal node) */
    int8_t   upnext;             /* state.upNext() */
    uint8_t  pad_;
} hk_mcts_node;

typedef struct hk_mcts_forest hk_mcts_forest;
int  hk_mcts_forest_create(const hk_game* g, int n_trees, int max_nodes_per_tree, hk_mcts_forest** out);
void hk_mcts_forest_destroy(hk_mcts_forest* f);
int  hk_mcts_forest_search(hk_mcts_forest* f, const hk_game_state* roots, const int32_t* fresh, int iterations, uint64_t seed,
                           hk_game_state* best_states, int32_t* n_best, int32_t* n_nodes, int32_t* status);
/* Nodes of one tree in creation order (node 0 = root; a node's parent is the node whose child list holds it): at most max_nodes
 * records are written, *n_nodes_out receives the tree size. */
int  hk_mcts_forest_nodes(const hk_mcts_forest* f, int tree, hk_mcts_node* nodes_out, int max_nodes, int32_t* n_nodes_out);
/* One-shot form: forest of n_roots fresh trees sized for `iterations`, searched, destroyed.  root_gen / root_episodes / root_values
 * [n_roots][HK_MAX_ACTIONS] (each may be NULL): generation index (-1 past the end), numEpisodes and totalValue of the root's children
 * in insertion order. */
int  hk_mcts_search_seq_batch(const hk_game* g, const hk_game_state* roots, int n_roots, int iterations, uint64_t seed,
                              hk_game_state* best_states, int32_t* n_best, int32_t* root_gen, int32_t* root_episodes,
                              float* root_values, int32_t* n_nodes);

/* The rollout policy's index distribution for `cnt` legal moves (KartMCTS.cs:266-269 via NextGaussian :218-236):
 * cdf_out[k] = P(index <= k) as a 32-bit threshold, exactly what the kernels sample from. cnt in 1..HK_MAX_ACTIONS. */
int hk_policy_cdf(int cnt, uint32_t* cdf_out /* [cnt] */);

/* ---- closed loop without PhysX: problem recipe + kinematic plant + checkpoint bookkeeping ---- */
/*
 * The reference closes its loop through Unity's PhysX ArcadeKart, which is out of scope; the planners themselves assume
 * the kinematic model x += dt v cos h, z += dt v sin h, h += dt w, v += dt a (MPC/KartMPCDynamics.cs:55-70, the model
 * LinearizedBicycle linearises, LQR/KartLQRDynamics.cs:40-62).  These entry points run that model for many independent
 * 2-kart races entirely on the GPU: per step and per agent the LQNG problem recipe of HierarchicalKartAgent.SolveLQR
 * (HierarchicalKartAgent.cs:699-1197, raycast-free branches), the solve, the actuator map (:1206-1224), the plant and
 * the checkpoint bookkeeping of OnTriggerEnter (:611-662, lane by DiscretePositionTracker.CalculateLane :116-148).
 */
#define HK_MAX_SECTIONS 64       /* Oval has 24, Complex 41 */

typedef struct hk_race_kart {    /* one kart of one race; 64 bytes */
    double  x, z, v, h;          /* LQR state order (MPC/KartMPC.cs:13-20); h in [0, 2 pi) (HierarchicalKartAgent.cs:734-736) */
    float   steer;               /* m_FinalStats.Steer (ArcadeKart.cs:300); max yaw rate = 0.4 steer (:505-510) */
    int32_t section;             /* m_SectionIndex: grows without bound, used % n_sections (RacingEnvController.cs:758-784) */
    int32_t lane;                /* m_Lane, 1..4 */
    int32_t laneChanges;         /* m_LaneChanges */
    int32_t illegalLaneChanges;  /* m_IllegalLaneChanges */
    int32_t sectionStep;         /* sectionTimes[m_SectionIndex]: episodeSteps at the last crossing (:650) */
    int32_t active;              /* 0 once the kart has reached goalSection (:651-654) */
    int32_t team;                /* RacingEnvController.getTeamID (:786-796); ignored by the 2-kart entry points (every kart its own team) */
} hk_race_kart;

#define HK_MAX_LAPS 8
typedef struct hk_race_plan {    /* per-agent tables: m_UpcomingLanes / m_UpcomingVelocities keyed by section % n_sections, + metrics */
    int8_t  lane[HK_MAX_SECTIONS];      /* 0 = key absent */
    float   vel[HK_MAX_SECTIONS];
    int8_t  oppLane[HK_MAX_SECTIONS];   /* opponentUpcomingLanes[other] (filled from MCTS bestStates, HierarchicalKartAgent.cs:396-400) */
    float   oppVel[HK_MAX_SECTIONS];
    int32_t sectionTimes[HK_MAX_SECTIONS];   /* sectionTimes[index] = episodeSteps (:650), keyed by index % n_sections; read by planWithMCTS :213 */
    int32_t lapStep[HK_MAX_LAPS];       /* episodeSteps at which lap k+1 was completed (what TelemetryViewer.cs:60-72 derives) */
    float   avgLaneDiff, avgVelDiff;    /* AverageLaneDifference / AverageVelDifference (KartAgent.cs:226-239) */
} hk_race_plan;

typedef struct hk_race_params {
    double  dt;                  /* (double)Time.fixedDeltaTime = (double)0.02f */
    float   accel, braking, coastingDrag, topSpeed;   /* ArcadeKart.Stats (7, 16, 5, 15) */
    float   gateHalfWidth;       /* lateral half extent of a checkpoint trigger */
    int32_t maxLaneChanges;      /* RacingEnvController.MaxLaneChanges */
    int32_t goalSection;         /* RacingEnvController.goalSection = laps * n_sections */
    int32_t highModeMcts;        /* HighLevelMode.MCTS (1) or Fixed (0): selects weights and the +2*bucket velocity margin */
    int32_t velocityBucketSize, treeSearchDepth;      /* DiscreteGameParams */
    int32_t planEvery;           /* the high level replans when episodeSteps % planEvery == 0 (100, :334) */
    int32_t horizon;             /* LQNG horizon (3, :1201) */
} hk_race_params;

typedef struct hk_track hk_track;   /* immutable device-resident geometry */
/* sections[n] as for hk_game_create; trigger_xz [n][2], forward_xz [n][2] (unit vector of the checkpoint's transform.forward
 * in (x, z)), lane_xz [n][4][2] lane collider centres. */
int  hk_track_create(const hk_section* sections, const double* trigger_xz, const double* forward_xz, const double* lane_xz,
                     int n_sections, hk_track** out);
void hk_track_destroy(hk_track* t);

/*
 * Problem recipe only (device-side K5 of SURVEY.md §7): for each of n_races races and each of its 2 agents builds the compact
 * description hk_lqng_assemble_solve_batch takes.  Problem index = 2 race + ego; player 0 = ego, player 1 = the other kart.
 * Host pointers.  Outputs [2 n_races][...] as documented at hk_lqng_assemble_solve_batch with N = 2.
 */
int hk_race_recipe(const hk_track* t, const hk_race_params* p, int n_races, const hk_race_kart* karts /* [n_races][2] */,
                   const hk_race_plan* plans /* [n_races][2] */,
                   double* x0, double* target, double* tw, double* cw, double* aw, double* otgt, double* otw);

/* One plant + bookkeeping step for n_karts karts with the given controls u [n_karts][2] = (KartLQR output: acceleration
 * command, angular velocity); karts and plans are updated in place (host pointers). */
int hk_race_step(const hk_track* t, const hk_race_params* p, int n_karts, int episode_step, const double* u,
                 hk_race_kart* karts, hk_race_plan* plans);

/* planFixed (HierarchicalKartAgent.cs:145-166) for n_karts agents, in place. */
int hk_race_plan_fixed(const hk_track* t, const hk_race_params* p, int n_karts, const hk_race_kart* karts, hk_race_plan* plans);

/*
 * The loop: n_steps steps (episodeSteps = first_step .. first_step + n_steps - 1) of n_races independent 2-kart races,
 * device-resident between the initial upload and the final download of karts/plans: replan (Fixed high level) when
 * episodeSteps % planEvery == 0 and > 0, recipe, assemble, LQNG solve, plant + bookkeeping.
 * u_last [n_races][2][2] (may be NULL) receives the controls of the last step; lqng_status_nonzero (may be NULL) the number
 * of solves that reported a zero pivot.
 */
int hk_race_run(const hk_track* t, const hk_race_params* p, int n_races, int first_step, int n_steps,
                hk_race_kart* karts, hk_race_plan* plans, double* u_last, int64_t* lqng_status_nonzero);

/*
 * The same loop with the MCTS high level (HighLevelMode.MCTS, HierarchicalKartAgent.cs:331-353; p->highModeMcts must be 1), entirely ON
 * THE DEVICE: root states (planWithMCTS :180-245), tree search and the waypoint hand-off (:366-402) are kernels between two steps;
 * karts / plans cross PCIe only at the start and the end of a call.  A planner holds what survives between planning events and between
 * calls — per agent the tree (currentRoot), CyclesRootProcessed, and the result of a search whose background thread has not finished:
 *   - a planning event happens when episodeSteps % planEvery == 0, 0 < episodeSteps < maxEpisodeSteps (:331), for active agents, and
 *     (first_iterations > 0) when the episode begins (step 0: planWithMCTS(T: 1.5), :85-93);
 *   - an agent without a tree starts a new one from its current root state (:175-262); an agent whose tree is still valid continues it
 *     while CyclesRootProcessed < reuse_cycles (:265-283) and otherwise does not plan at all; crossing a checkpoint drops the tree
 *     (currentRoot = null; CyclesRootProcessed = 0, :660-661);
 *   - the result (bestStates, CyclesRootProcessed = 1 / += 1) lands apply_delay steps after the search started — the reference searches
 *     on a background thread for T seconds of wall clock (:246-283), i.e. T / fixedDeltaTime physics steps; 0 = at once — and is handed
 *     to the plan tables at that step.
 * Tree of agent a (= 2 race + ego) searched by the event of step s: key = seed + (s / planEvery) * n_agents + a (hk_mcts_forest_search /
 * hk_mcts_search_batch streams), a function of the absolute step, so a loop advanced in blocks equals one long call.
 */
typedef struct hk_race_mcts_params {
    int32_t  mode;               /* 0: the sequential search the reference runs (hk_mcts_forest_search); 1: leaf-parallel (hk_mcts_search_batch) */
    int32_t  iterations;         /* budget of a replan (the reference: T = 0.9 s of wall clock, HierarchicalKartAgent.cs:172,271) */
    int32_t  first_iterations;   /* budget of the plan at episode start (T = 1.5 s, :93); 0 = no plan at step 0 */
    int32_t  rollouts_per_leaf;  /* mode 1 only */
    int32_t  reuse_cycles;       /* mode 0: bound on CyclesRootProcessed (3, :265); 0 = every plan starts a new tree */
    int32_t  apply_delay;        /* steps between the start of a search and the landing of its result; must be < planEvery */
    uint64_t seed;
    int32_t  max_tree_nodes;     /* mode 0: nodes reserved per agent's tree; 0 = 1 + (max(first_iterations, iterations) + reuse_cycles *
                                    iterations) * 2 treeSearchDepth.  A checkpoint crossing while a continued search is in flight resets
                                    CyclesRootProcessed, so a tree CAN be continued more often than reuse_cycles; a tree that runs out of
                                    nodes stops growing (tree_status 3 in hk_race_planner_state) */
    int32_t  pad_;
} hk_race_mcts_params;

typedef struct hk_race_planner hk_race_planner;
int  hk_race_planner_create(const hk_game* game, const hk_race_mcts_params* mp, int n_races, hk_race_planner** out);
void hk_race_planner_destroy(hk_race_planner* planner);
/* per agent: currentRoot != null, CyclesRootProcessed, status of the tree's last search as in hk_mcts_forest_search (each may be NULL) */
int  hk_race_planner_state(const hk_race_planner* planner, int32_t* root_valid, int32_t* cycles, int32_t* tree_status);
int  hk_race_run_planned(const hk_track* t, const hk_race_params* p, hk_race_planner* planner, int n_races, int first_step, int n_steps,
                         hk_race_kart* karts, hk_race_plan* plans, double* u_last, int64_t* lqng_status_nonzero);

/* One-call form without state between calls: a temporary planner with reuse_cycles = 0, apply_delay = 0, first_iterations = 0;
 * rollouts_per_leaf == 0 selects mode 0 (the reference's sequential search), > 0 mode 1 (leaf-parallel). */
int hk_race_run_mcts(const hk_track* t, const hk_race_params* p, const hk_game* game, int iterations, int rollouts_per_leaf,
                     uint64_t seed, int n_races, int first_step, int n_steps, hk_race_kart* karts, hk_race_plan* plans,
                     double* u_last, int64_t* lqng_status_nonzero);

/* ---- races with up to 4 karts and teams (the Duos scenes: 4 agents, 2 teams) ----------------------------------------------------------
 * The same closed loop for n_races races of karts_per_race (K = 2..4) karts each, agent id = K race + e, with everything SolveLQR does when
 * the environment has more than two agents (HierarchicalKartAgent.cs:702-725, 930-1003, 1004-1190):
 *   - players of agent e's game: [e] + its teammates + its opponents (environment order), of which only those within 8 m of e take part
 *     (float32 magnitude, :709-725) => N in 1..4 per problem and step; nearbyAgents = max(N - 1, 1) scales the weights (:932-962,
 *     :985-987, :1083-1085); teammates use multiplier / 2 (:1113) and the teammate-target weights (:1168-1187);
 *   - every player's cost is built in ITS private order [its opponents..., its teammates...] but multiplies the joint state ordered for
 *     the ego (quirk Q3 of SURVEY.md A.3) — reproduced, not fixed;
 *   - the solve runs every lqr_every-th step (4 when the environment has more than 2 agents, :317) and the controls are held in between;
 *   - a problem with N < 4 players is solved in the 4-player frame with decoupled dummy players (A = I, B = 0, Q = 0, R = I): their
 *     gains stay exactly zero and the real players' results are those of the N-player game.
 * beliefs [n_races][K][K]: agent e's opponentUpcomingLanes / opponentUpcomingVelocities about kart o (entry [e][e] unused).
 */
typedef struct hk_race_belief { int8_t lane[HK_MAX_SECTIONS]; float vel[HK_MAX_SECTIONS]; } hk_race_belief;

/* Problem recipe only.  Outputs in the 4-player layout of hk_lqng_assemble_solve_batch (N = 4): x0 [n][4][4], target [n][4][4],
 * tw [n][4][4], cw [n][4], aw [n][4][3][2], otgt [n][4][3][4], otw [n][4][3][3] with n = K n_races, dummy players zero (cw 1);
 * n_players [n] and players [n][4] (race-local kart of joint slot i, -1 = dummy). */
int hk_raceN_recipe(const hk_track* t, const hk_race_params* p, int karts_per_race, int n_races, const hk_race_kart* karts,
                    const hk_race_plan* plans, const hk_race_belief* beliefs, int32_t* n_players, int32_t* players,
                    double* x0, double* target, double* tw, double* cw, double* aw, double* otgt, double* otw);

/* The loop.  u_hold [n_races][K][2] carries the controls held between two solves across calls (in/out; zeros at the start).  planner may
 * be NULL (Fixed high level / the caller plans); with a planner made by hk_raceN_planner_create the MCTS high level runs on the device as
 * in hk_race_run_planned, with the game's team scoring (KartDiscreteGame.cs:271-310) deciding: root states hold every agent within
 * sectionWindow sections (HierarchicalKartAgent.cs:182-233) with its team, the hand-off also fills the beliefs about the other karts. */
int hk_raceN_planner_create(const hk_game* game, const hk_race_mcts_params* mp, int karts_per_race, int n_races, hk_race_planner** out);
int hk_raceN_run(const hk_track* t, const hk_race_params* p, hk_race_planner* planner, int karts_per_race, int lqr_every, int n_races,
                 int first_step, int n_steps, hk_race_kart* karts, hk_race_plan* plans, hk_race_belief* beliefs, double* u_hold,
                 int64_t* lqng_status_nonzero);

/*
 * The loops on race states that STAY in device memory between calls (a long race advanced in blocks, or states produced by other
 * kernels): the same work as hk_race_run / hk_race_run_planned / hk_raceN_run with DEVICE pointers in the same layouts and no upload or
 * download — a call to the host-pointer entries moves ~1 KB (2 karts) / ~2.3 KB (Duos) per agent each way, which costs more than 200 steps
 * of the loop itself.  cuda_stream: a cudaStream_t, NULL = the calling thread's own stream.  The call returns after the work has finished
 * (it brings back lqng_status_nonzero and the trees' status).  planner may be NULL.
 *   hk_race_run_device:  d_u (may be NULL) [n_races * 2][4]: the LQNG u0 record of the last step of every agent, its own controls first
 *   hk_raceN_run_device: d_u_hold [n_races * K][8]: the held u0 record of every agent (4-player frame, its own controls first; in/out,
 *                        zeros at the start of a race)
 */
int hk_race_run_device(const hk_track* t, const hk_race_params* p, hk_race_planner* planner, int n_races, int first_step, int n_steps,
                       hk_race_kart* d_karts, hk_race_plan* d_plans, double* d_u, int64_t* lqng_status_nonzero, void* cuda_stream);
int hk_raceN_run_device(const hk_track* t, const hk_race_params* p, hk_race_planner* planner, int karts_per_race, int lqr_every, int n_races,
                        int first_step, int n_steps, hk_race_kart* d_karts, hk_race_plan* d_plans, hk_race_belief* d_beliefs, double* d_u_hold,
                        int64_t* lqng_status_nonzero, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* HK_ABI_H */
