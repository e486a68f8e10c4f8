/* hk_oracle.h — CPU ORACLE (test infrastructure, NOT the product). See hk_oracle_lqng.c / hk_oracle_game.c.
 * Shares only the plain-data struct definitions of include/hk_abi.h with the product. */
#ifndef HK_ORACLE_H
#define HK_ORACLE_H
#include <stdint.h>
#include "../include/hk_abi.h"
#ifdef __cplusplus
extern "C" {
#endif
/* LQNG (double) */
void hk_oracle_bicycle_A(double dt, const double* x0, double* A);
void hk_oracle_bicycle_B(double dt, double* B);
void hk_oracle_cost(int n_other, const double* target, const double* tw, double cw, const double* aw, const double* otgt,
                    const double* otw, double* Q, double* q, double* R);
int hk_oracle_lqng_solve(int N, int horizon, int time_varying, const double* A, const double* B, const double* Q,
                         const double* q, const double* R, const double* x0, double* u0, double* P, double* alpha, double* traj);
int hk_oracle_lqng_solve_batch(int batch, int N, int horizon, int time_varying, const double* A, const double* B,
                               const double* Q, const double* q, const double* R, const double* x0, double* u0, double* P,
                               double* alpha, double* traj, int* status, int threads);
/* discrete game (float32/int32) */
typedef struct hk_oracle_game hk_oracle_game;
int  hk_oracle_game_create(const hk_section* s, int n_sections, const hk_kart* karts, int n_karts, const hk_kart* env_karts,
                           int n_env_karts, const hk_game_params* p, hk_oracle_game** out);
void hk_oracle_game_destroy(hk_oracle_game* g);
float hk_oracle_max_speed_for_radius_and_wear(const hk_kart* k, float radius, float wear);
float hk_oracle_compute_toc(const hk_kart* k, float distance, float radius, float tireWear, float initV, float finalV);
hk_kart_state hk_oracle_apply_action(const hk_oracle_game* g, const hk_kart_state* s, hk_action a);
int  hk_oracle_up_next(const hk_oracle_game* g, const hk_game_state* st);
int  hk_oracle_next_moves(const hk_oracle_game* g, const hk_game_state* st, hk_action* out, int* gen_index);
hk_game_state hk_oracle_make_move(const hk_oracle_game* g, const hk_game_state* st, hk_action a, int* err);
int  hk_oracle_is_over(const hk_oracle_game* g, const hk_game_state* st, float* scores, int* n_scores);
int  hk_oracle_policy_moves(const hk_oracle_game* g, const hk_game_state* st, hk_action* out, int* gen_index);
void hk_oracle_philox4x32_10(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);
void hk_oracle_policy_cdf(int cnt, uint32_t* cdf);
int  hk_oracle_policy_index(int cnt, const uint32_t* cdf, uint32_t u);
int  hk_oracle_reference_policy_index(int cnt, uint64_t* rng_state);
int  hk_oracle_rollout(const hk_oracle_game* g, const hk_game_state* leaf, int mode, uint64_t seed, uint64_t rollout_id,
                       uint64_t* rng_state, hk_action* actions_out, int* choice_out, float* scores, int* n_scores,
                       hk_game_state* terminal);
int  hk_oracle_rollouts(const hk_oracle_game* g, const hk_game_state* leaf, int64_t n_rollouts, int mode, uint64_t seed,
                        uint64_t rollout_offset, int64_t* visit, double* reward_sum, int64_t* nan_count, int64_t* plies_sum);
/* sequential tree search exactly as the reference's callers run it (hk_oracle_mcts.c): constructSearchTree with parallel == false */
typedef struct hk_oracle_tree hk_oracle_tree;
hk_oracle_tree* hk_oracle_tree_create(const hk_oracle_game* g, const hk_game_state* root);
void hk_oracle_tree_destroy(hk_oracle_tree* t);
int  hk_oracle_tree_search(hk_oracle_tree* t, int iterations, int mode, uint64_t key, uint64_t* rng_state);
int  hk_oracle_tree_best_states(hk_oracle_tree* t, int mode, uint64_t key, uint64_t* rng_state, hk_game_state* out, int max_out);
int  hk_oracle_tree_size(const hk_oracle_tree* t);
long long hk_oracle_tree_children_as_root(const hk_oracle_tree* t);
void hk_oracle_tree_dump(const hk_oracle_tree* t, int32_t* parent, int32_t* gen, float* totalValue, int32_t* numEpisodes,
                         int32_t* n_children, int32_t* first_child, int32_t* next_sibling, hk_game_state* states);
void hk_oracle_tree_set_key(hk_oracle_tree* t, uint64_t key);
uint64_t hk_oracle_tree_key(const hk_oracle_tree* t);
int  hk_oracle_tree_search_batch(const hk_oracle_game* g, const hk_game_state* roots, int n, int iterations, int mode, uint64_t seed,
                                 const uint64_t* rng_states, hk_game_state* best, int32_t* n_best, int max_seq, int32_t* root_gen,
                                 int32_t* root_episodes, float* root_values, int32_t* n_nodes, int threads);
/* closed loop without PhysX (hk_oracle_race.c): recipe, planFixed, plant + bookkeeping, the loop */
void hk_oracle_race_recipe_one(const hk_section* sections, const double* trig, const double* fwd, const double* lane, int n_sections,
                               const hk_race_params* p, const hk_race_kart* karts, const hk_race_plan* plans, int e,
                               double* x0, double* target, double* tw, double* cw, double* aw, double* otgt, double* otw);
void hk_oracle_raceN_recipe_one(const hk_section* sections, const double* trig, const double* fwd, const double* lane, int n_sections,
                                const hk_race_params* p, int K, const hk_race_kart* karts, const hk_race_plan* plans,
                                const hk_race_belief* beliefs, int e, int* n_players, int* players, double* x0, double* target, double* tw,
                                double* cw, double* aw, double* otgt, double* otw);
void hk_oracle_race_recipe(const hk_section* sections, const double* trig, const double* fwd, const double* lane, int n_sections,
                           const hk_race_params* p, int n_races, const hk_race_kart* karts, const hk_race_plan* plans,
                           double* x0, double* target, double* tw, double* cw, double* aw, double* otgt, double* otw);
void hk_oracle_race_plan_fixed(const hk_section* sections, int n_sections, const hk_race_params* p, int n_karts,
                               const hk_race_kart* karts, hk_race_plan* plans);
void hk_oracle_race_step(const hk_section* sections, const double* trig, const double* fwd, const double* lane, int n_sections,
                         const hk_race_params* p, int n_karts, int episode_step, const double* u, hk_race_kart* karts,
                         hk_race_plan* plans);
long long hk_oracle_race_run(const hk_section* sections, const double* trig, const double* fwd, const double* lane, int n_sections,
                             const hk_race_params* p, int n_races, int first_step, int n_steps, hk_race_kart* karts,
                             hk_race_plan* plans, double* u_last);
/* MCTS high level of the loop (hk_oracle_race.c): planWithMCTS's root state, the waypoint hand-off, the planner's schedule */
int  hk_oracle_race_mcts_root(const hk_race_params* p, const hk_game_params* gp, int n_sections, const hk_race_kart* karts,
                              const hk_race_plan* plans, int n_agents_in_race, int ego, hk_game_state* st, int* nearby);
void hk_oracle_race_apply_best(int n_sections, const hk_race_kart* karts, hk_race_plan* plans, int ego, const int* nearby,
                               const hk_game_state* best, int n_best);
typedef struct hk_oracle_planner hk_oracle_planner;
hk_oracle_planner* hk_oracle_planner_create(const hk_oracle_game* g, const hk_game_params* gp, const hk_race_mcts_params* mp, int n_races);
void hk_oracle_planner_destroy(hk_oracle_planner* pl);
void hk_oracle_planner_state(const hk_oracle_planner* pl, int32_t* root_valid, int32_t* cycles);
long long hk_oracle_race_run_planned(const hk_section* sections, const double* trig, const double* fwd, const double* lane, int n_sections,
                                     const hk_race_params* p, hk_oracle_planner* pl, int n_races, int first_step, int n_steps,
                                     hk_race_kart* karts, hk_race_plan* plans, double* u_last);
/* helpers for the KAT tests (track formulas) */
float hk_oracle_distance_to_travel(const hk_section* s, int a, int b);
float hk_oracle_radius_of_lane(const hk_section* s, int a, int b);
float hk_oracle_tire_load(const hk_section* s, float velocity, int a, int b);
#ifdef __cplusplus
}
#endif
#endif
