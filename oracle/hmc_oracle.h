/*
 * hmc_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the reference's hot path
 * (margaritacolberg/hybridmc: src/hardspheres.cc, src/crankshaft.cc,
 * src/entropy.cc, src/main.cc:45-305) used as the parity oracle for the CUDA
 * engine.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product path never does.
 *
 * Parity status: PINNED — checked against (a) the golden vectors of the
 * reference's own unit tests (tests/golden/reference_kats.json, transcribed
 * from src/hybridmc_test.cc) and (b) bit-exact event traces / trajectories of
 * the unmodified reference sources compiled into oracle/_ref/libhmc_ref.so
 * (tests/test_oracle_vs_ref.py, fixtures under tests/golden/).
 */
#ifndef HMC_ORACLE_H
#define HMC_ORACLE_H

#include "../include/hybridmc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_sim orc_sim;

/* ---- one replica ---------------------------------------------------------- */
orc_sim *orc_create(const hmc_params *p);
void orc_destroy(orc_sim *s);
void orc_set_pos(orc_sim *s, const double *pos /*[nbeads][3]*/);
void orc_get_pos(const orc_sim *s, double *pos);
void orc_get_vel(const orc_sim *s, double *vel);
void orc_set_config(orc_sim *s, uint64_t config);
uint64_t orc_get_config(const orc_sim *s);
void orc_init_config_from_pos(orc_sim *s);
void orc_set_s_bias(orc_sim *s, const double *s_bias, int nstates);
void orc_get_s_bias(const orc_sim *s, double *s_bias, int nstates);
void orc_init_s(orc_sim *s); /* init_s, entropy.cc:14-20 */
void orc_reset_clocks(orc_sim *s);
void orc_reset_counts(orc_sim *s);

/* initialize_system + run_step (+ sampling and energy check for MAIN/WL) for
 * one step index; normals[nbeads][3] are init_vel's draws before the COM
 * shift.  Returns hmc_error_flag bits (0 = ok).  For HMC_PHASE_WL `step` is
 * the trajectory index (outer iteration = step / nstates) and the ±gamma
 * update of wang_landau is applied. */
int orc_md_step(orc_sim *s, int phase, unsigned step, const double *normals);

/* n_moves crankshaft moves from a raw-word tape (see hybridmc_b200.h,
 * hmc_crankshaft_injected).  *cursor is advanced.  Returns error flags. */
int orc_crankshaft(orc_sim *s, unsigned move_offset, unsigned n_moves,
                   const uint32_t *words, const double *sin_tab,
                   const double *cos_tab, unsigned n_words, unsigned *cursor);

/* moves completed by the last orc_crankshaft call (< n_moves on tape underrun) */
unsigned orc_last_moves_done(const orc_sim *s);

void orc_enable_trace(orc_sim *s, unsigned capacity);
unsigned orc_get_trace(const orc_sim *s, hmc_event *out, unsigned max_events);
void orc_enable_samples(orc_sim *s, unsigned dist_rows, unsigned config_int_cap);
unsigned orc_get_dist(const orc_sim *s, double *dist);
unsigned orc_get_config_int(const orc_sim *s, uint64_t *config_int);
void orc_get_counts(const orc_sim *s, uint64_t *config_count, int nstates,
                    uint64_t *formed, uint64_t *broken);
/* [0] collisions, [1] cell crossings, [2] crankshaft attempts, [3] accepted,
 * [4] heap pops (valid + stale), [5] pair predictions made */
void orc_get_event_counts(const orc_sim *s, uint64_t ev[6]);
int orc_get_unique_config(const orc_sim *s, double *pos, double *vel);

/* ---- primitives (known-answer tests) -------------------------------------- */
int orc_t_until_inner_coll(double x, double y, double z, double vx, double vy,
                           double vz, double sigma2, double *t);
void orc_t_until_outer_coll(double x, double y, double z, double vx, double vy,
                            double vz, double sigma2, double *t);
int orc_v_after_coll(double x, double y, double z, double sigma2, double *vi,
                     double *vj, int has_dS, double dS, double m);
void orc_mindist(double l, double *d /*[3]*/);
void orc_minpos(double l, double *x /*[3]*/);
int orc_t_until_cell(double x, double y, double z, double l, unsigned ncell,
                     double lcell, double vx, double vy, double vz, unsigned ixo,
                     unsigned iyo, unsigned izo, double *dt, unsigned *wall,
                     unsigned *cell_new /*[3]*/);
void orc_shift_vel_cm(double *vel, unsigned nbeads);
int orc_check_local_dist(const orc_sim *s);
int orc_check_nonlocal_dist(const orc_sim *s);
void orc_rodrigues(const orc_sim *s, double sin_t, double cos_t, unsigned ind,
                   double *pos_trial);
uint64_t orc_config_int(const orc_sim *s, const double *pos_trial);
double orc_rc2_inner(const hmc_params *p, int p_bond);
double orc_rc2_outer(const hmc_params *p, int t_bond_nonbonded_under_stair,
                     int p_bond);

/* ---- entropy.cc ------------------------------------------------------------ */
void orc_compute_entropy(const uint64_t *config_count, double *s_bias,
                         int nstates, double pos_scale, double neg_scale);
double orc_compute_g_val(const uint64_t *config_count, int nstates);
int orc_g_test(const uint64_t *config_count, int nstates, double sig_level,
               double *g_val, double *g_crit, double *p_val);

/* ---- libstdc++ RNG restatement (std::mt19937 + distributions) --------------- */
typedef struct orc_mt19937 {
  uint32_t mt[624];
  int idx;
} orc_mt19937;
void orc_mt_seed_seq(orc_mt19937 *g, const uint32_t *seeds, unsigned n);
void orc_mt_seed(orc_mt19937 *g, uint32_t seed);
uint32_t orc_mt_next(orc_mt19937 *g);
double orc_mt_canonical(orc_mt19937 *g);
/* init_vel's draws: n values from a fresh std::normal_distribution(0, sigma) */
void orc_mt_normals(orc_mt19937 *g, double sigma, double *out, unsigned n);
/* init_pos (hardspheres.cc:37-138); returns 0 ok, -1 tries exceeded */
int orc_init_pos(const hmc_params *p, orc_mt19937 *g, double *pos);
/* Fill a crankshaft tape of n_words from a COPY of g (g is not advanced). */
void orc_mt_tape(const orc_mt19937 *g, uint32_t *words, double *sin_tab,
                 double *cos_tab, unsigned n_words);
void orc_mt_discard(orc_mt19937 *g, unsigned n);

/* ---- whole-program restatement of main (main.cc:362-491) --------------------
 * Runs EQ, WL and the G-test loop for one transition exactly as the reference
 * does with std::mt19937(seed_seq(seeds)).  Outputs: final s_bias[nstates],
 * number of G-test iterations, valid collision events processed.
 * max_gtest_iters > 0 caps the loop (0 = use p->max_g_test_count).
 * pos_in (optional) = starting structure instead of init_pos
 * (the --input-file path, main.cc:35-38). */
typedef struct orc_main_result {
  double s_bias[16];
  uint64_t config_count[16];
  uint64_t collisions, cell_crossings, heap_pops, predictions;
  uint32_t gtest_iters;
  int accepted;
  int err;
  double seconds_md; /* wall time inside EQ+WL+MAIN loops */
} orc_main_result;
int orc_run_main(const hmc_params *p, const double *pos_in,
                 unsigned max_gtest_iters, unsigned dist_rows_cap,
                 double *dist_out, unsigned *n_dist_rows,
                 orc_main_result *res);

/* bench.py's CPU leg: see hmc_oracle.c / ref_driver.cc:ref_bench_main_phase */
int orc_bench_main_phase(const hmc_params *p, const double *pos_in, unsigned eq_iters,
                         unsigned rounds, unsigned iters_per_round, double *seconds,
                         double *pos_final, uint64_t *config_final, uint64_t *collisions,
                         uint64_t *cell_crossings);

#ifdef __cplusplus
}
#endif
#endif
