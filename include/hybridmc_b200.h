/*
 * hybridmc_b200.h — C-ABI of the B200-native HybridMC ensemble engine.
 *
 * This is the drop-in boundary for the reference's hot path: everything the
 * reference executes inside run_trajectory / run_trajectory_eq / wang_landau
 * (/root/reference/src/main.cc:151-305) and the functions those call
 * (src/hardspheres.cc, src/crankshaft.cc, src/entropy.cc).  The host program
 * (C++17 `hybridmc` executable, or the ctypes mirror in hybridmc_b200/) calls
 * these entry points once per phase / iteration batch — never per event.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns 0 on success, <0 on error; the message for the
 *     last error on a handle is hmc_last_error(h) (hmc_last_error(NULL) for
 *     hmc_create failures).  No exception crosses the boundary.
 *   - the caller owns all host buffers; the library owns device memory.
 *   - one handle per device; a handle is not thread-safe, distinct handles are.
 *   - all work is ordered on the handle's CUDA stream; getters synchronise.
 *   - there is NO CPU fallback: without a CUDA device hmc_create fails.
 *
 * Ensemble layout: a handle holds T "transitions" (one hmc_params each; all
 * share nbeads / box / cell geometry) times R replicas per transition.
 * Replica index r in [0, T*R) belongs to transition r / R.
 */
#ifndef HYBRIDMC_B200_H
#define HYBRIDMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMC_MAX_BONDS 64 /* Config is a uint64 bitmask (src/config.h:18) */
#define HMC_MAX_BEADS 64
#define HMC_MAX_SEEDS 8

/* POD copy of the reference's Param (src/params.h:10-101) with the bond lists
 * of NonlocalBonds (src/config.h:21-56) flattened.  Field meaning and JSON key
 * are identical to from_json(Param) (src/main.cc:500-562); squares (rh2, ...)
 * are derived inside the library exactly as there (x*x). */
typedef struct hmc_params {
  double m;         /* "m" */
  double sigma;     /* "sigma_bb" */
  double near_min, near_max;   /* nearest-neighbour well */
  double nnear_min, nnear_max; /* next-nearest well */
  double rh, rc;               /* nonlocal hard core / bond radius */
  double stair;                /* valid iff has_stair */
  double p_rc;                 /* valid iff has_p_rc */
  double length;               /* box length */
  double del_t, del_t_wl;
  double gamma, gamma_f;
  double temp;
  double pos_scale, neg_scale, sig_level;
  uint64_t tries;
  uint32_t has_stair, has_p_rc;
  uint32_t nbeads, ncell;
  uint32_t nsteps, nsteps_eq, nsteps_wl, write_step;
  uint32_t mc_moves, mc_write;
  uint32_t total_iter, total_iter_eq;
  uint32_t max_g_test_count, max_nbonds;
  uint32_t n_nonlocal, n_transient, n_permanent, n_seeds;
  uint32_t nonlocal_bonds[HMC_MAX_BONDS][2];
  uint32_t transient_bonds[HMC_MAX_BONDS][2];
  uint32_t permanent_bonds[HMC_MAX_BONDS][2];
  uint32_t seeds[HMC_MAX_SEEDS];
} hmc_params;

/* Event types = index in the reference's Event variant (src/event.h:99-102). */
enum hmc_event_type {
  HMC_EV_MIN_NEAREST = 0,
  HMC_EV_MAX_NEAREST = 1,
  HMC_EV_MIN_NNEAREST = 2,
  HMC_EV_MAX_NNEAREST = 3,
  HMC_EV_MIN_NONLOCAL_INNER = 4,
  HMC_EV_MAX_NONLOCAL_INNER = 5,
  HMC_EV_MAX_NONLOCAL_OUTER = 6,
  HMC_EV_BEAD_CELL = 7
};

/* One valid (processed) event; for HMC_EV_BEAD_CELL j is the wall
 * (CellEvent::Wall, src/event.h:86: xpos,xneg,ypos,yneg,zpos,zneg). */
typedef struct hmc_event {
  double t;
  uint32_t type;
  uint32_t i, j;
  uint32_t pad;
} hmc_event;

/* Phases of the reference's main (src/main.cc:411-491). */
enum hmc_phase {
  HMC_PHASE_EQ = 0,  /* run_trajectory_eq   main.cc:151-169 (nsteps_eq, no MC, no sampling) */
  HMC_PHASE_WL = 1,  /* run_trajectory_wl   main.cc:249-305 (nsteps_wl, del_t_wl, ±gamma) */
  HMC_PHASE_MAIN = 2 /* run_trajectory      main.cc:171-247 (sampling + crankshaft) */
};

/* Per-replica error flags (sticky; reference throws / asserts at these points). */
enum hmc_error_flag {
  HMC_ERR_LOCAL_OVERLAP = 1,    /* main.cc:48-51 "local beads overlap" */
  HMC_ERR_NONLOCAL_OVERLAP = 2, /* main.cc:53-56 "nonlocal beads overlap" */
  HMC_ERR_ENERGY = 4,           /* main.cc:219-225 "energy is not conserved" */
  HMC_ERR_NAN = 8,              /* non-finite event time */
  HMC_ERR_TAPE_UNDERRUN = 16,   /* injected crankshaft tape exhausted */
  HMC_ERR_SAMPLE_OVERFLOW = 32, /* optional sample buffer full (samples dropped) */
  HMC_ERR_LIST_OVERFLOW = 64,   /* more live nonlocal predictions than the event list holds
                                   (HMC_LIST_CAP raises the capacity); the replica's
                                   trajectory is no longer the reference's */
  HMC_ERR_CRANK_STUCK = 128     /* a crankshaft move found no valid rotation in 2^20 trials */
};

typedef struct hmc_engine *hmc_handle;

/* ---- lifecycle --------------------------------------------------------- */

/* Build an ensemble on CUDA device `device`: n_transitions parameter sets
 * (same nbeads/length/ncell), replicas_per_transition replicas each.
 * Validates what initialize_system validates (ncell >= 4, lcell >= rc or
 * stair; main.cc:58-69) and from_json's stair check (main.cc:535-537).
 * Replaces: Param construction + System/Box/Cells/EventQueue set-up
 * (main.cc:362-380, 187-190). */
int hmc_create(const hmc_params *transitions, int n_transitions,
               int replicas_per_transition, int device, hmc_handle *out);
int hmc_destroy(hmc_handle h);
const char *hmc_last_error(hmc_handle h);
/* ABI version of this header (bumped on any signature change). */
int hmc_abi_version(void);

/* ---- state in / out ------------------------------------------------------ */

/* pos[R_total][nbeads][3], unwrapped coordinates as in System::pos
 * (src/system.h:15).  Replaces init_pos / read_input results (main.cc:26-43). */
int hmc_set_positions(hmc_handle h, const double *pos);
int hmc_get_positions(hmc_handle h, double *pos);
int hmc_get_velocities(hmc_handle h, double *vel);
/* config[R_total]: bond bitmask per replica (UpdateConfig::config). */
int hmc_set_config(hmc_handle h, const uint64_t *config);
int hmc_get_config(hmc_handle h, uint64_t *config);
/* Recompute every replica's bitmask from its positions:
 * init_update_config (src/hardspheres.cc:144-171). */
int hmc_init_config_from_positions(hmc_handle h);
/* s_bias[n_transitions][nstates] broadcast to the replicas of each transition
 * (System::s_bias; init_s src/entropy.cc:14-20 gives the starting values). */
int hmc_set_s_bias(hmc_handle h, const double *s_bias, int nstates);
/* s_bias[R_total][nstates]: per replica (they differ after the WL phase). */
int hmc_get_s_bias(hmc_handle h, double *s_bias, int nstates);
/* Zero bead clocks and collision counters and the step index
 * (main.cc:414-419, 436-439, 452-457). */
int hmc_reset_clocks(hmc_handle h);
/* Zero config_count and the sample cursors (main.cc:459-461). */
int hmc_reset_counts(hmc_handle h);

/* ---- production path: device Philox4x32-10 RNG ------------------------- */

/* Philox key; replica r uses counter word 3 = r. */
int hmc_seed(hmc_handle h, uint32_t seed0, uint32_t seed1);
/* Run n_iters iterations of `phase` for every replica, starting at iteration
 * index first_iter (the reference's `iter` / `iter_wl`):
 *   EQ   : n_iters x nsteps_eq MD steps                     (main.cc:428-431)
 *   WL   : n_iters outer WL iterations (nstates trajectories each, gamma
 *          schedule of main.cc:283-304, per-replica s_bias[native])
 *   MAIN : n_iters x (nsteps MD steps + mc_moves crankshafts), sampling
 *          config_count / dist / config_int                   (main.cc:463-468)
 * Asynchronous on the handle's stream. */
int hmc_run(hmc_handle h, int phase, unsigned first_iter, unsigned n_iters);
int hmc_sync(hmc_handle h);

/* ---- reference-exact path: host-injected RNG stream ---------------------- */

/* Run `nsteps` MD steps (step indices first_step..first_step+nsteps-1,
 * step_time = step*del_t as in run_step, main.cc:104) with the Maxwell-
 * Boltzmann draws supplied by the host: normals[R_total][nsteps][nbeads][3]
 * are the values init_vel assigns BEFORE the centre-of-mass shift
 * (src/hardspheres.cc:205-209).  phase selects del_t/del_t_wl and sampling;
 * for WL each "step" is one run_trajectory_wl call (trajectory index
 * first_step + k, i.e. outer iteration (first_step+k)/nstates) including the
 * ±gamma update. */
int hmc_md_steps_injected(hmc_handle h, int phase, unsigned first_step,
                          unsigned nsteps, const double *normals);
/* Run n_moves crankshaft moves per replica (src/crankshaft.cc:224-262) from a
 * raw 32-bit word tape: words[R_total][n_words] are successive outputs of the
 * host's std::mt19937; sin_tab/cos_tab[R_total][n_words] hold sin/cos of the
 * angle that libstdc++'s uniform_real_distribution(0,2pi) would form from
 * words[c], words[c+1] (host libm, so the rotation matrix is bit-identical to
 * the reference's).  consumed[R_total] returns the words each replica used.
 * move_offset = index of the first move within the iteration (for the
 * mc_write bookkeeping of main.cc:232-236). */
int hmc_crankshaft_injected(hmc_handle h, unsigned move_offset, unsigned n_moves,
                            const uint32_t *words, const double *sin_tab,
                            const double *cos_tab, unsigned n_words,
                            uint32_t *consumed);

/* moves[R_total]: crankshaft moves completed by the last hmc_crankshaft_injected
 * call (< n_moves when the tape ran out; the unfinished move is rolled back). */
int hmc_get_moves_done(hmc_handle h, uint32_t *moves);

/* ---- results ------------------------------------------------------------- */

/* config_count[n_transitions][nstates] (pooled over replicas; only MD-step
 * samples, main.cc:204,471-473), formed/broken[n_transitions] (CountBond),
 * err_flags[R_total].  Any pointer may be NULL. */
int hmc_get_counts(hmc_handle h, uint64_t *config_count, int nstates,
                   uint64_t *formed, uint64_t *broken, uint32_t *err_flags);
/* Device pointer of the u64 block [n_transitions][nstates] ++ formed[T] ++
 * broken[T] ++ events[4] for an in-place NCCL all-reduce by the caller
 * (torch.distributed / ncclAllReduce); *n_u64 = its length. */
int hmc_counts_device_ptr(hmc_handle h, void **dptr, uint64_t *n_u64);
/* events[0] valid collision events (types 0-6), [1] cell crossings,
 * [2] crankshaft attempts, [3] crankshaft accepted — summed over replicas. */
int hmc_get_event_counts(hmc_handle h, uint64_t events[4]);

/* Optional per-replica sample buffers (off by default).  rows_per_replica
 * bounds the dist rows / config_int entries kept per replica between
 * hmc_reset_counts calls. */
int hmc_enable_samples(hmc_handle h, unsigned dist_rows_per_replica,
                       unsigned config_int_per_replica);
/* dist[R_total][rows][n_nonlocal] (dist_between_nonlocal_beads,
 * src/hardspheres.cc:1084-1104, one row per MD step), n_rows[R_total]. */
int hmc_get_dist(hmc_handle h, double *dist, uint32_t *n_rows);
/* config_int[R_total][cap] in the reference's append order (MD samples and
 * crankshaft samples interleaved, main.cc:205,234), n[R_total]. */
int hmc_get_config_int(hmc_handle h, uint64_t *config_int, uint32_t *n);
/* Fine histograms of every nonlocal-bond distance, pooled per transition and split by
 * the state of the first transient bond (index 1: its distance < rc, else 0) — the
 * sufficient statistics of tools/mfpt.py (the on/off split of mfpt.py:104-125):
 * hist[n_transitions][n_nonlocal][2][nbins] over [0, rmax), one count per MD step. */
int hmc_enable_dist_hist(hmc_handle h, unsigned nbins, double rmax);
int hmc_get_dist_hist(hmc_handle h, uint64_t *hist);
/* Device pointer of that u64 histogram block (the first-passage-time accumulators) for an
 * in-place NCCL all-reduce when one transition's replicas are spread over several GPUs;
 * *n_u64 = n_transitions * n_nonlocal * 2 * nbins.  Fails if histograms are not enabled. */
int hmc_dist_hist_device_ptr(hmc_handle h, void **dptr, uint64_t *n_u64);
/* First recorded configuration with config == 1 per replica
 * (unique_config/{pos,vel}, main.cc:211-216): pos/vel[R_total][nbeads][3],
 * found[R_total]. */
int hmc_get_unique_config(hmc_handle h, double *pos, double *vel, uint8_t *found);

/* Event trace of one replica (parity / debugging): the next hmc_run /
 * hmc_md_steps_injected calls append every valid event of `replica` until
 * `capacity` is reached.  capacity 0 disables. */
int hmc_enable_trace(hmc_handle h, int replica, unsigned capacity);
int hmc_get_trace(hmc_handle h, hmc_event *out, unsigned max_events,
                  unsigned *n_events);

/* ---- entropy update (host-side, tiny) ------------------------------------ */

/* compute_entropy (src/entropy.cc:36-69) on pooled counts, in place. */
int hmc_compute_entropy(const uint64_t *config_count, double *s_bias, int nstates,
                        double pos_scale, double neg_scale);
/* compute_g_val (src/entropy.cc:72-89). */
double hmc_compute_g_val(const uint64_t *config_count, int nstates);
/* g_test (src/entropy.cc:92-116): returns 1 accepted, 0 rejected; g_val,
 * g_crit, p_val optional outputs. */
int hmc_g_test(const uint64_t *config_count, int nstates, double sig_level,
               double *g_val, double *g_crit, double *p_val);

/* Number of outer Wang-Landau iterations the reference's gamma schedule performs
 * (while gamma > gamma_f: ...; gamma = 1/iter_wl — main.cc:283-304). */
unsigned hmc_wl_outer_iterations(const hmc_params *p);

/* init_pos (src/hardspheres.cc:37-138) for ONE chain on the host with
 * std::mt19937(std::seed_seq(seeds)) — the reference's seeding contract
 * (main.cc:379-380).  pos[nbeads][3].  Returns -1 where the reference throws
 * "number of tries exceeded". */
int hmc_init_pos_chain(const hmc_params *p, const uint32_t *seeds, unsigned n_seeds,
                       double *pos);

/* Same with a caller-owned engine: `mt19937` points to a std::mt19937 that keeps
 * its state across the call (C++ hosts only). */
int hmc_init_pos_engine(const hmc_params *p, void *mt19937, double *pos);

/* ---- whole program (host-side phase loop of main.cc:362-491) ------------------ */

typedef struct hmc_program_result {
  double s_bias[16];
  uint64_t config_count[16]; /* last G-test iteration */
  uint64_t collisions, cell_crossings;
  uint64_t config_final;
  uint32_t gtest_iters;
  int accepted;
  uint32_t err_flags;
  uint32_t pad;
  double seconds;
  double pos_final[HMC_MAX_BEADS * 3];
  char message[256];
} hmc_program_result;

/* The reference program for ONE replica with the reference's RNG contract,
 * std::mt19937(std::seed_seq(p->seeds)): init_pos (or pos_in, the --input-file
 * path), total_iter_eq equilibration iterations, Wang-Landau, then G-test
 * iterations until acceptance or max_gtest_iters (0 = p->max_g_test_count).
 * The host draws every variate with libstdc++ exactly as the reference does and
 * injects it, so s_bias / counts / dist are bit-identical to the reference's.
 * dist_out (optional): up to dist_rows_cap rows of n_nonlocal doubles. */
int hmc_run_program_exact(const hmc_params *p, int device, const double *pos_in,
                          unsigned max_gtest_iters, unsigned dist_rows_cap,
                          double *dist_out, unsigned *n_dist_rows,
                          hmc_program_result *res);

/* Output hooks for a host program that writes the reference's datasets
 * (src/main.cc:383-409,443-446,477-495).  Any pointer may be NULL. */
typedef struct hmc_sink {
  void *user;
  /* once per G-test iteration: config_count row, updated s_bias, G-test verdict
   * (main.cc:471-491; g_test prints the same numbers, entropy.cc:106-113) */
  void (*on_round)(void *user, unsigned round, const uint64_t *config_count,
                   const double *s_bias, int nstates, int accepted, double g_val,
                   double g_crit, double p_val);
  /* dist rows of that iteration, one per MD step (DistWriter, main.cc:199-200) */
  void (*on_dist)(void *user, const double *rows, unsigned nrows, unsigned nbonds);
  /* config/int entries of that iteration in append order (main.cc:205,234,245) */
  void (*on_config_int)(void *user, const uint64_t *config_int, unsigned n);
  /* unique_config/{pos,vel,config}: first recorded config == 1 (main.cc:211-216) */
  void (*on_unique)(void *user, const double *pos, const double *vel, uint64_t config,
                    unsigned nbeads);
  /* snapshot point after every G-test iteration (main.cc:480-486): textual
   * std::mt19937 state as written by operator<< (snapshot.cc:43-51) */
  void (*on_snapshot)(void *user, const double *pos, const double *s_bias, int nstates,
                      uint64_t config, const char *mt_state);
} hmc_sink;

/* Optional restart state (read_snapshot, src/snapshot.cc:70-120). */
typedef struct hmc_snapshot_in {
  const double *pos;     /* [nbeads][3] */
  const double *s_bias;  /* [nstates] */
  const char *mt_state;  /* textual std::mt19937 state */
  uint64_t config;
} hmc_snapshot_in;

/* hmc_run_program_exact with output hooks and an optional snapshot to resume from
 * (snap overrides pos_in, as initialize_pos does, main.cc:31-34). */
int hmc_run_program_exact_sink(const hmc_params *p, int device, const double *pos_in,
                               const hmc_snapshot_in *snap, unsigned max_gtest_iters,
                               const hmc_sink *sink, hmc_program_result *res);

/* Measured FP64 issue rate of `device` in TFLOP/s: DFMA chains, and DMUL+DADD
 * chains (the ceiling for this engine, which is compiled without FMA
 * contraction for bit parity).  Roofline denominator for bench.py. */
int hmc_fp64_peak(int device, double *fma_tflops, double *muladd_tflops);

/* ---- ensemble snapshot / resume ---------------------------------------------
 * Together with hmc_get_positions / hmc_get_config / hmc_get_s_bias this is the
 * ensemble form of the reference's snapshot (pos, s_bias, RNG state, config;
 * src/snapshot.cc:14-68): the schedule state below replaces the mt19937 text because
 * the Philox streams are counter-based.  Restoring all four on a fresh handle built
 * from the same parameters continues the run bit-identically. */
typedef struct hmc_schedule_state {
  double t_now;          /* bead clocks (all beads share it at a step boundary) */
  uint64_t md_draws;     /* initialize_system calls so far (Philox draw index) */
  uint64_t mc_draws;     /* crankshaft blocks so far */
  uint32_t seed0, seed1; /* Philox key */
} hmc_schedule_state;
int hmc_get_schedule_state(hmc_handle h, hmc_schedule_state *out);
int hmc_set_schedule_state(hmc_handle h, const hmc_schedule_state *in);
/* s_bias[R_total][nstates], one row per replica (counterpart of hmc_get_s_bias; the
 * per-transition broadcast of hmc_set_s_bias loses per-replica values mid-WL). */
int hmc_set_s_bias_per_replica(hmc_handle h, const double *s_bias, int nstates);

/* ---- HDF5 output -----------------------------------
 * Dependency-free writer of the HDF5 files the reference produces through H5Cpp
 * (datasets: src/writer.h, src/config.cc:34-60; root attributes: src/main.cc:397-409,
 * src/config.cc:18-27; snapshot: src/snapshot.cc:14-68).  Items are collected in memory
 * and the file is written once by hmc_h5_write (superblock v0, symbol-table groups,
 * contiguous datasets: readable by every libhdf5 / h5py).  `path` may contain '/'
 * (groups are created on demand, as H5File::createGroup in main.cc:387).  dims == NULL
 * or rank 0 is a scalar; nbytes must equal prod(dims) * element size; HMC_H5_VSTR is a
 * scalar variable-length string (data = the characters, nbytes = their count).
 * Errors: <0, message in hmc_last_error(NULL). */
typedef enum hmc_h5_type {
  HMC_H5_F64 = 0, /* H5::PredType::NATIVE_DOUBLE */
  HMC_H5_F32 = 1, /* NATIVE_FLOAT  (s_bias, main.cc:409) */
  HMC_H5_U64 = 2, /* NATIVE_UINT64 (config, config_count) */
  HMC_H5_U32 = 3, /* NATIVE_UINT   (nsteps, bond lists) */
  HMC_H5_VSTR = 4 /* H5::StrType(0, H5T_VARIABLE) (snapshot.cc:44) */
} hmc_h5_type;
typedef struct hmc_h5_file *hmc_h5;
int hmc_h5_create(hmc_h5 *out);
int hmc_h5_dataset(hmc_h5 f, const char *path, int dtype, int rank, const uint64_t *dims,
                   const void *data, uint64_t nbytes);
int hmc_h5_attribute(hmc_h5 f, const char *name, int dtype, int rank, const uint64_t *dims,
                     const void *data, uint64_t nbytes);
int hmc_h5_write(hmc_h5 f, const char *filename, int do_fsync);
int hmc_h5_destroy(hmc_h5 f);

/* ---- multi-GPU: replicas of one transition spread over several ranks ----------------
 * One process per GPU.  Each rank builds a handle with its share of every transition's
 * replicas; the only exchange of the path is an integer all-reduce of the per-state
 * counts (and of the first-passage-time histograms) once per G-test round — SURVEY.md
 * section 8(e); it replaces the reference's hmc-per-process + files-on-disk hand-off
 * (tools/init_json.py:14-36) and its planned `hmc_allreduce_counts(h, ncclComm_t)`.
 * The communicator is the library's own (NCCL loaded with dlopen at first use), so a
 * C/C++ host needs neither torch nor MPI. */
#define HMC_COMM_ID_BYTES 128
typedef struct hmc_comm_s *hmc_comm;
/* rank 0: make an id and hand its 128 bytes to the other ranks by any means */
int hmc_comm_unique_id(uint8_t id[HMC_COMM_ID_BYTES]);
int hmc_comm_create(const uint8_t id[HMC_COMM_ID_BYTES], int rank, int world, int device,
                    hmc_comm *out);
/* the same with the id passed through a file all ranks can see: rank 0 writes it (temporary
 * file + rename), the others wait for it up to 120 s; rank 0 removes it in hmc_comm_destroy.
 * `path` must not exist when the ranks start (a leftover of a crashed launch would be read
 * as this launch's id): use a fresh name per launch. */
int hmc_comm_create_from_file(const char *path, int rank, int world, int device, hmc_comm *out);
/* adopt an ncclComm_t the host already owns (it stays the host's; same libnccl.so.2) */
int hmc_comm_wrap_nccl(void *nccl_comm, int rank, int world, int device, hmc_comm *out);
int hmc_comm_destroy(hmc_comm c);
int hmc_comm_rank(hmc_comm c);
int hmc_comm_world(hmc_comm c);
int hmc_comm_nccl_version(void);
/* collectives issued through this communicator so far */
uint64_t hmc_comm_collectives(hmc_comm c);
/* Sum config_count[T][nstates] over all ranks, in place in the accumulator block,
 * ordered on the handle's stream (asynchronous; the getters synchronise) — once per
 * G-test round, after the round's hmc_run and before hmc_get_counts (the counts are
 * zeroed by hmc_reset_counts at the start of every round).  u64 sums: exact and
 * order-independent, so every rank then computes the same s_bias update redundantly.
 * A NULL or 1-rank communicator is a no-op. */
int hmc_allreduce_counts(hmc_handle h, hmc_comm c);
/* The running totals behind it (formed[T] ++ broken[T] ++ events[4]) and the distance
 * histograms (hmc_enable_dist_hist) accumulate over the rounds: sum them ONCE, when the
 * loop has ended (summing an already summed block again would count it world times). */
int hmc_allreduce_totals(hmc_handle h, hmc_comm c);
int hmc_allreduce_hist(hmc_handle h, hmc_comm c);
/* Host-buffer helpers for the O(states) bookkeeping around it (blocking):
 * op 0 sum, 1 max, 2 min. */
int hmc_comm_allreduce_host_u64(hmc_comm c, uint64_t *buf, uint64_t n, int op);
int hmc_comm_bcast_host(hmc_comm c, void *buf, uint64_t nbytes, int root);
int hmc_comm_allreduce_device_u64(hmc_comm c, void *dptr, uint64_t n, void *cuda_stream);
/* RNG identity of the handle's replicas: replica j of transition t draws the Philox
 * streams of replica  t * total_per_transition + first + j  of the whole (multi-rank)
 * ensemble, so a split run reproduces the one-rank run bit for bit.  Default: local
 * indices (first = 0, total = replicas_per_transition). */
int hmc_set_replica_ids(hmc_handle h, unsigned first, unsigned total_per_transition);

/* Capacity (entries) of a replica's list of live nonlocal predictions in shared memory.
 * The default keeps 16 warps per SM resident and covers open chains with a wide margin; a
 * compact state (most bonds of a protein formed) can hold more, which raises
 * HMC_ERR_LIST_OVERFLOW: run it again with a larger list (fewer replicas per SM).  Every
 * nonlocal pair plus two beads' worth is the most a list can hold; larger values are clipped.
 * 0: default.  Takes effect at the next launch (the list lives only within a launch). */
int hmc_set_list_capacity(hmc_handle h, unsigned entries);
unsigned hmc_get_list_capacity(hmc_handle h);

/* ---- independent-program mode ----------------------------------------------------------
 * Every replica keeps its own s_bias and its own config_count, i.e. runs the reference's
 * whole program like one CPU process with its own seed (R replicas = R reference seeds
 * at equal sample counts; the statistical-parity tests use it). */
/* per-replica counts config_count[R_total][nstates] (zeroed by hmc_reset_counts) */
int hmc_enable_replica_counts(hmc_handle h, int on);
int hmc_get_replica_counts(hmc_handle h, uint64_t *config_count);
/* skip[R_total] != 0: the replica sits out the following hmc_run calls (its program has
 * finished).  NULL clears the mask. */
int hmc_set_replica_skip(hmc_handle h, const uint8_t *skip);

/* ---- the ensemble program (production path) ---------------------------------------------
 * The phase loop of the reference's main (src/main.cc:411-491: equilibrate ->
 * Wang-Landau -> G-test rounds) for T transitions of one lattice layer, `replicas`
 * replicas each, on device Philox streams.  Pooled mode: the counts of a transition's
 * replicas are summed (over all ranks of `comm`) before compute_entropy / g_test, so one
 * round of `iters_per_round` iterations samples replicas * iters_per_round * nsteps /
 * write_step configurations.  Independent mode: see above.  Everything but O(T * nstates)
 * arithmetic and the output hooks runs on the GPU. */
typedef struct hmc_ensemble_options {
  uint32_t replicas;        /* per transition, on THIS rank */
  uint32_t iters_per_round; /* 0: ceil(total_iter / (replicas * world)), at least 1 */
  uint32_t eq_iters;        /* 0: min(total_iter_eq, 8) */
  uint32_t max_rounds;      /* 0: max_g_test_count */
  uint32_t wang_landau;     /* 1: per-replica WL pre-estimate, pooled by its exact mean */
  uint32_t independent;     /* 1: independent-program mode (no pooling, no comm) */
  uint32_t hist_bins;       /* >0: pooled distance histograms over [0, hist_rmax) */
  uint32_t dist_rows;       /* >0: keep this many raw dist rows per transition and round
                               (replica-major subsample) for the `dist` dataset */
  uint32_t config_int_rows; /* >0: likewise config/int entries */
  uint32_t seed0, seed1;    /* Philox key */
  uint32_t wl_replicas;     /* wang_landau = 1: how many replicas of a transition (over all
                               ranks, spread evenly: global id % stride == 0) run the
                               Wang-Landau phase; the others wait.  0: all of them.  The
                               phase is 1/gamma_f sequential steps per replica whatever the
                               ensemble size, so in a layer of many transitions a few
                               replicas per transition give the start value at a fraction
                               of the device time */
  double hist_rmax;
  hmc_comm comm;            /* NULL: one rank */
  uint32_t list_capacity;   /* hmc_set_list_capacity; 0: default */
  uint32_t pad2;
  const double *start_s_bias; /* NULL: init_s (entropy.cc:14-20); else [T][16] start values
                               of the flat-histogram iteration, e.g. the converged value of
                               the same bond one lattice layer earlier (fewer rounds) */
} hmc_ensemble_options;

typedef struct hmc_ensemble_result { /* one per transition */
  double s_bias[16];
  uint64_t config_count[16]; /* of the accepting (or last) round */
  uint64_t formed, broken;
  uint32_t rounds;
  int accepted;
  uint32_t err_flags;        /* OR over the transition's replicas (all ranks) */
  int have_unique;
  double unique_pos[HMC_MAX_BEADS * 3]; /* first recorded config == 1 (lowest replica id) */
  double unique_vel[HMC_MAX_BEADS * 3];
} hmc_ensemble_result;

typedef struct hmc_ensemble_totals {
  uint64_t collisions, cell_crossings; /* all ranks */
  double seconds, seconds_device;      /* wall time of the loop / kernel time of this rank */
  uint64_t collectives;
  char message[256];
} hmc_ensemble_totals;

typedef struct hmc_ensemble_sink {
  void *user;
  /* after every round, for every transition that was still running */
  void (*on_round)(void *user, int transition, unsigned round, const uint64_t *config_count,
                   const double *s_bias, int nstates, int accepted, double g_val, double g_crit,
                   double p_val);
  /* raw samples of the round (options dist_rows / config_int_rows), this rank's replicas */
  void (*on_dist)(void *user, int transition, const double *rows, unsigned nrows, unsigned nbonds);
  void (*on_config_int)(void *user, int transition, const uint64_t *config_int, unsigned n);
  /* once, at the end: pooled histograms hist[n_nonlocal][2][nbins] of the transition */
  void (*on_hist)(void *user, int transition, const uint64_t *hist, unsigned n_nonlocal,
                  unsigned nbins);
  /* independent mode, at the end: s_bias[replicas][nstates], rounds[replicas],
   * accepted[replicas] of the transition */
  void (*on_replicas)(void *user, int transition, const double *s_bias, const uint32_t *rounds,
                      const uint8_t *accepted, unsigned replicas, int nstates);
} hmc_ensemble_sink;

/* start_pos[T][nbeads][3]: one start structure per transition (init_pos or the previous
 * layer's unique_config/pos).  res[T].  Returns 0, or <0 with totals->message set; a
 * replica error flag (reference: uncaught exception) stops the loop after the round in
 * which it was raised and is reported in res[t].err_flags with return value -8. */
int hmc_run_program_ensemble(const hmc_params *transitions, int n_transitions, int device,
                             const double *start_pos, const hmc_ensemble_options *opt,
                             const hmc_ensemble_sink *sink, hmc_ensemble_result *res,
                             hmc_ensemble_totals *totals);

/* ---- launch statistics (for bench.py) ------------------------------------ */

/* kernels launched by this handle so far, and the device time (ms, CUDA
 * events on the handle's stream) of the last hmc_run / injected call. */
int hmc_get_launch_stats(hmc_handle h, uint64_t *n_launches, float *last_ms);
/* Replicas of an nbeads-bead chain that run concurrently on `device` (SMs x resident warps
 * x replicas per warp): ensembles smaller than this leave SMs idle, so a campaign tops thin
 * lattice layers up to it.  <0: no such device. */
int hmc_resident_replicas(unsigned nbeads, int device);
/* Debug builds (-DHMC_DEBUG_BOUNDS, libhybridmc_b200_bounds.so): out-of-range indices into
 * the shared-memory replica image counted so far on the current device.  -1: this library is
 * not such a build. */
long long hmc_debug_bounds_violations(void);
/* Interval timer on the handle's stream (CUDA events): everything enqueued between the two
 * calls — kernels, the NCCL all-reduce of hmc_allreduce_counts, copies — and the host gaps
 * between them.  hmc_timer_stop synchronises and returns the milliseconds. */
int hmc_timer_start(hmc_handle h);
int hmc_timer_stop(hmc_handle h, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* HYBRIDMC_B200_H */
