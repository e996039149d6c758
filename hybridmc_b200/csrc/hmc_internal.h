// hmc_internal.h — structures shared by the kernels (hmc_kernels.cu) and the
// C-ABI host layer (hmc_api.cu).  Not installed; the public surface is
// include/hybridmc_b200.h.
#ifndef HMC_INTERNAL_H
#define HMC_INTERNAL_H

#include "../../include/hybridmc_b200.h"
#include <cuda_runtime.h>
#include <stdint.h>

// Per-transition constants (one hmc_params each); everything that may differ
// between the transitions of one protein.
struct TransDev {
  double rc2, stair2, p_rc2; // squares as from_json derives them (main.cc:515-533)
  uint32_t has_stair, has_p_rc, max_nbonds;
  uint32_t n_nonlocal, n_transient;
  uint32_t pad;
  // nonlocal bonds in the (i,j)-lexicographic scan order of
  // dist_between_nonlocal_beads (hardspheres.cc:1091-1103) = column order of `dist`
  uint8_t nl_i[HMC_MAX_BONDS], nl_j[HMC_MAX_BONDS];
  // transient bonds in JSON list order: bit k of the config (config.h:30-45)
  uint8_t tb_i[HMC_MAX_BONDS], tb_j[HMC_MAX_BONDS];
};

enum { HMC_RNG_PHILOX = 0, HMC_RNG_INJECTED = 1 };
enum { HMC_OP_RUN = 0, HMC_OP_MD_ONLY = 1, HMC_OP_CRANK_ONLY = 2 };

// Kernel argument block (passed by value, __grid_constant__).
struct KArgs {
  // geometry / potential shared by all transitions
  int nb, ncell, npairs, nstates;
  int list_cap; // capacity of the per-replica list of live nonlocal predictions (multiple of 8)
  uint32_t off_tnn, off_list, off_plist; // byte offsets inside a replica's shared-memory image
  double l, inv_l, lcell, inv_lcell;
  double near_min2, near_max2, nnear_min2, nnear_max2, rh2;
  // lookup tables (constant-bank indexed loads instead of select chains)
  double s_in_tab[3];   // inner radius^2 by pair class: nearest, next-nearest, nonlocal core
  double s_out_tab[3];  // outer radius^2 by pair class (nonlocal entry unused)
  double s_mid_tab[2];  // (inner^2 + outer^2) / 2 of the two local wells
  double sigma2_tab[5]; // wall radius^2 by event type 0..4
  // ncell <= 10: cells as one-hot masks, 10 bits per axis; nbr_tab[c] = bits c-1, c, c+1
  // (periodic)
  int cell_masks;
  uint16_t nbr_tab[16];
  double one;       // 1.0 (an operand the compiler cannot fold, see predict_pair)
  double m, vsigma; // vsigma = sqrt(temp/m), init_vel (hardspheres.cc:202)
  double del_t, del_t_wl, gamma0;
  uint32_t nsteps, nsteps_eq, nsteps_wl, write_step, mc_moves, mc_write;
  // tables
  const TransDev *trans;
  const uint8_t *pairinfo; // [T][nb][64] (symmetric): bits 0-6 transient bit index+1, bit 7 permanent
  const unsigned long long *bondmask; // [T][nb]: bit k set iff pairinfo of (b,k) is non-zero
  const uchar2 *pair_ij;   // [npairs] -> (i,j), i<j
  // ensemble
  int n_replicas, replicas_per_trans;
  // replica state (global memory, host layout)
  double *pos;       // [R][nb][3]
  double *vel;       // [R][nb][3]
  uint64_t *config;  // [R]
  double *s_bias;    // [R][nstates]
  uint32_t *err;     // [R]
  // accumulators: [T][nstates] counts ++ formed[T] ++ broken[T] ++ events[4]
  unsigned long long *acc;
  unsigned long long *acc_rep; // optional per-replica counts [R][nstates]
  const uint8_t *skip;         // optional [R]: replica sits this launch out
  // RNG identity of replica j of transition t: t * rng_total + rng_first + j
  uint32_t rng_first, rng_total;
  int n_trans;
  // optional samples
  double *dist;          // [R][dist_cap][n_nonlocal_max]
  uint32_t *dist_n;      // [R]
  uint32_t dist_cap, dist_stride;
  unsigned long long *cfg_int; // [R][cfg_cap]
  uint32_t *cfg_n;
  uint32_t cfg_cap;
  unsigned long long *hist; // [T][n_nonlocal_max][nbins]
  uint32_t hist_bins;
  double hist_scale; // nbins / rmax
  double *uniq_pos, *uniq_vel; // [R][nb][3]
  uint8_t *uniq_found;         // [R]
  // trace of one replica
  hmc_event *trace;
  uint32_t *trace_n;
  int trace_replica;
  uint32_t trace_cap;
  // schedule
  int op, phase;
  uint32_t first, count; // iterations (RUN), steps (MD_ONLY), moves (CRANK_ONLY)
  uint32_t move_offset;
  double t_now; // bead clocks at launch (all beads share it at step boundaries)
  // rng
  int rng_mode;
  uint32_t seed0, seed1;
  unsigned long long md_draw_base, mc_draw_base;
  const double *normals; // injected: [R][count][nb][3]
  const uint32_t *tape_words;
  const double *tape_sin, *tape_cos; // [R][tape_n]
  uint32_t tape_n;
  uint32_t *tape_consumed; // [R] words used ++ [R] moves completed
  // work distribution
  unsigned int *work_counter;
  uint32_t smem_per_replica;
  uint32_t max_events_per_step;
  int short_chain; // (nb-1)*near_max < 1.45*length: bead-bead differences stay below 1.5 boxes
  double chain_max; // (nb-1)*near_max: the fully stretched chain
  // packed single-precision constants of the partner pre-filter (read as 64-bit
  // constant-bank operands, so they occupy no registers)
  float2 pf_invl2, pf_nl2, pf_nrh2, pf_magic, pf_nmagic, pf_c1e5, pf_c1e2;
  float pf_xc, pf_l;
  int two_pass;    // partner re-prediction as pre-filter + compacted exact pass (chains >= 16 beads)
};

// launcher implemented in hmc_kernels.cu; group = lanes per replica (8/16/32)
cudaError_t hmc_launch(const KArgs &a, int group, cudaStream_t stream, int *grid_out,
                       int *block_out);
size_t hmc_smem_per_replica(int nb, int group);
int hmc_list_capacity(int nb, int group, int override_entries);
void hmc_smem_layout(int nb, int group, int cap_override, uint32_t *off_tnn, uint32_t *off_list, uint32_t *off_plist,
                     uint32_t *total);
int hmc_pick_group(int nb);
long long hmc_bounds_violations_device();

// message returned by hmc_last_error(NULL) (handle-less entry points)
void hmc_set_global_error(const char *msg);

#endif
