// TEST INFRASTRUCTURE (oracle/), NOT PRODUCT CODE.
//
// extern "C" driver around the UNMODIFIED reference sources
// (/root/reference/src/{hardspheres,crankshaft,config,entropy,main}.cc, compiled
// by oracle/Makefile into oracle/_ref/libhmc_ref.so).  It lets the tests
//   * replay initialize_system / run_step with host-injected velocities and log
//     every valid event ("the reference fed the same RNG stream"),
//   * run the reference's own run_trajectory_eq / wang_landau / run_trajectory /
//     crankshaft / compute_entropy / g_test with std::mt19937(seed_seq(seeds)),
//   * expose libstdc++'s RNG so the C oracle's restatement can be pinned.
// Nothing here re-implements reference arithmetic: the bodies only sequence
// calls into reference functions the way main.cc does.
#include "crankshaft.h"
#include "entropy.h"
#include "hardspheres.h"
#include "json.hpp"
#include "snapshot.h"
#include "writer.h"

#include "../include/hybridmc_b200.h"

#include <chrono>
#include <cstring>
#include <set>
#include <sstream>

// non-static functions defined in the reference's main.cc
void from_json(const nlohmann::json &json, Param &p);
void run_trajectory_eq(System &sys, Random &mt, const Param &p, const Box &box,
                       UpdateConfig &update_config, CountBond &count_bond,
                       double wall_time, unsigned int iter);
void run_trajectory(System &sys, Random &mt, const Param &p, const Box &box,
                    std::vector<double> &dist, UpdateConfig &update_config,
                    UpdateConfigWriter &update_config_writer, PosWriter &pos_writer,
                    VelWriter &vel_writer, ConfigWriter &config_writer,
                    DistWriter &dist_writer, std::set<Config> &store_config,
                    ConfigInt &store_config_int, CountBond &count_bond,
                    double wall_time, unsigned int iter);
void wang_landau(System &sys, Random &mt, const Param &p, const Box &box,
                 UpdateConfig &update_config, CountBond &count_bond,
                 const unsigned int nstates, std::vector<double> &s_bias);

namespace {

nlohmann::json bonds_json(const uint32_t (*b)[2], unsigned n) {
  nlohmann::json a = nlohmann::json::array();
  for (unsigned k = 0; k < n; k++) a.push_back({b[k][0], b[k][1]});
  return a;
}

Param to_param(const hmc_params &h) {
  nlohmann::json j;
  j["m"] = h.m;
  j["sigma_bb"] = h.sigma;
  j["near_min"] = h.near_min;
  j["near_max"] = h.near_max;
  j["nnear_min"] = h.nnear_min;
  j["nnear_max"] = h.nnear_max;
  j["rh"] = h.rh;
  j["rc"] = h.rc;
  if (h.has_stair) j["stair"] = h.stair;
  if (h.has_p_rc) j["p_rc"] = h.p_rc;
  j["nonlocal_bonds"] = bonds_json(h.nonlocal_bonds, h.n_nonlocal);
  j["transient_bonds"] = bonds_json(h.transient_bonds, h.n_transient);
  j["permanent_bonds"] = bonds_json(h.permanent_bonds, h.n_permanent);
  j["tries"] = h.tries;
  j["nbeads"] = h.nbeads;
  j["length"] = h.length;
  j["ncell"] = h.ncell;
  j["nsteps"] = h.nsteps;
  j["nsteps_eq"] = h.nsteps_eq;
  j["del_t"] = h.del_t;
  j["nsteps_wl"] = h.nsteps_wl;
  j["del_t_wl"] = h.del_t_wl;
  j["gamma"] = h.gamma;
  j["gamma_f"] = h.gamma_f;
  j["write_step"] = h.write_step;
  std::vector<unsigned> seeds(h.seeds, h.seeds + h.n_seeds);
  j["seeds"] = seeds;
  j["temp"] = h.temp;
  j["mc_moves"] = h.mc_moves;
  j["mc_write"] = h.mc_write;
  j["total_iter"] = h.total_iter;
  j["total_iter_eq"] = h.total_iter_eq;
  j["pos_scale"] = h.pos_scale;
  j["neg_scale"] = h.neg_scale;
  j["sig_level"] = h.sig_level;
  j["max_nbonds"] = h.max_nbonds;
  j["max_g_test_count"] = h.max_g_test_count;
  Param p = j;
  return p;
}

} // namespace

struct ref_sim {
  Param p;
  Box box;
  System sys;
  UpdateConfig cfg;
  CountBond cb{};
  Random mt;
  unsigned nstates;
  std::vector<hmc_event> trace;
  unsigned trace_cap = 0;
  uint64_t n_coll = 0, n_cell = 0, n_pop = 0;
  ref_sim(const hmc_params &h)
      : p(to_param(h)), box(p.length), sys(p.nbeads),
        nstates(1u << p.transient_bonds.get_nbonds()) {
    sys.s_bias.assign(nstates, 0.0);
  }
};

namespace {

struct TraceVisitor {
  ref_sim *s;
  template <class Ev> void log(const Ev &ev, unsigned type, unsigned j) {
    if (type == HMC_EV_BEAD_CELL)
      s->n_cell++;
    else
      s->n_coll++;
    if (s->trace.size() < s->trace_cap)
      s->trace.push_back(hmc_event{ev.t, type, ev.i, j, 0});
  }
};

// run_step's loop (main.cc:107-132) with a hook on valid events
void traced_run_step(ref_sim *s, Cells &cells, EventQueue &q, double step_time) {
  TraceVisitor tv{s};
  while (!q.empty()) {
    const Event event = q.top();
    if (std::visit([=](auto &&ev) { return ev.t > step_time; }, event)) break;
    q.pop();
    s->n_pop++;
    std::visit(
        [&](auto &&ev) {
          const bool valid =
              process_event(ev, s->sys, s->p, s->box, q, cells, s->cfg, s->cb);
          if (valid) {
            using T = std::decay_t<decltype(ev)>;
            if constexpr (std::is_same_v<T, BeadCellEvent>)
              tv.log(ev, HMC_EV_BEAD_CELL, unsigned(ev.wall));
            else
              tv.log(ev, unsigned(event.index()), ev.j);
          }
        },
        event);
  }
  for (unsigned i = 0; i < s->p.nbeads; i++)
    update_pos(s->sys.pos[i], s->sys.vel[i], s->sys.times[i], step_time);
}

// initialize_system (main.cc:45-95) with init_vel's draws replaced by `normals`
int injected_initialize(ref_sim *s, Cells &cells, EventQueue &q, const double *normals) {
  int err = 0;
  const Param &p = s->p;
  if (!check_local_dist(s->sys.pos, s->box, p.near_min2, p.near_max2, p.nnear_min2,
                        p.nnear_max2))
    err |= HMC_ERR_LOCAL_OVERLAP;
  if (!check_nonlocal_dist(s->sys.pos, s->box, p.rc2, p.rh2, p.stair2, p.p_rc2,
                           p.transient_bonds, p.permanent_bonds))
    err |= HMC_ERR_NONLOCAL_OVERLAP;
  init_cells(s->sys.pos, s->box, cells);
  for (unsigned i = 0; i < p.nbeads; i++)
    s->sys.vel[i] = Vec3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
  shift_vel_CM_to_0(s->sys.vel);
  init_cell_events(s->sys.pos, s->sys.vel, p.nbeads, s->box, s->sys.counter, q,
                   s->sys.times, cells);
  init_nearest_bond_events(s->sys.pos, s->sys.vel, p.nbeads, s->box, s->sys.counter, q,
                           s->sys.times, p.near_min2, p.near_max2);
  init_nnearest_bond_events(s->sys.pos, s->sys.vel, p.nbeads, s->box, s->sys.counter, q,
                            s->sys.times, p.nnear_min2, p.nnear_max2);
  add_events_for_all_beads(s->sys.pos, s->sys.vel, p.nbeads, p.rh2, p.rc2, p.stair2,
                           p.p_rc2, s->box, s->sys.counter, q, s->sys.times, cells,
                           p.transient_bonds, p.permanent_bonds, s->cfg, p.max_nbonds);
  return err;
}

} // namespace

extern "C" {

ref_sim *ref_create(const hmc_params *h) {
  try {
    return new ref_sim(*h);
  } catch (...) {
    return nullptr;
  }
}
void ref_destroy(ref_sim *s) { delete s; }

void ref_set_pos(ref_sim *s, const double *pos) {
  for (unsigned i = 0; i < s->p.nbeads; i++)
    s->sys.pos[i] = Vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
}
void ref_get_pos(const ref_sim *s, double *pos) {
  std::memcpy(pos, s->sys.pos.data(), sizeof(double) * 3 * s->p.nbeads);
}
void ref_get_vel(const ref_sim *s, double *vel) {
  std::memcpy(vel, s->sys.vel.data(), sizeof(double) * 3 * s->p.nbeads);
}
void ref_set_config(ref_sim *s, uint64_t c) { s->cfg.config = c; }
uint64_t ref_get_config(const ref_sim *s) { return s->cfg.config; }
void ref_init_config_from_pos(ref_sim *s) {
  init_update_config(s->sys.pos, s->cfg, s->box, s->p.rc2, s->p.transient_bonds);
}
void ref_set_s_bias(ref_sim *s, const double *sb, int n) {
  for (int i = 0; i < n && i < int(s->nstates); i++) s->sys.s_bias[i] = sb[i];
}
void ref_get_s_bias(const ref_sim *s, double *sb, int n) {
  for (int i = 0; i < n && i < int(s->nstates); i++) sb[i] = s->sys.s_bias[i];
}
void ref_init_s(ref_sim *s) {
  s->sys.s_bias.clear();
  init_s(s->sys.s_bias, s->p.transient_bonds.get_nbonds());
}
void ref_reset_clocks(ref_sim *s) {
  for (unsigned i = 0; i < s->p.nbeads; i++) {
    s->sys.times[i] = 0.0;
    s->sys.counter[i] = 0;
  }
}
void ref_enable_trace(ref_sim *s, unsigned cap) {
  s->trace.clear();
  s->trace.reserve(cap);
  s->trace_cap = cap;
}
unsigned ref_get_trace(const ref_sim *s, hmc_event *out, unsigned max_events) {
  unsigned n = unsigned(s->trace.size());
  if (n > max_events) n = max_events;
  if (out) std::memcpy(out, s->trace.data(), n * sizeof(hmc_event));
  return unsigned(s->n_coll + s->n_cell);
}
void ref_get_event_counts(const ref_sim *s, uint64_t ev[6]) {
  ev[0] = s->n_coll;
  ev[1] = s->n_cell;
  ev[2] = ev[3] = 0;
  ev[4] = s->n_pop;
  ev[5] = 0;
}
void ref_get_bond_counts(const ref_sim *s, uint64_t *formed, uint64_t *broken) {
  *formed = s->cb.formed;
  *broken = s->cb.broken;
}

// same contract as orc_md_step (hmc_oracle.h)
int ref_md_step(ref_sim *s, int phase, unsigned step, const double *normals) {
  const Param &p = s->p;
  EventQueue q;
  Cells cells{p.ncell, p.length / p.ncell};
  int err = injected_initialize(s, cells, q, normals);
  if (phase == HMC_PHASE_EQ) {
    traced_run_step(s, cells, q, step * p.del_t);
    return err;
  }
  const double e0 = compute_hamiltonian(s->sys.vel, s->sys.s_bias, s->cfg.config, p.m);
  if (phase == HMC_PHASE_WL) {
    const unsigned iter_wl = step / s->nstates;
    for (unsigned st = iter_wl * p.nsteps_wl; st < (iter_wl + 1) * p.nsteps_wl; st++) {
      traced_run_step(s, cells, q, st * p.del_t_wl);
      const double e1 = compute_hamiltonian(s->sys.vel, s->sys.s_bias, s->cfg.config, p.m);
      if (!(std::abs(1 - (e1 / e0)) < 1e-6)) err |= HMC_ERR_ENERGY;
    }
    const double gamma = iter_wl == 0 ? p.gamma : 1.0 / double(iter_wl);
    if (s->cfg.config == 0)
      s->sys.s_bias[s->nstates - 1] -= gamma;
    else
      s->sys.s_bias[s->nstates - 1] += gamma;
    return err;
  }
  traced_run_step(s, cells, q, step * p.del_t);
  const double e1 = compute_hamiltonian(s->sys.vel, s->sys.s_bias, s->cfg.config, p.m);
  if (!(std::abs(1 - (e1 / e0)) < 1e-6)) err |= HMC_ERR_ENERGY;
  return err;
}

void ref_dist(const ref_sim *s, double *out) {
  std::vector<double> d;
  dist_between_nonlocal_beads(s->sys.pos, s->box, s->p.nonlocal_bonds, d);
  std::memcpy(out, d.data(), d.size() * sizeof(double));
}

// ---- libstdc++ RNG, for pinning the oracle's restatement ------------------
void ref_mt_seed_seq(ref_sim *s, const uint32_t *seeds, unsigned n) {
  std::seed_seq seq(seeds, seeds + n);
  s->mt.seed(seq);
}
void ref_mt_seed(ref_sim *s, uint32_t seed) { s->mt.seed(seed); }
uint32_t ref_mt_next(ref_sim *s) { return uint32_t(s->mt()); }
void ref_mt_discard(ref_sim *s, unsigned n) { s->mt.discard(n); }
void ref_mt_normals(ref_sim *s, double sigma, double *out, unsigned n) {
  std::normal_distribution<> g{0.0, sigma};
  for (unsigned k = 0; k < n; k++) out[k] = g(s->mt);
}
void ref_mt_tape(const ref_sim *s, uint32_t *words, double *sin_tab, double *cos_tab,
                 unsigned n) {
  Random c = s->mt;
  for (unsigned k = 0; k < n; k++) words[k] = uint32_t(c());
  Random base = s->mt;
  for (unsigned k = 0; k < n; k++) {
    if (k + 1 < n) {
      Random at = base;
      std::uniform_real_distribution<> theta(0.0, 2 * M_PI);
      const double th = theta(at);
      sin_tab[k] = std::sin(th);
      cos_tab[k] = std::cos(th);
    } else {
      sin_tab[k] = 0.0;
      cos_tab[k] = 1.0;
    }
    base();
  }
}
// init_pos with the sim's engine (hardspheres.cc:37-138)
int ref_init_pos(ref_sim *s) {
  try {
    init_pos(s->sys.pos, s->box, s->mt, s->p);
  } catch (...) {
    return -1;
  }
  return 0;
}
// init_vel + COM shift with the sim's engine (hardspheres.cc:197-212)
void ref_init_vel(ref_sim *s) { init_vel(s->sys.vel, s->mt, s->p.temp, s->p.m); }

// n_moves of the reference's crankshaft with the sim's engine; returns words used
unsigned ref_crankshaft(ref_sim *s, unsigned n_moves) {
  Random before = s->mt;
  const Param &p = s->p;
  for (unsigned i = 0; i < n_moves; i++)
    crankshaft(s->sys.pos, s->cfg, s->box, p.near_min2, p.near_max2, p.nnear_min2,
               p.nnear_max2, p.rc2, p.rh2, p.stair2, p.p_rc2, p.transient_bonds,
               p.permanent_bonds, s->mt, s->sys.s_bias);
  unsigned used = 0;
  while (!(before == s->mt) && used < (1u << 24)) {
    before();
    used++;
  }
  return used;
}

// ---- primitives for known-answer cross-checks ------------------------------
int ref_t_until_inner_coll(double x, double y, double z, double vx, double vy, double vz,
                           double sigma2, double *t) {
  return t_until_inner_coll(x, y, z, vx, vy, vz, sigma2, *t) ? 1 : 0;
}
void ref_t_until_outer_coll(double x, double y, double z, double vx, double vy, double vz,
                            double sigma2, double *t) {
  t_until_outer_coll(x, y, z, vx, vy, vz, sigma2, *t);
}
int ref_v_after_coll(double x, double y, double z, double sigma2, double *vi, double *vj,
                     int has_dS, double dS, double m) {
  std::optional<double> o;
  if (has_dS) o = dS;
  return v_after_coll(x, y, z, sigma2, vi[0], vi[1], vi[2], vj[0], vj[1], vj[2], o, m) ? 1 : 0;
}
int ref_t_until_cell(double x, double y, double z, double l, unsigned ncell, double lcell,
                     double vx, double vy, double vz, unsigned ixo, unsigned iyo,
                     unsigned izo, double *dt, unsigned *wall, unsigned *cell_new) {
  Box box{l};
  Cells cells{ncell, lcell};
  BeadCellEvent::Wall w = BeadCellEvent::xpos;
  const bool c = t_until_cell(x, y, z, box, cells, vx, vy, vz, ixo, iyo, izo, *dt, w,
                              cell_new[0], cell_new[1], cell_new[2]);
  *wall = unsigned(w);
  return c ? 1 : 0;
}
void ref_compute_entropy(const uint64_t *cc, double *sb, int n, double pos_scale,
                         double neg_scale) {
  std::vector<uint64_t> c(cc, cc + n);
  std::vector<double> s(sb, sb + n);
  std::stringstream sink;
  auto *old = std::cout.rdbuf(sink.rdbuf());
  compute_entropy(c, s, n, pos_scale, neg_scale);
  std::cout.rdbuf(old);
  std::memcpy(sb, s.data(), sizeof(double) * n);
}
double ref_compute_g_val(const uint64_t *cc, int n) {
  std::vector<uint64_t> c(cc, cc + n);
  return compute_g_val(c, n);
}
int ref_g_test(const uint64_t *cc, int n, double sig) {
  std::vector<uint64_t> c(cc, cc + n);
  std::stringstream sink;
  auto *old = std::cout.rdbuf(sink.rdbuf());
  const bool ok = g_test(c, n, sig);
  std::cout.rdbuf(old);
  return ok ? 1 : 0;
}

// ---- whole program: main.cc:362-491 sequenced over the reference's own
// run_trajectory_eq / wang_landau / run_trajectory (file output = no-op stubs)
struct ref_main_result {
  double s_bias[16];
  uint64_t config_count[16];
  uint32_t gtest_iters;
  int accepted;
  int err;
  double seconds_md;
  double pos_final[HMC_MAX_BEADS * 3];
  uint64_t config_final;
};

int ref_run_main(const hmc_params *h, const double *pos_in, unsigned max_gtest_iters,
                 ref_main_result *res) {
  std::memset(res, 0, sizeof *res);
  std::stringstream sink;
  auto *old = std::cout.rdbuf(sink.rdbuf());
  int rc = 0;
  try {
    const Param p = to_param(*h);
    Param p_eq = p;
    const Box box{p.length};
    UpdateConfig update_config;
    const unsigned t_bonds = p.transient_bonds.get_nbonds();
    const unsigned nbonds = p.nonlocal_bonds.get_nbonds();
    const unsigned nstates = std::pow(2, t_bonds);
    ConfigInt store_config_int;
    std::vector<uint64_t> config_count(nstates);
    std::vector<double> dist(nbonds);
    CountBond count_bond = {};
    System sys(p.nbeads);
    std::seed_seq seq(p.seeds.begin(), p.seeds.end());
    Random mt(seq);
    H5::H5File file("unused", H5F_ACC_TRUNC);
    UpdateConfigWriter update_config_writer(file);
    H5::Group group{file.createGroup("unique_config")};
    PosWriter pos_writer{group, "pos", p.nbeads};
    VelWriter vel_writer{group, "vel", p.nbeads};
    ConfigWriter config_writer{group, "config"};
    DistWriter dist_writer{file, "dist", nbonds};
    std::set<Config> store_config;

    for (unsigned i = 0; i < p.nbeads; i++) {
      sys.times[i] = 0.0;
      sys.counter[i] = 0.0;
    }
    double wall_time = 0.0;
    UpdateConfig update_config_eq;
    p_eq.nsteps = p.nsteps_eq;
    p_eq.total_iter = p.total_iter_eq;
    // initialize_pos (main.cc:26-43)
    if (pos_in) {
      for (unsigned i = 0; i < p.nbeads; i++)
        sys.pos[i] = Vec3(pos_in[3 * i], pos_in[3 * i + 1], pos_in[3 * i + 2]);
      init_update_config(sys.pos, update_config_eq, box, p.rc2, p.transient_bonds);
      init_s(sys.s_bias, t_bonds);
    } else {
      init_pos(sys.pos, box, mt, p_eq);
      init_s(sys.s_bias, t_bonds);
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned iter = 0; iter < p_eq.total_iter; iter++)
      run_trajectory_eq(sys, mt, p_eq, box, update_config_eq, count_bond, wall_time, iter);
    init_update_config(sys.pos, update_config, box, p.rc2, p.transient_bonds);
    for (unsigned i = 0; i < p.nbeads; i++) {
      sys.times[i] = 0.0;
      sys.counter[i] = 0.0;
    }
    wang_landau(sys, mt, p, box, update_config, count_bond, nstates, sys.s_bias);
    unsigned g_test_count = 0;
    const unsigned cap = max_gtest_iters ? max_gtest_iters : p.max_g_test_count;
    bool accepted = false;
    do {
      for (unsigned i = 0; i < p.nbeads; i++) {
        sys.times[i] = 0.0;
        sys.counter[i] = 0.0;
      }
      wall_time = 0.0;
      store_config_int.clear();
      std::fill(config_count.begin(), config_count.end(), 0);
      for (unsigned iter = 0; iter < p.total_iter; iter++)
        run_trajectory(sys, mt, p, box, dist, update_config, update_config_writer,
                       pos_writer, vel_writer, config_writer, dist_writer, store_config,
                       store_config_int, count_bond, wall_time, iter);
      for (unsigned i = 0; i < store_config_int.size(); i++)
        config_count[store_config_int[i]]++;
      compute_entropy(config_count, sys.s_bias, nstates, p.pos_scale, p.neg_scale);
    } while (++g_test_count < cap &&
             !(accepted = g_test(config_count, nstates, p.sig_level)));
    res->seconds_md =
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (unsigned i = 0; i < nstates && i < 16; i++) {
      res->s_bias[i] = sys.s_bias[i];
      res->config_count[i] = config_count[i];
    }
    res->gtest_iters = g_test_count;
    res->accepted = accepted ? 1 : 0;
    std::memcpy(res->pos_final, sys.pos.data(), sizeof(double) * 3 * p.nbeads);
    res->config_final = update_config.config;
  } catch (const std::exception &e) {
    res->err = 1;
    rc = -1;
  }
  std::cout.rdbuf(old);
  return rc;
}

// CPU-baseline leg of bench.py (same workload as the GPU arm): the reference's OWN
// run_trajectory (main.cc:171-247: nsteps MD intervals with fresh velocities, mc_moves
// crankshaft moves, sampling) called `rounds` x `iters_per_round` times from a given start
// structure, compute_entropy after every round as the G-test loop does (main.cc:451-491).
// `eq_iters` iterations of run_trajectory_eq come first, untimed.  Only calls into
// unmodified reference functions are timed; the reference has no event counter, so the
// number of events of this (deterministic) run comes from the oracle's bit-identical
// replay orc_bench_main_phase (tests/test_oracle_vs_ref.py pins the two on pos_final).
int ref_bench_main_phase(const hmc_params *h, const double *pos_in, unsigned eq_iters,
                         unsigned rounds, unsigned iters_per_round, double *seconds,
                         double *pos_final, uint64_t *config_final) {
  std::stringstream sink;
  auto *old = std::cout.rdbuf(sink.rdbuf());
  int rc = 0;
  try {
    const Param p = to_param(*h);
    Param p_eq = p;
    const Box box{p.length};
    UpdateConfig update_config;
    const unsigned t_bonds = p.transient_bonds.get_nbonds();
    const unsigned nbonds = p.nonlocal_bonds.get_nbonds();
    const unsigned nstates = std::pow(2, t_bonds);
    ConfigInt store_config_int;
    std::vector<uint64_t> config_count(nstates);
    std::vector<double> dist(nbonds);
    CountBond count_bond = {};
    System sys(p.nbeads);
    std::seed_seq seq(p.seeds.begin(), p.seeds.end());
    Random mt(seq);
    H5::H5File file("unused", H5F_ACC_TRUNC);
    UpdateConfigWriter update_config_writer(file);
    H5::Group group{file.createGroup("unique_config")};
    PosWriter pos_writer{group, "pos", p.nbeads};
    VelWriter vel_writer{group, "vel", p.nbeads};
    ConfigWriter config_writer{group, "config"};
    DistWriter dist_writer{file, "dist", nbonds};
    std::set<Config> store_config;
    for (unsigned i = 0; i < p.nbeads; i++) {
      sys.times[i] = 0.0;
      sys.counter[i] = 0.0;
    }
    double wall_time = 0.0;
    UpdateConfig update_config_eq;
    p_eq.nsteps = p.nsteps_eq;
    p_eq.total_iter = eq_iters;
    for (unsigned i = 0; i < p.nbeads; i++)
      sys.pos[i] = Vec3(pos_in[3 * i], pos_in[3 * i + 1], pos_in[3 * i + 2]);
    init_update_config(sys.pos, update_config_eq, box, p.rc2, p.transient_bonds);
    init_s(sys.s_bias, t_bonds);
    for (unsigned iter = 0; iter < eq_iters; iter++)
      run_trajectory_eq(sys, mt, p_eq, box, update_config_eq, count_bond, wall_time, iter);
    init_update_config(sys.pos, update_config, box, p.rc2, p.transient_bonds);
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned r = 0; r < rounds; r++) {
      for (unsigned i = 0; i < p.nbeads; i++) {
        sys.times[i] = 0.0;
        sys.counter[i] = 0.0;
      }
      wall_time = 0.0;
      store_config_int.clear();
      std::fill(config_count.begin(), config_count.end(), 0);
      for (unsigned iter = 0; iter < iters_per_round; iter++)
        run_trajectory(sys, mt, p, box, dist, update_config, update_config_writer,
                       pos_writer, vel_writer, config_writer, dist_writer, store_config,
                       store_config_int, count_bond, wall_time, iter);
      for (unsigned i = 0; i < store_config_int.size(); i++)
        config_count[store_config_int[i]]++;
      compute_entropy(config_count, sys.s_bias, nstates, p.pos_scale, p.neg_scale);
    }
    *seconds =
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::memcpy(pos_final, sys.pos.data(), sizeof(double) * 3 * p.nbeads);
    *config_final = update_config.config;
  } catch (const std::exception &e) {
    rc = -1;
  }
  std::cout.rdbuf(old);
  return rc;
}

// CPU-baseline leg: time `n_steps` MD steps (initialize_system + run_step via the
// traced loop, velocities from the sim's engine) and return valid collisions.
double ref_time_md(ref_sim *s, unsigned first_step, unsigned n_steps, uint64_t *collisions,
                   uint64_t *cells_crossed) {
  const Param &p = s->p;
  const uint64_t c0 = s->n_coll, x0 = s->n_cell;
  std::vector<double> normals(3 * p.nbeads);
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned k = 0; k < n_steps; k++) {
    std::normal_distribution<> g{0.0, std::sqrt(p.temp / p.m)};
    for (auto &v : normals) v = g(s->mt);
    ref_md_step(s, HMC_PHASE_EQ, first_step + k, normals.data());
  }
  const double dt =
      std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  *collisions = s->n_coll - c0;
  *cells_crossed = s->n_cell - x0;
  return dt;
}

} // extern "C"
