// hmc_program.cu — host-side phase loop of the reference's main (src/main.cc:362-491)
// over the C-ABI engine: equilibrate -> Wang-Landau -> G-test loop.
//
//   hmc_run_program_exact     ONE replica driven by std::mt19937(std::seed_seq(seeds)),
//                             the reference's RNG contract (main.cc:379-380): the host
//                             draws init_vel's normals and the crankshaft word tape with
//                             libstdc++ exactly as the reference does and injects them,
//                             so the GPU trajectory is the reference's, bit for bit.
//   hmc_run_program_ensemble  R replicas with device Philox streams and pooled counts
//                             (the production path).
// Only orchestration and O(nstates) arithmetic happen here; every trajectory
// operation runs in the CUDA kernel.
#include "hmc_internal.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <random>
#include <sstream>
#include <string>
#include <vector>


namespace {

struct Tape {
  std::vector<uint32_t> w;
  std::vector<double> sn, cs;
};

// the next n words of `mt` (copy) with sin/cos of the angle libstdc++'s
// uniform_real_distribution(0, 2*pi) forms from words c, c+1 (crankshaft.cc:235,247)
void fill_tape(const std::mt19937 &mt, unsigned n, Tape &t) {
  t.w.resize(n);
  t.sn.resize(n);
  t.cs.resize(n);
  std::mt19937 c = mt;
  for (unsigned k = 0; k < n; k++) t.w[k] = uint32_t(c());
  std::mt19937 base = mt;
  for (unsigned k = 0; k < n; k++) {
    if (k + 1 < n) {
      std::mt19937 at = base;
      std::uniform_real_distribution<> theta(0.0, 2 * M_PI);
      const double th = theta(at);
      t.sn[k] = std::sin(th);
      t.cs[k] = std::cos(th);
    } else {
      t.sn[k] = 0.0;
      t.cs[k] = 1.0;
    }
    base();
  }
}

// init_vel's draws for one step (hardspheres.cc:197-209): fresh distribution object
void draw_normals(std::mt19937 &mt, double sigma, double *out, unsigned n) {
  std::normal_distribution<> g{0.0, sigma};
  for (unsigned k = 0; k < n; k++) out[k] = g(mt);
}

unsigned wl_iters(const hmc_params &p) { // main.cc:283-304
  double gamma = p.gamma;
  unsigned it = 0;
  while (gamma > p.gamma_f) {
    it += 1;
    gamma = 1.0 / double(it);
  }
  return it;
}

#define RUN(call)                                                                        \
  do {                                                                                   \
    const int rc_ = (call);                                                              \
    if (rc_ != 0) {                                                                      \
      if (res) std::snprintf(res->message, sizeof res->message, "%s: %s", #call,          \
                             hmc_last_error(h));                                         \
      if (h) hmc_destroy(h);                                                             \
      return rc_;                                                                        \
    }                                                                                    \
  } while (0)

} // namespace

namespace {

struct DistCollector { // sink used by the plain entry point
  double *out;
  unsigned cap, n, nbonds;
};
void collect_dist(void *user, const double *rows, unsigned nrows, unsigned nbonds) {
  DistCollector *c = static_cast<DistCollector *>(user);
  const unsigned take = std::min(nrows, c->cap - c->n);
  std::memcpy(c->out + size_t(c->n) * nbonds, rows, sizeof(double) * size_t(take) * nbonds);
  c->n += take;
}

} // namespace

extern "C" {

int hmc_run_program_exact(const hmc_params *p, int device, const double *pos_in,
                          unsigned max_gtest_iters, unsigned dist_rows_cap, double *dist_out,
                          unsigned *n_dist_rows, hmc_program_result *res) {
  DistCollector c{dist_out, dist_rows_cap, 0, p ? p->n_nonlocal : 0};
  hmc_sink sink{};
  sink.user = &c;
  if (dist_rows_cap && dist_out) sink.on_dist = collect_dist;
  const int rc = hmc_run_program_exact_sink(p, device, pos_in, nullptr, max_gtest_iters, &sink, res);
  if (n_dist_rows) *n_dist_rows = c.n;
  return rc;
}

int hmc_run_program_exact_sink(const hmc_params *p, int device, const double *pos_in,
                               const hmc_snapshot_in *snap, unsigned max_gtest_iters,
                               const hmc_sink *sink, hmc_program_result *res) {
  if (!p || !res) return -1;
  const hmc_sink no_sink{};
  if (!sink) sink = &no_sink;
  std::memset(res, 0, sizeof *res);
  const unsigned nb = p->nbeads, ns = 1u << p->n_transient;
  if (ns > 16) {
    std::snprintf(res->message, sizeof res->message, "more than 16 states");
    return -1;
  }
  hmc_handle h = nullptr;
  {
    const int rc = hmc_create(p, 1, 1, device, &h);
    if (rc != 0) {
      std::snprintf(res->message, sizeof res->message, "%s", hmc_last_error(nullptr));
      return rc;
    }
  }
  std::seed_seq seq(p->seeds, p->seeds + p->n_seeds); // main.cc:379-380
  std::mt19937 mt(seq);
  const double vsig = std::sqrt(p->temp / p->m);
  std::vector<double> pos(3 * nb);
  // initialize_pos, main.cc:26-43
  std::vector<double> sb(ns);
  for (unsigned i = 0; i < ns; i++) sb[i] = 3 * (p->n_transient - unsigned(__builtin_popcount(i)));
  if (snap) { // read_snapshot, snapshot.cc:70-120
    std::memcpy(pos.data(), snap->pos, sizeof(double) * 3 * nb);
    std::memcpy(sb.data(), snap->s_bias, sizeof(double) * ns);
    std::stringstream ss(snap->mt_state);
    ss >> mt;
    if (ss.fail()) {
      std::snprintf(res->message, sizeof res->message, "snapshot: bad mt19937 state");
      hmc_destroy(h);
      return -6;
    }
  } else if (pos_in) std::memcpy(pos.data(), pos_in, sizeof(double) * 3 * nb);
  else if (hmc_init_pos_engine(p, &mt, pos.data()) != 0) {
    std::snprintf(res->message, sizeof res->message,
                  "number of tries exceeded while placing bead in box");
    hmc_destroy(h);
    return -4;
  }
  RUN(hmc_set_positions(h, pos.data()));
  if (snap) {
    const uint64_t c = snap->config;
    RUN(hmc_set_config(h, &c));
  } else if (pos_in) RUN(hmc_init_config_from_positions(h));
  RUN(hmc_set_s_bias(h, sb.data(), int(ns))); // init_s, entropy.cc:14-20
  const auto t0 = std::chrono::steady_clock::now();

  // equilibration, main.cc:411-431: every step's draws are known up front (no MC moves)
  RUN(hmc_reset_clocks(h));
  {
    const unsigned nsteps = p->total_iter_eq * p->nsteps_eq;
    const unsigned chunk = 4096;
    std::vector<double> normals;
    for (unsigned s0 = 0; s0 < nsteps; s0 += chunk) {
      const unsigned n = std::min(chunk, nsteps - s0);
      normals.resize(size_t(n) * 3 * nb);
      for (unsigned k = 0; k < n; k++) draw_normals(mt, vsig, &normals[size_t(k) * 3 * nb], 3 * nb);
      RUN(hmc_md_steps_injected(h, HMC_PHASE_EQ, s0, n, normals.data()));
    }
  }
  // Wang-Landau, main.cc:433-441
  RUN(hmc_init_config_from_positions(h));
  RUN(hmc_reset_clocks(h));
  {
    const unsigned ntraj = wl_iters(*p) * ns;
    std::vector<double> normals(size_t(ntraj) * 3 * nb);
    for (unsigned k = 0; k < ntraj; k++) draw_normals(mt, vsig, &normals[size_t(k) * 3 * nb], 3 * nb);
    if (ntraj) RUN(hmc_md_steps_injected(h, HMC_PHASE_WL, 0, ntraj, normals.data()));
  }
  RUN(hmc_get_s_bias(h, sb.data(), int(ns)));
  // per-iteration sample buffers: one dist row per MD step; config/int gets one entry
  // per recorded MD step plus mc_write-1 per block of crankshaft moves
  const unsigned dist_cap = sink->on_dist ? p->total_iter * p->nsteps : 0;
  const unsigned cfg_cap =
      sink->on_config_int
          ? p->total_iter * p->nsteps / p->write_step + 2 + p->total_iter * p->mc_write
          : 0;
  if (dist_cap || cfg_cap) RUN(hmc_enable_samples(h, dist_cap, cfg_cap));
  bool unique_sent = false;

  // the G-test loop, main.cc:451-491
  const unsigned cap = max_gtest_iters ? max_gtest_iters : p->max_g_test_count;
  unsigned gcount = 0;
  int accepted = 0;
  std::vector<uint64_t> cc(ns);
  std::vector<double> normals(size_t(p->nsteps) * 3 * nb);
  Tape tape;
  const unsigned tape_n = 2048;
  do {
    RUN(hmc_reset_clocks(h));
    RUN(hmc_reset_counts(h));
    for (unsigned iter = 0; iter < p->total_iter; iter++) {
      for (unsigned k = 0; k < p->nsteps; k++)
        draw_normals(mt, vsig, &normals[size_t(k) * 3 * nb], 3 * nb);
      RUN(hmc_md_steps_injected(h, HMC_PHASE_MAIN, iter * p->nsteps, p->nsteps, normals.data()));
      unsigned done = 0;
      while (done < p->mc_moves) {
        fill_tape(mt, tape_n, tape);
        uint32_t used = 0;
        // one launch per attempt to finish the remaining moves; on a tape underrun the
        // device rolls the unfinished move back and reports how far it got
        RUN(hmc_crankshaft_injected(h, done, p->mc_moves - done, tape.w.data(), tape.sn.data(),
                                    tape.cs.data(), tape_n, &used));
        uint32_t moves_done = 0;
        RUN(hmc_get_moves_done(h, &moves_done));
        if (moves_done == 0) {
          std::snprintf(res->message, sizeof res->message, "crankshaft move needs > %u words", tape_n);
          hmc_destroy(h);
          return -5;
        }
        mt.discard(used);
        done += moves_done;
      }
    }
    RUN(hmc_get_counts(h, cc.data(), int(ns), nullptr, nullptr, nullptr));
    if (dist_cap) { // DistWriter appends one row per MD step (main.cc:199-200)
      std::vector<double> rows(size_t(dist_cap) * std::max(1u, p->n_nonlocal));
      uint32_t n = 0;
      RUN(hmc_get_dist(h, rows.data(), &n));
      sink->on_dist(sink->user, rows.data(), std::min<uint32_t>(n, dist_cap), p->n_nonlocal);
    }
    if (cfg_cap) {
      std::vector<uint64_t> ci(cfg_cap);
      uint32_t n = 0;
      RUN(hmc_get_config_int(h, ci.data(), &n));
      sink->on_config_int(sink->user, ci.data(), std::min<uint32_t>(n, cfg_cap));
    }
    if (sink->on_unique && !unique_sent) {
      std::vector<double> up(3 * nb), uv(3 * nb);
      uint8_t found = 0;
      RUN(hmc_get_unique_config(h, up.data(), uv.data(), &found));
      if (found) {
        sink->on_unique(sink->user, up.data(), uv.data(), 1, nb);
        unique_sent = true;
      }
    }
    hmc_compute_entropy(cc.data(), sb.data(), int(ns), p->pos_scale, p->neg_scale);
    RUN(hmc_set_s_bias(h, sb.data(), int(ns)));
    gcount++;
    double gv = 0, gc = 0, pv = 0;
    accepted = hmc_g_test(cc.data(), int(ns), p->sig_level, &gv, &gc, &pv);
    if (sink->on_round)
      sink->on_round(sink->user, gcount - 1, cc.data(), sb.data(), int(ns), accepted, gv, gc, pv);
    if (sink->on_snapshot) {
      std::vector<double> cur(3 * nb);
      uint64_t cfg = 0;
      RUN(hmc_get_positions(h, cur.data()));
      RUN(hmc_get_config(h, &cfg));
      std::stringstream ss;
      ss << mt;
      sink->on_snapshot(sink->user, cur.data(), sb.data(), int(ns), cfg, ss.str().c_str());
    }
  } while (gcount < cap && !accepted);

  res->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (unsigned i = 0; i < ns; i++) {
    res->s_bias[i] = sb[i];
    res->config_count[i] = cc[i];
  }
  res->gtest_iters = gcount;
  res->accepted = accepted;
  uint64_t ev[4];
  RUN(hmc_get_event_counts(h, ev));
  res->collisions = ev[0];
  res->cell_crossings = ev[1];
  uint32_t err = 0;
  RUN(hmc_get_counts(h, nullptr, int(ns), nullptr, nullptr, &err));
  res->err_flags = err & ~uint32_t(HMC_ERR_TAPE_UNDERRUN);
  RUN(hmc_get_positions(h, res->pos_final));
  uint64_t cfg = 0;
  RUN(hmc_get_config(h, &cfg));
  res->config_final = cfg;
  hmc_destroy(h);
  return 0;
}

} // extern "C"
