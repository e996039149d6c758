// hmc_program.cu — host-side phase loop of the reference's main (src/main.cc:362-491)
// over the C-ABI engine: equilibrate -> Wang-Landau -> G-test loop.
//
//   hmc_run_program_exact     ONE replica driven by std::mt19937(std::seed_seq(seeds)),
//                             the reference's RNG contract (main.cc:379-380): the host
//                             draws init_vel's normals and the crankshaft word tape with
//                             libstdc++ exactly as the reference does and injects them,
//                             so the GPU trajectory is the reference's, bit for bit.
//   hmc_run_program_ensemble  T transitions x R replicas with device Philox streams and
//                             pooled counts (the production path), optionally spread
//                             over the ranks of an hmc_comm (one NCCL all-reduce of the
//                             counts per G-test round), or with every replica running
//                             its own program (independent mode).
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
    uint32_t round_err = 0;
    RUN(hmc_get_counts(h, cc.data(), int(ns), nullptr, nullptr, &round_err));
    // the reference throws at the first violated check (main.cc:48-56,219-225): stop at the
    // end of the round that raised a flag instead of running on to max_g_test_count
    if (round_err & ~uint32_t(HMC_ERR_TAPE_UNDERRUN)) {
      gcount++;
      break;
    }
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

// ------------------------------------------------------------ ensemble program --

namespace {

struct EnsembleFail {
  hmc_handle h;
  hmc_ensemble_totals *tot;
  int operator()(int rc, const char *what) const {
    if (tot) std::snprintf(tot->message, sizeof tot->message, "%s: %s", what,
                           hmc_last_error(h) && *hmc_last_error(h) ? hmc_last_error(h)
                                                                   : hmc_last_error(nullptr));
    if (h) hmc_destroy(h);
    return rc;
  }
};

#define ERUN(call)                                                                       \
  do {                                                                                   \
    const int rc_ = (call);                                                              \
    if (rc_ != 0) return efail(rc_, #call);                                              \
  } while (0)

// flags that end a program (the reference throws at these points); a full sample buffer
// only drops samples
const uint32_t FATAL_FLAGS = ~uint32_t(HMC_ERR_SAMPLE_OVERFLOW);

} // namespace

extern "C" int hmc_run_program_ensemble(const hmc_params *tr, int T, int device,
                                        const double *start_pos,
                                        const hmc_ensemble_options *opt,
                                        const hmc_ensemble_sink *sink, hmc_ensemble_result *res,
                                        hmc_ensemble_totals *tot) {
  if (tot) std::memset(tot, 0, sizeof *tot);
  if (!tr || T < 1 || !start_pos || !opt || !res || opt->replicas < 1) {
    if (tot) std::snprintf(tot->message, sizeof tot->message, "hmc_run_program_ensemble: bad arguments");
    return -1;
  }
  const hmc_ensemble_sink no_sink{};
  if (!sink) sink = &no_sink;
  const hmc_params &p = tr[0];
  const unsigned nb = p.nbeads, ns = 1u << p.n_transient, R = opt->replicas;
  if (ns > 16) {
    if (tot) std::snprintf(tot->message, sizeof tot->message, "more than 16 states");
    return -1;
  }
  hmc_comm comm = opt->independent ? nullptr : opt->comm;
  const unsigned world = unsigned(hmc_comm_world(comm)), rank = unsigned(hmc_comm_rank(comm));
  const size_t Rl = size_t(T) * R; // replicas of this rank
  for (int t = 0; t < T; t++) std::memset(&res[t], 0, sizeof res[t]);

  hmc_handle h = nullptr;
  {
    const int rc = hmc_create(tr, T, int(R), device, &h);
    if (rc != 0) {
      if (tot) std::snprintf(tot->message, sizeof tot->message, "%s", hmc_last_error(nullptr));
      return rc;
    }
  }
  const EnsembleFail efail{h, tot};
  ERUN(hmc_set_replica_ids(h, rank * R, R * world));
  if (opt->list_capacity) ERUN(hmc_set_list_capacity(h, opt->list_capacity));
  {
    std::vector<double> pos(Rl * nb * 3);
    for (int t = 0; t < T; t++)
      for (unsigned j = 0; j < R; j++)
        std::memcpy(&pos[(size_t(t) * R + j) * nb * 3], start_pos + size_t(t) * nb * 3,
                    sizeof(double) * nb * 3);
    ERUN(hmc_set_positions(h, pos.data()));
  }
  ERUN(hmc_seed(h, opt->seed0, opt->seed1));
  ERUN(hmc_init_config_from_positions(h));
  std::vector<double> sb(size_t(T) * ns); // pooled mode: one row per transition
  for (int t = 0; t < T; t++)
    for (unsigned i = 0; i < ns; i++) // init_s, entropy.cc:14-20
      sb[size_t(t) * ns + i] = opt->start_s_bias
                                   ? opt->start_s_bias[size_t(t) * 16 + i]
                                   : 3 * (p.n_transient - unsigned(__builtin_popcount(i)));
  ERUN(hmc_set_s_bias(h, sb.data(), int(ns)));
  if (opt->hist_bins) ERUN(hmc_enable_dist_hist(h, opt->hist_bins, opt->hist_rmax));
  if (opt->independent) ERUN(hmc_enable_replica_counts(h, 1));

  const unsigned iters = opt->iters_per_round
                             ? opt->iters_per_round
                             : std::max(1u, (p.total_iter + R * world - 1) / (R * world));
  const unsigned eq = opt->eq_iters ? opt->eq_iters : std::max(1u, std::min(p.total_iter_eq, 8u));
  const unsigned cap = opt->max_rounds ? opt->max_rounds : p.max_g_test_count;
  const unsigned rows_round = iters * p.nsteps; // dist rows a replica produces per round
  const unsigned dist_cap =
      (opt->dist_rows && sink->on_dist) ? std::min(rows_round, (opt->dist_rows + R - 1) / R) : 0;
  const unsigned cfg_round = iters * p.nsteps / p.write_step + 2 + iters * p.mc_write;
  const unsigned cfg_cap = (opt->config_int_rows && sink->on_config_int)
                               ? std::min(cfg_round, (opt->config_int_rows + R - 1) / R)
                               : 0;
  if (dist_cap || cfg_cap) ERUN(hmc_enable_samples(h, dist_cap, cfg_cap));
  unsigned nl_max = 1;
  for (int t = 0; t < T; t++) nl_max = std::max(nl_max, tr[t].n_nonlocal);

  const auto t0 = std::chrono::steady_clock::now();
  double dev_ms = 0.0;
  auto add_ms = [&]() {
    uint64_t n = 0;
    float ms = 0.f;
    if (hmc_get_launch_stats(h, &n, &ms) == 0) dev_ms += ms;
  };

  // equilibration, main.cc:411-431
  ERUN(hmc_reset_clocks(h));
  ERUN(hmc_run(h, HMC_PHASE_EQ, 0, eq));
  ERUN(hmc_sync(h));
  add_ms();
  ERUN(hmc_init_config_from_positions(h));
  ERUN(hmc_reset_clocks(h));

  // Wang-Landau, main.cc:433-441: every replica its own estimate
  std::vector<double> sb_rep(Rl * ns);
  if (opt->wang_landau) {
    const unsigned nwl = wl_iters(p);
    // pooled mode: replicas whose global id is a multiple of `stride` take part
    const unsigned total = R * world;
    const unsigned stride = (!opt->independent && opt->wl_replicas && opt->wl_replicas < total)
                                ? total / opt->wl_replicas : 1u;
    std::vector<uint8_t> wait(Rl, 0);
    if (stride > 1) {
      for (size_t r = 0; r < Rl; r++) wait[r] = ((rank * R + unsigned(r % R)) % stride) != 0;
      ERUN(hmc_set_replica_skip(h, wait.data()));
    }
    if (nwl) {
      ERUN(hmc_run(h, HMC_PHASE_WL, 0, nwl));
      ERUN(hmc_sync(h));
      add_ms();
    }
    if (stride > 1) ERUN(hmc_set_replica_skip(h, nullptr));
    ERUN(hmc_get_s_bias(h, sb_rep.data(), int(ns)));
    if (!opt->independent) {
      // pooled start value: the mean over the participating replicas of the transition,
      // formed from exact fixed-point sums (2^-32 resolution) so that it does not depend on
      // how the replicas are spread over ranks
      std::vector<uint64_t> fx(size_t(T) * ns + size_t(T), 0);
      for (size_t r = 0; r < Rl; r++) {
        if (wait[r]) continue;
        for (unsigned i = 0; i < ns; i++)
          fx[(r / R) * ns + i] += uint64_t(int64_t(std::llround(sb_rep[r * ns + i] * 4294967296.0)));
        fx[size_t(T) * ns + r / R] += 1; // participants of the transition
      }
      ERUN(hmc_comm_allreduce_host_u64(comm, fx.data(), fx.size(), 0));
      for (int t = 0; t < T; t++)
        for (unsigned i = 0; i < ns; i++)
          sb[size_t(t) * ns + i] = double(int64_t(fx[size_t(t) * ns + i])) /
                                   (4294967296.0 * double(std::max<uint64_t>(1, fx[size_t(T) * ns + t])));
      ERUN(hmc_set_s_bias(h, sb.data(), int(ns)));
    }
  } else if (opt->independent) {
    ERUN(hmc_get_s_bias(h, sb_rep.data(), int(ns)));
  }

  // the G-test rounds, main.cc:451-491
  std::vector<uint8_t> done_t(T, 0), failed_t(T, 0), skip(Rl, 0), done_r(opt->independent ? Rl : 0, 0);
  std::vector<uint32_t> rounds_r(opt->independent ? Rl : 0, 0);
  std::vector<uint64_t> cc(size_t(T) * ns), cc_rep(opt->independent ? Rl * ns : 0);
  std::vector<uint32_t> err(Rl);
  std::vector<uint64_t> err_t(T);
  std::vector<double> dist_buf;
  std::vector<uint32_t> dist_n;
  std::vector<uint64_t> cfg_buf;
  std::vector<uint32_t> cfg_n;
  int rc_final = 0;
  for (unsigned rnd = 0; rnd < cap; rnd++) {
    ERUN(hmc_reset_clocks(h));
    ERUN(hmc_reset_counts(h));
    ERUN(hmc_set_replica_skip(h, skip.data()));
    ERUN(hmc_run(h, HMC_PHASE_MAIN, 0, iters));
    ERUN(hmc_allreduce_counts(h, comm)); // the round's one exchange (on the engine's stream)
    ERUN(hmc_get_counts(h, cc.data(), int(ns), nullptr, nullptr, err.data()));
    add_ms();
    // replica error flags end the program, as the reference's exceptions do
    for (int t = 0; t < T; t++) {
      uint32_t e = 0;
      for (unsigned j = 0; j < R; j++) e |= err[size_t(t) * R + j];
      err_t[t] = e;
    }
    if (comm && world > 1) {
      // OR over ranks, bit by bit (the sum of 0/1 flags per bit is exact)
      std::vector<uint64_t> bits(size_t(T) * 32);
      for (int t = 0; t < T; t++)
        for (int b = 0; b < 32; b++) bits[size_t(t) * 32 + b] = (err_t[t] >> b) & 1u;
      ERUN(hmc_comm_allreduce_host_u64(comm, bits.data(), bits.size(), 1));
      for (int t = 0; t < T; t++) {
        uint64_t e = 0;
        for (int b = 0; b < 32; b++) e |= (bits[size_t(t) * 32 + b] ? 1ull : 0ull) << b;
        err_t[t] = e;
      }
    }
    // (per transition, like the reference's one process per transition: the others go on)
    for (int t = 0; t < T; t++) {
      if (done_t[t]) continue;
      res[t].err_flags |= uint32_t(err_t[t]);
      if (uint32_t(err_t[t]) & FATAL_FLAGS) {
        failed_t[t] = 1;
        rc_final = -8;
      }
    }
    if (dist_cap) {
      dist_buf.resize(Rl * dist_cap * nl_max);
      dist_n.resize(Rl);
      ERUN(hmc_get_dist(h, dist_buf.data(), dist_n.data()));
      std::vector<double> rows;
      for (int t = 0; t < T; t++) {
        if (done_t[t]) continue;
        rows.clear();
        const unsigned nbonds = tr[t].n_nonlocal;
        for (unsigned j = 0; j < R; j++) {
          const size_t r = size_t(t) * R + j;
          const unsigned n = std::min(dist_n[r], dist_cap);
          for (unsigned q = 0; q < n; q++) {
            const double *src = &dist_buf[(r * dist_cap + q) * nl_max];
            rows.insert(rows.end(), src, src + nbonds);
          }
        }
        if (nbonds) sink->on_dist(sink->user, t, rows.data(), unsigned(rows.size() / nbonds), nbonds);
      }
    }
    if (cfg_cap) {
      cfg_buf.resize(Rl * cfg_cap);
      cfg_n.resize(Rl);
      ERUN(hmc_get_config_int(h, cfg_buf.data(), cfg_n.data()));
      std::vector<uint64_t> vals;
      for (int t = 0; t < T; t++) {
        if (done_t[t]) continue;
        vals.clear();
        for (unsigned j = 0; j < R; j++) {
          const size_t r = size_t(t) * R + j;
          const unsigned n = std::min(cfg_n[r], cfg_cap);
          vals.insert(vals.end(), &cfg_buf[r * cfg_cap], &cfg_buf[r * cfg_cap] + n);
        }
        sink->on_config_int(sink->user, t, vals.data(), unsigned(vals.size()));
      }
    }
    for (int t = 0; t < T; t++)
      if (failed_t[t] && !done_t[t]) {
        done_t[t] = 1; // stops here; res[t].accepted stays 0 and err_flags says why
        for (unsigned j = 0; j < R; j++) skip[size_t(t) * R + j] = 1;
        if (opt->independent)
          for (unsigned j = 0; j < R; j++) done_r[size_t(t) * R + j] = 1;
        if (tot) std::snprintf(tot->message, sizeof tot->message,
                               "replica error flags raised (first in transition %d, round %u; see err_flags)",
                               t, rnd);
      }
    bool all_done = true;
    if (!opt->independent) {
      for (int t = 0; t < T; t++) {
        if (done_t[t]) continue;
        uint64_t *c = &cc[size_t(t) * ns];
        double *s = &sb[size_t(t) * ns];
        hmc_compute_entropy(c, s, int(ns), p.pos_scale, p.neg_scale);
        double gv = 0, gc = 0, pv = 0;
        const int ok = hmc_g_test(c, int(ns), p.sig_level, &gv, &gc, &pv);
        res[t].rounds++;
        res[t].accepted = ok;
        for (unsigned i = 0; i < ns; i++) {
          res[t].s_bias[i] = s[i];
          res[t].config_count[i] = c[i];
        }
        if (sink->on_round) sink->on_round(sink->user, t, rnd, c, s, int(ns), ok, gv, gc, pv);
        if (ok) {
          done_t[t] = 1;
          for (unsigned j = 0; j < R; j++) skip[size_t(t) * R + j] = 1;
        } else {
          all_done = false;
        }
      }
      ERUN(hmc_set_s_bias(h, sb.data(), int(ns)));
    } else {
      ERUN(hmc_get_replica_counts(h, cc_rep.data()));
      for (size_t r = 0; r < Rl; r++) {
        if (done_r[r]) continue;
        uint64_t *c = &cc_rep[r * ns];
        double *s = &sb_rep[r * ns];
        hmc_compute_entropy(c, s, int(ns), p.pos_scale, p.neg_scale);
        rounds_r[r]++;
        if (hmc_g_test(c, int(ns), p.sig_level, nullptr, nullptr, nullptr)) {
          done_r[r] = 1;
          skip[r] = 1;
        } else {
          all_done = false;
        }
      }
      ERUN(hmc_set_s_bias_per_replica(h, sb_rep.data(), int(ns)));
      for (int t = 0; t < T; t++) {
        res[t].rounds = rnd + 1;
        bool every = true;
        for (unsigned j = 0; j < R; j++) every = every && done_r[size_t(t) * R + j];
        res[t].accepted = every ? 1 : 0;
        for (unsigned i = 0; i < ns; i++) res[t].config_count[i] = cc[size_t(t) * ns + i];
      }
    }
    if (all_done) break;
  }

  // totals, histograms, hand-off structures
  ERUN(hmc_allreduce_totals(h, comm));
  {
    std::vector<uint64_t> formed(T), broken(T);
    ERUN(hmc_get_counts(h, nullptr, int(ns), formed.data(), broken.data(), nullptr));
    for (int t = 0; t < T; t++) {
      res[t].formed = formed[t];
      res[t].broken = broken[t];
    }
    uint64_t ev[4];
    ERUN(hmc_get_event_counts(h, ev));
    if (tot) {
      tot->collisions = ev[0];
      tot->cell_crossings = ev[1];
    }
  }
  if (opt->hist_bins) {
    ERUN(hmc_allreduce_hist(h, comm));
    if (sink->on_hist) {
      std::vector<uint64_t> hist(size_t(T) * nl_max * 2 * opt->hist_bins);
      ERUN(hmc_get_dist_hist(h, hist.data()));
      for (int t = 0; t < T; t++)
        sink->on_hist(sink->user, t, &hist[size_t(t) * nl_max * 2 * opt->hist_bins], nl_max,
                      opt->hist_bins);
    }
  }
  {
    // unique_config (main.cc:211-216): the bonded structure of the lowest replica id that
    // recorded one; only its owner contributes non-zero words to the (exact) sum
    std::vector<double> up(Rl * nb * 3), uv(Rl * nb * 3);
    std::vector<uint8_t> found(Rl);
    ERUN(hmc_get_unique_config(h, up.data(), uv.data(), found.data()));
    std::vector<uint64_t> key(T);
    std::vector<int> local(T, -1);
    for (int t = 0; t < T; t++) {
      key[t] = ~0ull;
      for (unsigned j = 0; j < R; j++)
        if (found[size_t(t) * R + j]) {
          local[t] = int(j);
          key[t] = uint64_t(rank) * R + j;
          break;
        }
    }
    std::vector<uint64_t> best = key;
    ERUN(hmc_comm_allreduce_host_u64(comm, best.data(), best.size(), 2));
    const size_t words = size_t(nb) * 3;
    std::vector<uint64_t> pack(size_t(T) * words * 2, 0);
    for (int t = 0; t < T; t++)
      if (local[t] >= 0 && key[t] == best[t]) {
        const size_t r = size_t(t) * R + size_t(local[t]);
        std::memcpy(&pack[size_t(t) * words * 2], &up[r * words], sizeof(double) * words);
        std::memcpy(&pack[size_t(t) * words * 2 + words], &uv[r * words], sizeof(double) * words);
      }
    ERUN(hmc_comm_allreduce_host_u64(comm, pack.data(), pack.size(), 0));
    for (int t = 0; t < T; t++) {
      res[t].have_unique = best[t] != ~0ull;
      if (res[t].have_unique) {
        std::memcpy(res[t].unique_pos, &pack[size_t(t) * words * 2], sizeof(double) * words);
        std::memcpy(res[t].unique_vel, &pack[size_t(t) * words * 2 + words], sizeof(double) * words);
      }
    }
  }
  if (opt->independent) {
    for (int t = 0; t < T; t++) {
      if (sink->on_replicas)
        sink->on_replicas(sink->user, t, &sb_rep[size_t(t) * R * ns], &rounds_r[size_t(t) * R],
                          &done_r[size_t(t) * R], R, int(ns));
      // the transition's value: mean over its replicas (tools/avg_s_bias.py averages seeds)
      for (unsigned i = 0; i < ns; i++) {
        double m = 0.0;
        for (unsigned j = 0; j < R; j++) m += sb_rep[(size_t(t) * R + j) * ns + i];
        res[t].s_bias[i] = m / double(R);
      }
    }
  }
  if (tot) {
    tot->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    tot->seconds_device = dev_ms * 1e-3;
    tot->collectives = hmc_comm_collectives(comm);
  }
  hmc_destroy(h);
  return rc_final;
}
