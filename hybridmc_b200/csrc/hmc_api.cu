// hmc_api.cu — the C-ABI of include/hybridmc_b200.h: handle, device memory, table
// construction, launches.  Host-side orchestration only; every trajectory operation
// runs in hmc_kernels.cu on the GPU.  There is no CPU fallback.
#include "hmc_internal.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define HMC_ABI_VERSION 2

namespace {
thread_local std::string g_create_error;
}
void hmc_set_global_error(const char *msg) { g_create_error = msg; }

struct hmc_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t tm0 = nullptr, tm1 = nullptr; // hmc_timer_start / hmc_timer_stop
  bool timing_pending = false;
  float last_ms = 0.f;
  uint64_t n_launches = 0;
  int last_grid = 0, last_block = 0;
  std::string err;

  std::vector<hmc_params> params;
  int T = 0, Rper = 0, R = 0, nb = 0, nstates = 0, npairs = 0, nslots = 0;
  int group = 32;
  unsigned n_nonlocal_max = 0;

  // device buffers
  TransDev *d_trans = nullptr;
  uint8_t *d_pairinfo = nullptr;
  unsigned long long *d_bondmask = nullptr;
  uchar2 *d_pair_ij = nullptr;
  double *d_pos = nullptr, *d_vel = nullptr, *d_sbias = nullptr;
  uint64_t *d_config = nullptr;
  uint32_t *d_err = nullptr;
  unsigned long long *d_acc = nullptr;
  size_t n_acc = 0;
  unsigned long long *d_acc_rep = nullptr; // per-replica counts (independent-program mode)
  uint8_t *d_skip = nullptr;
  bool skip_on = false;
  uint32_t rng_first = 0, rng_total = 0;
  int list_cap_override = 0; // hmc_set_list_capacity
  double *d_dist = nullptr;
  uint32_t *d_dist_n = nullptr;
  uint32_t dist_cap = 0;
  unsigned long long *d_cfg_int = nullptr;
  uint32_t *d_cfg_n = nullptr;
  uint32_t cfg_cap = 0;
  unsigned long long *d_hist = nullptr;
  uint32_t hist_bins = 0;
  double hist_rmax = 0.0;
  double *d_uniq_pos = nullptr, *d_uniq_vel = nullptr;
  uint8_t *d_uniq_found = nullptr;
  hmc_event *d_trace = nullptr;
  uint32_t *d_trace_n = nullptr;
  int trace_replica = -1;
  uint32_t trace_cap = 0;
  unsigned int *d_work = nullptr;
  // staging for injected streams
  double *d_normals = nullptr;
  size_t normals_cap = 0;
  uint32_t *d_tape_words = nullptr, *d_tape_consumed = nullptr;
  double *d_tape_sin = nullptr, *d_tape_cos = nullptr;
  size_t tape_cap = 0;

  // schedule state
  double t_now = 0.0;
  uint32_t seed0 = 0, seed1 = 0;
  unsigned long long md_draws = 0, mc_draws = 0;
};

namespace {

int fail(hmc_handle h, const std::string &msg, int code = -1) {
  if (h) h->err = msg;
  else g_create_error = msg;
  return code;
}

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(h, std::string(#call) + ": " + cudaGetErrorString(e_), -2);            \
  } while (0)

// Zero-initialised device buffer.  cudaMemset runs on the legacy default stream and is
// asynchronous to the host, while the engine's stream is NON-BLOCKING (no implicit ordering
// with the default stream): without the synchronisation below the zeroing can overtake — and
// wipe — the first upload on the engine's stream (seen when several processes share a GPU:
// tests/test_gpu_program_ensemble.py::test_concurrent_executables_share_one_gpu).
template <class T> cudaError_t dalloc(T **p, size_t n) {
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), n * sizeof(T));
  if (e == cudaSuccess) e = cudaMemset(*p, 0, n * sizeof(T));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  return e;
}

void normalise(uint32_t &i, uint32_t &j) { // NonlocalBonds ctor, config.cc:6-13
  if (i > j) {
    const uint32_t t = i;
    i = j;
    j = t;
  }
}

// first list position of pair (i,j), or -1 (std::find, config.h:34)
int bond_index(const uint32_t (*b)[2], unsigned n, unsigned i, unsigned j) {
  for (unsigned k = 0; k < n; k++) {
    uint32_t a = b[k][0], c = b[k][1];
    normalise(a, c);
    if (a == i && c == j) return int(k);
  }
  return -1;
}

void fill_kargs(hmc_handle h, KArgs &a) {
  const hmc_params &p = h->params[0];
  std::memset(&a, 0, sizeof a);
  a.nb = h->nb;
  a.ncell = int(p.ncell);
  a.npairs = h->npairs;
  a.list_cap = hmc_list_capacity(h->nb, h->group, h->list_cap_override);
  a.nstates = h->nstates;
  a.l = p.length;
  a.inv_l = 1.0 / p.length;       // Box, box.h:15
  a.lcell = p.length / p.ncell;   // Cells{ncell, length/ncell}, main.cc:188
  a.inv_lcell = 1 / a.lcell;      // cells.h:29
  a.near_min2 = p.near_min * p.near_min;
  a.near_max2 = p.near_max * p.near_max;
  a.nnear_min2 = p.nnear_min * p.nnear_min;
  a.nnear_max2 = p.nnear_max * p.nnear_max;
  a.rh2 = p.rh * p.rh;
  a.s_in_tab[0] = a.near_min2;
  a.s_in_tab[1] = a.nnear_min2;
  a.s_in_tab[2] = a.rh2;
  a.s_out_tab[0] = a.near_max2;
  a.s_out_tab[1] = a.nnear_max2;
  a.s_out_tab[2] = a.rh2;
  a.s_mid_tab[0] = 0.5 * (a.near_min2 + a.near_max2);
  a.s_mid_tab[1] = 0.5 * (a.nnear_min2 + a.nnear_max2);
  a.sigma2_tab[0] = a.near_min2;
  a.sigma2_tab[1] = a.near_max2;
  a.sigma2_tab[2] = a.nnear_min2;
  a.sigma2_tab[3] = a.nnear_max2;
  a.sigma2_tab[4] = a.rh2;
  a.one = 1.0;
  a.cell_masks = p.ncell <= 10 ? 1 : 0;
  if (const char *cm = getenv("HMC_CELL_MASKS")) a.cell_masks = a.cell_masks && atoi(cm) != 0; // test knob
  for (unsigned c = 0; c < 16; c++) {
    const unsigned nc = p.ncell;
    a.nbr_tab[c] = (a.cell_masks && c < nc)
                       ? uint16_t((1u << c) | (1u << ((c + 1) % nc)) | (1u << ((c + nc - 1) % nc)))
                       : uint16_t(0);
  }
  a.m = p.m;
  a.vsigma = std::sqrt(p.temp / p.m);
  a.del_t = p.del_t;
  a.del_t_wl = p.del_t_wl;
  a.gamma0 = p.gamma;
  a.nsteps = p.nsteps;
  a.nsteps_eq = p.nsteps_eq;
  a.nsteps_wl = p.nsteps_wl;
  a.write_step = p.write_step;
  a.mc_moves = p.mc_moves;
  a.mc_write = p.mc_write;
  a.trans = h->d_trans;
  a.pairinfo = h->d_pairinfo;
  a.bondmask = h->d_bondmask;
  a.pair_ij = h->d_pair_ij;
  a.n_replicas = h->R;
  a.replicas_per_trans = h->Rper;
  a.pos = h->d_pos;
  a.vel = h->d_vel;
  a.config = reinterpret_cast<uint64_t *>(h->d_config);
  a.s_bias = h->d_sbias;
  a.err = h->d_err;
  a.acc = h->d_acc;
  a.acc_rep = h->d_acc_rep;
  a.skip = h->skip_on ? h->d_skip : nullptr;
  a.rng_first = h->rng_first;
  a.rng_total = h->rng_total ? h->rng_total : uint32_t(h->Rper);
  a.n_trans = h->T;
  a.dist = h->d_dist;
  a.dist_n = h->d_dist_n;
  a.dist_cap = h->dist_cap;
  a.dist_stride = h->n_nonlocal_max ? h->n_nonlocal_max : 1;
  a.cfg_int = h->d_cfg_int;
  a.cfg_n = h->d_cfg_n;
  a.cfg_cap = h->cfg_cap;
  a.hist = h->d_hist;
  a.hist_bins = h->hist_bins;
  a.hist_scale = h->hist_bins ? double(h->hist_bins) / h->hist_rmax : 0.0;
  a.uniq_pos = h->d_uniq_pos;
  a.uniq_vel = h->d_uniq_vel;
  a.uniq_found = h->d_uniq_found;
  a.trace = h->d_trace;
  a.trace_n = h->d_trace_n;
  a.trace_replica = h->trace_replica;
  a.trace_cap = h->trace_cap;
  a.t_now = h->t_now;
  a.seed0 = h->seed0;
  a.seed1 = h->seed1;
  a.md_draw_base = h->md_draws;
  a.mc_draw_base = h->mc_draws;
  a.work_counter = h->d_work;
  hmc_smem_layout(h->nb, h->group, h->list_cap_override, &a.off_tnn, &a.off_list, &a.off_plist,
                  &a.smem_per_replica);
  a.short_chain = (double(h->nb - 1) * p.near_max < 1.45 * p.length) ? 1 : 0;
  a.chain_max = double(h->nb - 1) * p.near_max;
  {
    const float fl = float(p.length), finv = float(1.0 / p.length), frh2 = float(p.rh * p.rh);
    a.pf_invl2 = make_float2(finv, finv);
    a.pf_nl2 = make_float2(-fl, -fl);
    a.pf_nrh2 = make_float2(-frh2, -frh2);
    a.pf_magic = make_float2(12582912.0f, 12582912.0f); // 1.5 * 2^23: rint by adding
    a.pf_nmagic = make_float2(-12582912.0f, -12582912.0f);
    a.pf_c1e5 = make_float2(1e-5f, 1e-5f);
    a.pf_c1e2 = make_float2(1e-2f, 1e-2f);
    a.pf_xc = float(a.chain_max) + fl;
    a.pf_l = fl;
  }
  {
    const char *tp = getenv("HMC_TWO_PASS"); // experiment switch: 0 / 1 overrides the heuristic
    // needs: one-image wrap (short chain) and partners in adjacent cells closer than half a
    // box in every axis (ncell >= 5), so that single precision picks the same image
    const bool can = a.short_chain && p.ncell >= 5;
    a.two_pass = (tp ? atoi(tp) != 0 : h->nb >= 16) && can;
  }
  a.max_events_per_step = 1u << 18; // ~30x the largest step of the shipped configs
}

int launch(hmc_handle h, const KArgs &a) {
  CK(cudaSetDevice(h->device));
  CK(cudaMemsetAsync(h->d_work, 0, sizeof(unsigned int), h->stream));
  CK(cudaEventRecord(h->ev0, h->stream));
  CK(hmc_launch(a, h->group, h->stream, &h->last_grid, &h->last_block));
  CK(cudaEventRecord(h->ev1, h->stream));
  h->timing_pending = true;
  h->n_launches++;
  return 0;
}

int sync(hmc_handle h) {
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->timing_pending) {
    CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    h->timing_pending = false;
  }
  return 0;
}

unsigned wl_outer_iterations(const hmc_params &p) { // main.cc:283-304
  double gamma = p.gamma;
  unsigned it = 0;
  while (gamma > p.gamma_f) {
    it += 1;
    gamma = 1.0 / double(it);
  }
  return it;
}

// regularised lower incomplete gamma P(a,x) (series / Lentz continued fraction):
// chi-squared cdf for the G-test.  The reference calls Boost.Math here
// (entropy.cc:98-104); Boost is not available, the textbook forms are.
double gamma_p(double a, double x) {
  if (x <= 0) return 0.0;
  const double gln = std::lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int k = 0; k < 100000; k++) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
    }
    return sum * std::exp(-x + a * std::log(x) - gln);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, hh = d;
  for (int i = 1; i < 100000; i++) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    hh *= del;
    if (std::fabs(del - 1.0) < 1e-16) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - gln) * hh;
}

double compute_norm(const uint64_t *cc, int n) { // entropy.cc:24-34
  double norm = 0.0;
  for (int i = 0; i < n; i++) norm += cc[i];
  norm /= double(n);
  return norm;
}

} // namespace

extern "C" {

int hmc_abi_version(void) { return HMC_ABI_VERSION; }

const char *hmc_last_error(hmc_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int hmc_create(const hmc_params *tr, int n_transitions, int replicas_per_transition, int device,
               hmc_handle *out) {
  hmc_handle h = nullptr; // errors before the handle exists go to the thread-local slot
  if (!tr || !out || n_transitions < 1 || replicas_per_transition < 1)
    return fail(h, "hmc_create: bad arguments");
  *out = nullptr;
  const hmc_params &p0 = tr[0];
  if (p0.nbeads < 4 || p0.nbeads > HMC_MAX_BEADS)
    return fail(h, "hmc_create: nbeads must be in [4, 64]");
  if (p0.n_transient > 10) return fail(h, "hmc_create: more than 10 transient bonds per run");
  for (int t = 0; t < n_transitions; t++) {
    const hmc_params &p = tr[t];
    if (p.nbeads != p0.nbeads || p.length != p0.length || p.ncell != p0.ncell ||
        p.n_transient != p0.n_transient || p.m != p0.m || p.temp != p0.temp ||
        p.near_min != p0.near_min || p.near_max != p0.near_max || p.nnear_min != p0.nnear_min ||
        p.nnear_max != p0.nnear_max || p.rh != p0.rh || p.del_t != p0.del_t ||
        p.del_t_wl != p0.del_t_wl || p.nsteps != p0.nsteps || p.nsteps_eq != p0.nsteps_eq ||
        p.nsteps_wl != p0.nsteps_wl || p.write_step != p0.write_step ||
        p.mc_moves != p0.mc_moves || p.mc_write != p0.mc_write || p.gamma != p0.gamma)
      return fail(h, "hmc_create: transitions of one ensemble must share the chain, box and schedule");
    if (p.n_nonlocal > HMC_MAX_BONDS || p.n_transient > HMC_MAX_BONDS ||
        p.n_permanent > HMC_MAX_BONDS)
      return fail(h, "hmc_create: too many bonds");
    if (p.has_stair && (p.stair < p.rc || (p.has_p_rc && p.stair < p.p_rc)))
      return fail(h, "stair boundary is smaller than rc"); // main.cc:535-537
    if (!(p.ncell >= 4)) return fail(h, "bead ncell must be at least 4"); // main.cc:58-60
    if (p.ncell > 255) return fail(h, "hmc_create: ncell > 255");
    const double check_rc = p.has_stair ? p.stair : p.rc;
    if (!(p.length / p.ncell >= check_rc))
      return fail(h, "bead lcell must be at least rc"); // main.cc:62-69
    if (p.write_step == 0) return fail(h, "hmc_create: write_step must be > 0");
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1)
    return fail(h, std::string("hmc_create: no CUDA device (") + cudaGetErrorString(e) +
                       "); this engine has no CPU fallback",
                -3);
  if (device < 0 || device >= ndev) return fail(h, "hmc_create: bad device index");

  hmc_handle eng = new hmc_engine();
  h = eng;
  h->device = device;
  h->params.assign(tr, tr + n_transitions);
  h->T = n_transitions;
  h->Rper = replicas_per_transition;
  h->R = h->T * h->Rper;
  h->nb = int(p0.nbeads);
  h->nstates = 1 << p0.n_transient;
  h->npairs = h->nb * (h->nb - 1) / 2;
  h->nslots = h->nb + h->npairs;
  h->group = hmc_pick_group(h->nb);
  if (const char *g = std::getenv("HMC_GROUP")) { // tuning knob: lanes per replica
    const int v = std::atoi(g);
    if (v == 8 || v == 16 || v == 32) h->group = v;
  }

#define CKC(call)                                                                        \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      g_create_error = std::string(#call) + ": " + cudaGetErrorString(e_);               \
      hmc_destroy(eng);                                                                  \
      return -2;                                                                         \
    }                                                                                    \
  } while (0)

  CKC(cudaSetDevice(device));
  CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CKC(cudaEventCreate(&h->ev0));
  CKC(cudaEventCreate(&h->ev1));
  CKC(cudaEventCreate(&h->tm0));
  CKC(cudaEventCreate(&h->tm1));

  // tables
  const int nb = h->nb, np = h->npairs;
  std::vector<uchar2> pair_ij(np);
  for (int i = 0, q = 0; i < nb; i++)
    for (int j = i + 1; j < nb; j++, q++) pair_ij[q] = make_uchar2((unsigned char)i, (unsigned char)j);
  const size_t pstride = size_t(nb) * 64; // symmetric [nb][64] matrix per transition
  std::vector<uint8_t> pairinfo(size_t(h->T) * pstride, 0);
  std::vector<unsigned long long> bondmask(size_t(h->T) * nb, 0ull);
  std::vector<TransDev> trans(h->T);
  for (int t = 0; t < h->T; t++) {
    const hmc_params &p = tr[t];
    TransDev &d = trans[t];
    std::memset(&d, 0, sizeof d);
    d.rc2 = p.rc * p.rc;
    d.stair2 = p.has_stair ? p.stair * p.stair : 0.0;
    d.p_rc2 = p.has_p_rc ? p.p_rc * p.p_rc : 0.0;
    d.has_stair = p.has_stair ? 1u : 0u;
    d.has_p_rc = p.has_p_rc ? 1u : 0u;
    d.max_nbonds = p.max_nbonds;
    d.n_transient = p.n_transient;
    for (unsigned k = 0; k < p.n_transient; k++) {
      uint32_t i = p.transient_bonds[k][0], j = p.transient_bonds[k][1];
      normalise(i, j);
      d.tb_i[k] = uint8_t(i);
      d.tb_j[k] = uint8_t(j);
    }
    unsigned nl = 0;
    for (int i = 0, q = 0; i < nb; i++)
      for (int j = i + 1; j < nb; j++, q++) {
        if (j - i < 3) continue;
        const int tbk = bond_index(p.transient_bonds, p.n_transient, i, j);
        const int pbk = bond_index(p.permanent_bonds, p.n_permanent, i, j);
        uint8_t info = 0;
        if (tbk >= 0) info |= uint8_t(tbk + 1);
        if (pbk >= 0) info |= 0x80u;
        pairinfo[size_t(t) * pstride + size_t(i) * 64 + j] = info;
        pairinfo[size_t(t) * pstride + size_t(j) * 64 + i] = info;
        if (info) {
          bondmask[size_t(t) * nb + i] |= 1ull << j;
          bondmask[size_t(t) * nb + j] |= 1ull << i;
        }
        if (bond_index(p.nonlocal_bonds, p.n_nonlocal, i, j) >= 0 && nl < HMC_MAX_BONDS) {
          d.nl_i[nl] = uint8_t(i);
          d.nl_j[nl] = uint8_t(j);
          nl++;
        }
      }
    d.n_nonlocal = nl;
    if (nl > h->n_nonlocal_max) h->n_nonlocal_max = nl;
  }
  CKC(dalloc(&h->d_trans, size_t(h->T)));
  CKC(cudaMemcpy(h->d_trans, trans.data(), sizeof(TransDev) * h->T, cudaMemcpyHostToDevice));
  CKC(dalloc(&h->d_pairinfo, pairinfo.size()));
  CKC(cudaMemcpy(h->d_pairinfo, pairinfo.data(), pairinfo.size(), cudaMemcpyHostToDevice));
  CKC(dalloc(&h->d_bondmask, bondmask.size()));
  CKC(cudaMemcpy(h->d_bondmask, bondmask.data(), sizeof(unsigned long long) * bondmask.size(),
                 cudaMemcpyHostToDevice));
  CKC(dalloc(&h->d_pair_ij, size_t(np)));
  CKC(cudaMemcpy(h->d_pair_ij, pair_ij.data(), sizeof(uchar2) * np, cudaMemcpyHostToDevice));

  const size_t R = size_t(h->R);
  CKC(dalloc(&h->d_pos, R * nb * 3));
  CKC(dalloc(&h->d_vel, R * nb * 3));
  CKC(dalloc(&h->d_sbias, R * h->nstates));
  CKC(dalloc(&h->d_config, R));
  CKC(dalloc(&h->d_err, R));
  h->n_acc = size_t(h->T) * h->nstates + 2 * size_t(h->T) + 4;
  CKC(dalloc(&h->d_acc, h->n_acc));
  CKC(dalloc(&h->d_uniq_pos, R * nb * 3));
  CKC(dalloc(&h->d_uniq_vel, R * nb * 3));
  CKC(dalloc(&h->d_uniq_found, R));
  CKC(dalloc(&h->d_trace_n, size_t(1)));
  CKC(dalloc(&h->d_work, size_t(1)));
#undef CKC
  *out = h;
  return 0;
}

int hmc_destroy(hmc_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  void *bufs[] = {h->d_bondmask, h->d_trans,    h->d_pairinfo,  h->d_pair_ij,    h->d_pos,        h->d_vel,
                  h->d_sbias,    h->d_config,    h->d_err,        h->d_acc,        h->d_dist,
                  h->d_dist_n,   h->d_cfg_int,   h->d_cfg_n,      h->d_hist,       h->d_uniq_pos,
                  h->d_uniq_vel, h->d_uniq_found, h->d_trace,     h->d_trace_n,    h->d_work,
                  h->d_normals,  h->d_tape_words, h->d_tape_consumed, h->d_tape_sin, h->d_tape_cos,
                  h->d_acc_rep,  h->d_skip};
  for (void *b : bufs)
    if (b) cudaFree(b);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->tm0) cudaEventDestroy(h->tm0);
  if (h->tm1) cudaEventDestroy(h->tm1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

// ---- state ----------------------------------------------------------------

int hmc_set_positions(hmc_handle h, const double *pos) {
  if (!h || !pos) return fail(h, "hmc_set_positions: null");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_pos, pos, sizeof(double) * size_t(h->R) * h->nb * 3,
                     cudaMemcpyHostToDevice, h->stream));
  return sync(h);
}
int hmc_get_positions(hmc_handle h, double *pos) {
  if (!h || !pos) return fail(h, "hmc_get_positions: null");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(pos, h->d_pos, sizeof(double) * size_t(h->R) * h->nb * 3,
                     cudaMemcpyDeviceToHost, h->stream));
  return sync(h);
}
int hmc_get_velocities(hmc_handle h, double *vel) {
  if (!h || !vel) return fail(h, "hmc_get_velocities: null");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(vel, h->d_vel, sizeof(double) * size_t(h->R) * h->nb * 3,
                     cudaMemcpyDeviceToHost, h->stream));
  return sync(h);
}
int hmc_set_config(hmc_handle h, const uint64_t *config) {
  if (!h || !config) return fail(h, "hmc_set_config: null");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_config, config, sizeof(uint64_t) * size_t(h->R), cudaMemcpyHostToDevice,
                     h->stream));
  return sync(h);
}
int hmc_get_config(hmc_handle h, uint64_t *config) {
  if (!h || !config) return fail(h, "hmc_get_config: null");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(config, h->d_config, sizeof(uint64_t) * size_t(h->R), cudaMemcpyDeviceToHost,
                     h->stream));
  return sync(h);
}

int hmc_init_config_from_positions(hmc_handle h) {
  // init_update_config, hardspheres.cc:144-171 (once per phase start; O(R*bonds))
  if (!h) return fail(h, "null handle");
  std::vector<double> pos(size_t(h->R) * h->nb * 3);
  int rc = hmc_get_positions(h, pos.data());
  if (rc) return rc;
  std::vector<uint64_t> cfg(h->R, 0);
  const double l = h->params[0].length, inv_l = 1.0 / l;
  for (int r = 0; r < h->R; r++) {
    const hmc_params &p = h->params[r / h->Rper];
    const double rc2 = p.rc * p.rc;
    const double *q = pos.data() + size_t(r) * h->nb * 3;
    uint64_t c = 0;
    // only pairs on the transient-bond list can set a bit; the first list entry of a
    // pair owns it (std::find, config.h:34)
    for (unsigned k = 0; k < p.n_transient; k++) {
      uint32_t i = p.transient_bonds[k][0], j = p.transient_bonds[k][1];
      normalise(i, j);
      if (j < i + 3 || int(j) >= h->nb) continue;
      if (bond_index(p.transient_bonds, p.n_transient, i, j) != int(k)) continue;
      double dx = q[3 * j] - q[3 * i], dy = q[3 * j + 1] - q[3 * i + 1],
             dz = q[3 * j + 2] - q[3 * i + 2];
      dx -= std::round(dx * inv_l) * l;
      dy -= std::round(dy * inv_l) * l;
      dz -= std::round(dz * inv_l) * l;
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (d2 < rc2) c ^= (uint64_t(1) << k);
    }
    cfg[r] = c;
  }
  return hmc_set_config(h, cfg.data());
}

int hmc_set_s_bias(hmc_handle h, const double *s_bias, int nstates) {
  if (!h || !s_bias) return fail(h, "hmc_set_s_bias: null");
  if (nstates != h->nstates) return fail(h, "hmc_set_s_bias: nstates mismatch");
  std::vector<double> all(size_t(h->R) * nstates);
  for (int r = 0; r < h->R; r++)
    std::memcpy(&all[size_t(r) * nstates], s_bias + size_t(r / h->Rper) * nstates,
                sizeof(double) * nstates);
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_sbias, all.data(), sizeof(double) * all.size(), cudaMemcpyHostToDevice,
                     h->stream));
  return sync(h);
}
int hmc_get_s_bias(hmc_handle h, double *s_bias, int nstates) {
  if (!h || !s_bias) return fail(h, "hmc_get_s_bias: null");
  if (nstates != h->nstates) return fail(h, "hmc_get_s_bias: nstates mismatch");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(s_bias, h->d_sbias, sizeof(double) * size_t(h->R) * nstates,
                     cudaMemcpyDeviceToHost, h->stream));
  return sync(h);
}

int hmc_reset_clocks(hmc_handle h) {
  if (!h) return fail(h, "null handle");
  int rc = sync(h);
  h->t_now = 0.0;
  return rc;
}

int hmc_reset_counts(hmc_handle h) {
  if (!h) return fail(h, "null handle");
  CK(cudaSetDevice(h->device));
  // config_count only; formed/broken/events keep accumulating like CountBond does
  CK(cudaMemsetAsync(h->d_acc, 0, sizeof(unsigned long long) * size_t(h->T) * h->nstates,
                     h->stream));
  if (h->d_acc_rep)
    CK(cudaMemsetAsync(h->d_acc_rep, 0, sizeof(unsigned long long) * size_t(h->R) * h->nstates,
                       h->stream));
  if (h->d_dist_n) CK(cudaMemsetAsync(h->d_dist_n, 0, sizeof(uint32_t) * h->R, h->stream));
  if (h->d_cfg_n) CK(cudaMemsetAsync(h->d_cfg_n, 0, sizeof(uint32_t) * h->R, h->stream));
  // distance histograms keep accumulating over rounds, like the reference's `dist`
  // dataset; hmc_enable_dist_hist re-zeroes them
  return sync(h);
}

int hmc_seed(hmc_handle h, uint32_t seed0, uint32_t seed1) {
  if (!h) return fail(h, "null handle");
  h->seed0 = seed0;
  h->seed1 = seed1;
  h->md_draws = 0;
  h->mc_draws = 0;
  return 0;
}

// ---- runs -------------------------------------------------------------------

int hmc_run(hmc_handle h, int phase, unsigned first_iter, unsigned n_iters) {
  if (!h) return fail(h, "null handle");
  if (phase < HMC_PHASE_EQ || phase > HMC_PHASE_MAIN) return fail(h, "hmc_run: bad phase");
  if (n_iters == 0) return 0;
  const hmc_params &p = h->params[0];
  {
    // the first step must not lie before the bead clocks: the reference would integrate
    // backwards in time there (run_step, main.cc:104,135-138) and leave every well
    const double first_t = phase == HMC_PHASE_WL ? double(first_iter * p.nsteps_wl) * p.del_t_wl
                           : double(first_iter * (phase == HMC_PHASE_EQ ? p.nsteps_eq : p.nsteps)) * p.del_t;
    if (first_t < h->t_now)
      return fail(h, "hmc_run: step time before the bead clocks; call hmc_reset_clocks first");
  }
  KArgs a;
  fill_kargs(h, a);
  a.op = HMC_OP_RUN;
  a.phase = phase;
  a.first = first_iter;
  a.count = n_iters;
  a.rng_mode = HMC_RNG_PHILOX;
  int rc = launch(h, a);
  if (rc) return rc;
  if (phase == HMC_PHASE_EQ) {
    h->md_draws += (unsigned long long)n_iters * p.nsteps_eq;
    h->t_now = double((first_iter + n_iters) * p.nsteps_eq - 1) * p.del_t;
  } else if (phase == HMC_PHASE_MAIN) {
    h->md_draws += (unsigned long long)n_iters * p.nsteps;
    h->mc_draws += n_iters;
    h->t_now = double((first_iter + n_iters) * p.nsteps - 1) * p.del_t;
  } else {
    h->md_draws += (unsigned long long)n_iters * h->nstates;
    h->t_now = double((first_iter + n_iters) * p.nsteps_wl - 1) * p.del_t_wl;
  }
  return 0;
}

int hmc_sync(hmc_handle h) {
  if (!h) return fail(h, "null handle");
  return sync(h);
}

int hmc_md_steps_injected(hmc_handle h, int phase, unsigned first_step, unsigned nsteps,
                          const double *normals) {
  if (!h || !normals) return fail(h, "hmc_md_steps_injected: null");
  if (phase < HMC_PHASE_EQ || phase > HMC_PHASE_MAIN) return fail(h, "bad phase");
  if (nsteps == 0) return 0;
  const hmc_params &p = h->params[0];
  {
    const double first_t = phase == HMC_PHASE_WL
                               ? double((first_step / unsigned(h->nstates)) * p.nsteps_wl) * p.del_t_wl
                               : double(first_step) * p.del_t;
    if (first_t < h->t_now)
      return fail(h, "hmc_md_steps_injected: step time before the bead clocks; call hmc_reset_clocks first");
  }
  const size_t n = size_t(h->R) * nsteps * h->nb * 3;
  CK(cudaSetDevice(h->device));
  if (n > h->normals_cap) {
    CK(cudaStreamSynchronize(h->stream));
    if (h->d_normals) cudaFree(h->d_normals);
    h->d_normals = nullptr;
    CK(cudaMalloc(reinterpret_cast<void **>(&h->d_normals), n * sizeof(double)));
    h->normals_cap = n;
  }
  CK(cudaMemcpyAsync(h->d_normals, normals, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  KArgs a;
  fill_kargs(h, a);
  a.op = HMC_OP_MD_ONLY;
  a.phase = phase;
  a.first = first_step;
  a.count = nsteps;
  a.rng_mode = HMC_RNG_INJECTED;
  a.normals = h->d_normals;
  int rc = launch(h, a);
  if (rc) return rc;
  if (phase == HMC_PHASE_WL) {
    const unsigned iter_wl = (first_step + nsteps - 1) / unsigned(h->nstates);
    h->t_now = double((iter_wl + 1) * p.nsteps_wl - 1) * p.del_t_wl;
  } else {
    h->t_now = double(first_step + nsteps - 1) * p.del_t;
  }
  return sync(h); // the host buffer may be reused by the caller
}

int hmc_crankshaft_injected(hmc_handle h, unsigned move_offset, unsigned n_moves,
                            const uint32_t *words, const double *sin_tab, const double *cos_tab,
                            unsigned n_words, uint32_t *consumed) {
  if (!h || !words || !sin_tab || !cos_tab) return fail(h, "hmc_crankshaft_injected: null");
  if (n_moves == 0) return 0;
  const size_t n = size_t(h->R) * n_words;
  CK(cudaSetDevice(h->device));
  if (n > h->tape_cap) {
    CK(cudaStreamSynchronize(h->stream));
    if (h->d_tape_words) cudaFree(h->d_tape_words);
    if (h->d_tape_sin) cudaFree(h->d_tape_sin);
    if (h->d_tape_cos) cudaFree(h->d_tape_cos);
    h->d_tape_words = nullptr;
    h->d_tape_sin = h->d_tape_cos = nullptr;
    CK(cudaMalloc(reinterpret_cast<void **>(&h->d_tape_words), n * sizeof(uint32_t)));
    CK(cudaMalloc(reinterpret_cast<void **>(&h->d_tape_sin), n * sizeof(double)));
    CK(cudaMalloc(reinterpret_cast<void **>(&h->d_tape_cos), n * sizeof(double)));
    h->tape_cap = n;
  }
  if (!h->d_tape_consumed) CK(dalloc(&h->d_tape_consumed, 2 * size_t(h->R)));
  CK(cudaMemcpyAsync(h->d_tape_words, words, n * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_tape_sin, sin_tab, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_tape_cos, cos_tab, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  KArgs a;
  fill_kargs(h, a);
  a.op = HMC_OP_CRANK_ONLY;
  a.phase = HMC_PHASE_MAIN;
  a.count = n_moves;
  a.move_offset = move_offset;
  a.rng_mode = HMC_RNG_INJECTED;
  a.tape_words = h->d_tape_words;
  a.tape_sin = h->d_tape_sin;
  a.tape_cos = h->d_tape_cos;
  a.tape_n = n_words;
  a.tape_consumed = h->d_tape_consumed;
  int rc = launch(h, a);
  if (rc) return rc;
  if (consumed)
    CK(cudaMemcpyAsync(consumed, h->d_tape_consumed, sizeof(uint32_t) * h->R, cudaMemcpyDeviceToHost,
                       h->stream));
  return sync(h);
}

int hmc_get_moves_done(hmc_handle h, uint32_t *moves) {
  if (!h || !moves) return fail(h, "hmc_get_moves_done: null");
  if (!h->d_tape_consumed) return fail(h, "hmc_get_moves_done: no crankshaft call yet");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(moves, h->d_tape_consumed + h->R, sizeof(uint32_t) * h->R,
                     cudaMemcpyDeviceToHost, h->stream));
  return sync(h);
}

// ---- results ----------------------------------------------------------------

int hmc_get_counts(hmc_handle h, uint64_t *config_count, int nstates, uint64_t *formed,
                   uint64_t *broken, uint32_t *err_flags) {
  if (!h) return fail(h, "null handle");
  if (config_count && nstates != h->nstates) return fail(h, "hmc_get_counts: nstates mismatch");
  std::vector<unsigned long long> acc(h->n_acc);
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(acc.data(), h->d_acc, sizeof(unsigned long long) * h->n_acc,
                     cudaMemcpyDeviceToHost, h->stream));
  if (err_flags)
    CK(cudaMemcpyAsync(err_flags, h->d_err, sizeof(uint32_t) * h->R, cudaMemcpyDeviceToHost,
                       h->stream));
  int rc = sync(h);
  if (rc) return rc;
  const size_t base = size_t(h->T) * h->nstates;
  if (config_count)
    for (size_t k = 0; k < base; k++) config_count[k] = acc[k];
  for (int t = 0; t < h->T; t++) {
    if (formed) formed[t] = acc[base + t];
    if (broken) broken[t] = acc[base + h->T + t];
  }
  return 0;
}

int hmc_set_replica_ids(hmc_handle h, unsigned first, unsigned total_per_transition) {
  if (!h) return fail(h, "null handle");
  if (total_per_transition && first + unsigned(h->Rper) > total_per_transition)
    return fail(h, "hmc_set_replica_ids: first + replicas_per_transition exceeds the total");
  h->rng_first = first;
  h->rng_total = total_per_transition;
  return 0;
}

int hmc_set_list_capacity(hmc_handle h, unsigned entries) {
  if (!h) return fail(h, "null handle");
  h->list_cap_override = int(entries);
  return 0;
}

unsigned hmc_get_list_capacity(hmc_handle h) {
  return h ? unsigned(hmc_list_capacity(h->nb, h->group, h->list_cap_override)) : 0u;
}

int hmc_enable_replica_counts(hmc_handle h, int on) {
  if (!h) return fail(h, "null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (on && !h->d_acc_rep) {
    CK(dalloc(&h->d_acc_rep, size_t(h->R) * h->nstates));
  } else if (!on && h->d_acc_rep) {
    cudaFree(h->d_acc_rep);
    h->d_acc_rep = nullptr;
  }
  return 0;
}

int hmc_get_replica_counts(hmc_handle h, uint64_t *config_count) {
  if (!h || !config_count) return fail(h, "hmc_get_replica_counts: null");
  if (!h->d_acc_rep) return fail(h, "hmc_get_replica_counts: not enabled");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(config_count, h->d_acc_rep, sizeof(uint64_t) * size_t(h->R) * h->nstates,
                     cudaMemcpyDeviceToHost, h->stream));
  return sync(h);
}

int hmc_set_replica_skip(hmc_handle h, const uint8_t *skip) {
  if (!h) return fail(h, "null handle");
  CK(cudaSetDevice(h->device));
  if (!skip) {
    h->skip_on = false;
    return 0;
  }
  if (!h->d_skip) CK(dalloc(&h->d_skip, size_t(h->R)));
  CK(cudaMemcpyAsync(h->d_skip, skip, size_t(h->R), cudaMemcpyHostToDevice, h->stream));
  h->skip_on = true;
  return sync(h);
}

int hmc_allreduce_counts(hmc_handle h, hmc_comm c) {
  if (!h) return fail(h, "null handle");
  if (!c || hmc_comm_world(c) == 1) return 0;
  CK(cudaSetDevice(h->device));
  if (hmc_comm_allreduce_device_u64(c, h->d_acc, uint64_t(h->T) * h->nstates, h->stream) != 0)
    return fail(h, std::string("hmc_allreduce_counts: ") + hmc_last_error(nullptr), -7);
  return 0;
}

int hmc_allreduce_totals(hmc_handle h, hmc_comm c) {
  if (!h) return fail(h, "null handle");
  if (!c || hmc_comm_world(c) == 1) return 0;
  CK(cudaSetDevice(h->device));
  const uint64_t off = uint64_t(h->T) * h->nstates;
  if (hmc_comm_allreduce_device_u64(c, h->d_acc + off, h->n_acc - off, h->stream) != 0)
    return fail(h, std::string("hmc_allreduce_totals: ") + hmc_last_error(nullptr), -7);
  return 0;
}

int hmc_allreduce_hist(hmc_handle h, hmc_comm c) {
  if (!h) return fail(h, "null handle");
  if (!h->d_hist) return fail(h, "hmc_allreduce_hist: histograms are not enabled");
  if (!c || hmc_comm_world(c) == 1) return 0;
  CK(cudaSetDevice(h->device));
  const uint64_t n = uint64_t(h->T) * (h->n_nonlocal_max ? h->n_nonlocal_max : 1) * 2 * h->hist_bins;
  if (hmc_comm_allreduce_device_u64(c, h->d_hist, n, h->stream) != 0)
    return fail(h, std::string("hmc_allreduce_hist: ") + hmc_last_error(nullptr), -7);
  return 0;
}

int hmc_counts_device_ptr(hmc_handle h, void **dptr, uint64_t *n_u64) {
  if (!h || !dptr || !n_u64) return fail(h, "hmc_counts_device_ptr: null");
  *dptr = h->d_acc;
  *n_u64 = h->n_acc;
  return 0;
}

int hmc_get_event_counts(hmc_handle h, uint64_t events[4]) {
  if (!h || !events) return fail(h, "hmc_get_event_counts: null");
  unsigned long long ev[4];
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(ev, h->d_acc + (h->n_acc - 4), sizeof ev, cudaMemcpyDeviceToHost, h->stream));
  int rc = sync(h);
  for (int k = 0; k < 4; k++) events[k] = ev[k];
  return rc;
}

int hmc_enable_samples(hmc_handle h, unsigned dist_rows, unsigned cfg_cap) {
  if (!h) return fail(h, "null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->d_dist) cudaFree(h->d_dist);
  if (h->d_cfg_int) cudaFree(h->d_cfg_int);
  h->d_dist = nullptr;
  h->d_cfg_int = nullptr;
  h->dist_cap = 0;
  h->cfg_cap = 0;
  const size_t stride = h->n_nonlocal_max ? h->n_nonlocal_max : 1;
  if (dist_rows) {
    CK(dalloc(&h->d_dist, size_t(h->R) * dist_rows * stride));
    if (!h->d_dist_n) CK(dalloc(&h->d_dist_n, size_t(h->R)));
    h->dist_cap = dist_rows;
  }
  if (cfg_cap) {
    CK(dalloc(&h->d_cfg_int, size_t(h->R) * cfg_cap));
    if (!h->d_cfg_n) CK(dalloc(&h->d_cfg_n, size_t(h->R)));
    h->cfg_cap = cfg_cap;
  }
  // (on the engine's stream: a default-stream memset is not ordered with it, see dalloc)
  if (h->d_dist_n) CK(cudaMemsetAsync(h->d_dist_n, 0, sizeof(uint32_t) * h->R, h->stream));
  if (h->d_cfg_n) CK(cudaMemsetAsync(h->d_cfg_n, 0, sizeof(uint32_t) * h->R, h->stream));
  return sync(h);
}

int hmc_get_dist(hmc_handle h, double *dist, uint32_t *n_rows) {
  if (!h) return fail(h, "null handle");
  if (!h->dist_cap) return fail(h, "hmc_get_dist: samples not enabled");
  const size_t stride = h->n_nonlocal_max ? h->n_nonlocal_max : 1;
  CK(cudaSetDevice(h->device));
  if (dist)
    CK(cudaMemcpyAsync(dist, h->d_dist, sizeof(double) * size_t(h->R) * h->dist_cap * stride,
                       cudaMemcpyDeviceToHost, h->stream));
  if (n_rows)
    CK(cudaMemcpyAsync(n_rows, h->d_dist_n, sizeof(uint32_t) * h->R, cudaMemcpyDeviceToHost,
                       h->stream));
  return sync(h);
}

int hmc_get_config_int(hmc_handle h, uint64_t *config_int, uint32_t *n) {
  if (!h) return fail(h, "null handle");
  if (!h->cfg_cap) return fail(h, "hmc_get_config_int: samples not enabled");
  CK(cudaSetDevice(h->device));
  if (config_int)
    CK(cudaMemcpyAsync(config_int, h->d_cfg_int, sizeof(uint64_t) * size_t(h->R) * h->cfg_cap,
                       cudaMemcpyDeviceToHost, h->stream));
  if (n)
    CK(cudaMemcpyAsync(n, h->d_cfg_n, sizeof(uint32_t) * h->R, cudaMemcpyDeviceToHost, h->stream));
  return sync(h);
}

int hmc_enable_dist_hist(hmc_handle h, unsigned nbins, double rmax) {
  if (!h) return fail(h, "null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->d_hist) cudaFree(h->d_hist);
  h->d_hist = nullptr;
  h->hist_bins = 0;
  if (nbins) {
    if (!(rmax > 0)) return fail(h, "hmc_enable_dist_hist: rmax must be > 0");
    const size_t stride = h->n_nonlocal_max ? h->n_nonlocal_max : 1;
    CK(dalloc(&h->d_hist, size_t(h->T) * stride * 2 * nbins));
    h->hist_bins = nbins;
    h->hist_rmax = rmax;
  }
  return 0;
}

int hmc_get_dist_hist(hmc_handle h, uint64_t *hist) {
  if (!h || !hist) return fail(h, "hmc_get_dist_hist: null");
  if (!h->hist_bins) return fail(h, "hmc_get_dist_hist: histogram not enabled");
  const size_t stride = h->n_nonlocal_max ? h->n_nonlocal_max : 1;
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(hist, h->d_hist, sizeof(uint64_t) * size_t(h->T) * stride * 2 * h->hist_bins,
                     cudaMemcpyDeviceToHost, h->stream));
  return sync(h);
}

int hmc_dist_hist_device_ptr(hmc_handle h, void **dptr, uint64_t *n_u64) {
  if (!h || !dptr || !n_u64) return fail(h, "hmc_dist_hist_device_ptr: null");
  if (!h->hist_bins) return fail(h, "hmc_dist_hist_device_ptr: histogram not enabled");
  const size_t stride = h->n_nonlocal_max ? h->n_nonlocal_max : 1;
  *dptr = h->d_hist;
  *n_u64 = uint64_t(h->T) * stride * 2 * h->hist_bins;
  return 0;
}

int hmc_get_unique_config(hmc_handle h, double *pos, double *vel, uint8_t *found) {
  if (!h) return fail(h, "null handle");
  const size_t n = size_t(h->R) * h->nb * 3;
  CK(cudaSetDevice(h->device));
  if (pos) CK(cudaMemcpyAsync(pos, h->d_uniq_pos, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  if (vel) CK(cudaMemcpyAsync(vel, h->d_uniq_vel, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  if (found) CK(cudaMemcpyAsync(found, h->d_uniq_found, size_t(h->R), cudaMemcpyDeviceToHost, h->stream));
  return sync(h);
}

int hmc_enable_trace(hmc_handle h, int replica, unsigned capacity) {
  if (!h) return fail(h, "null handle");
  if (capacity && (replica < 0 || replica >= h->R)) return fail(h, "hmc_enable_trace: bad replica");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->d_trace) cudaFree(h->d_trace);
  h->d_trace = nullptr;
  h->trace_cap = 0;
  h->trace_replica = -1;
  CK(cudaMemsetAsync(h->d_trace_n, 0, sizeof(uint32_t), h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (capacity) {
    CK(dalloc(&h->d_trace, size_t(capacity)));
    h->trace_cap = capacity;
    h->trace_replica = replica;
  }
  return 0;
}

int hmc_get_trace(hmc_handle h, hmc_event *out, unsigned max_events, unsigned *n_events) {
  if (!h) return fail(h, "null handle");
  uint32_t n = 0;
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(&n, h->d_trace_n, sizeof n, cudaMemcpyDeviceToHost, h->stream));
  int rc = sync(h);
  if (rc) return rc;
  if (n_events) *n_events = n;
  unsigned k = n < h->trace_cap ? n : h->trace_cap;
  if (k > max_events) k = max_events;
  if (out && k) {
    CK(cudaMemcpyAsync(out, h->d_trace, sizeof(hmc_event) * k, cudaMemcpyDeviceToHost, h->stream));
    rc = sync(h);
  }
  return rc;
}

// ---- entropy (host, O(nstates)) -------------------------------------------------

int hmc_compute_entropy(const uint64_t *cc, double *sb, int n, double pos_scale,
                        double neg_scale) {
  if (!cc || !sb || n < 1) return -1;
  const double norm = compute_norm(cc, n);
  const int nat = n - 1;
  const double native_count = double(cc[nat]);
  double s_native;
  if (native_count > 0) s_native = sb[nat] + pos_scale * std::log(native_count / norm);
  else s_native = sb[nat] - neg_scale * std::log(norm);
  for (int i = 0; i < n; i++) {
    double si;
    if (cc[i] > 0) si = sb[i] + pos_scale * std::log(cc[i] / norm);
    else si = sb[i] - neg_scale * std::log(norm);
    sb[i] = si - s_native;
  }
  return 0;
}

double hmc_compute_g_val(const uint64_t *cc, int n) {
  const double norm = compute_norm(cc, n);
  double g = 0.0;
  for (int i = 0; i < n; i++)
    if (cc[i] > 0) g += cc[i] * std::log(cc[i] / norm);
  g *= 2.0;
  return g;
}

int hmc_g_test(const uint64_t *cc, int n, double sig, double *g_val, double *g_crit,
               double *p_val) {
  const double dof = n - 1;
  const double g = hmc_compute_g_val(cc, n);
  const double pv = 1.0 - gamma_p(0.5 * dof, 0.5 * g);
  double lo = 0.0, hi = 1.0;
  while (1.0 - gamma_p(0.5 * dof, 0.5 * hi) > sig) hi *= 2.0;
  for (int k = 0; k < 200; k++) {
    const double mid = 0.5 * (lo + hi);
    if (1.0 - gamma_p(0.5 * dof, 0.5 * mid) > sig) lo = mid;
    else hi = mid;
  }
  const double crit = 0.5 * (lo + hi);
  if (g_val) *g_val = g;
  if (g_crit) *g_crit = crit;
  if (p_val) *p_val = pv;
  return (g < crit) ? 1 : 0;
}

int hmc_get_schedule_state(hmc_handle h, hmc_schedule_state *out) {
  if (!h || !out) return fail(h, "hmc_get_schedule_state: null");
  const int rc = sync(h);
  out->t_now = h->t_now;
  out->md_draws = h->md_draws;
  out->mc_draws = h->mc_draws;
  out->seed0 = h->seed0;
  out->seed1 = h->seed1;
  return rc;
}

int hmc_set_schedule_state(hmc_handle h, const hmc_schedule_state *in) {
  if (!h || !in) return fail(h, "hmc_set_schedule_state: null");
  const int rc = sync(h);
  h->t_now = in->t_now;
  h->md_draws = in->md_draws;
  h->mc_draws = in->mc_draws;
  h->seed0 = in->seed0;
  h->seed1 = in->seed1;
  return rc;
}

int hmc_set_s_bias_per_replica(hmc_handle h, const double *s_bias, int nstates) {
  if (!h || !s_bias) return fail(h, "hmc_set_s_bias_per_replica: null");
  if (nstates != h->nstates) return fail(h, "hmc_set_s_bias_per_replica: nstates mismatch");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_sbias, s_bias, sizeof(double) * size_t(h->R) * nstates,
                     cudaMemcpyHostToDevice, h->stream));
  return sync(h);
}

int hmc_get_launch_stats(hmc_handle h, uint64_t *n_launches, float *last_ms) {
  if (!h) return fail(h, "null handle");
  int rc = sync(h);
  if (n_launches) *n_launches = h->n_launches;
  if (last_ms) *last_ms = h->last_ms;
  return rc;
}

int hmc_resident_replicas(unsigned nbeads, int device) {
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  // 16 warps per SM (the kernel's register budget), 32 / G replicas per warp
  return sms * 16 * (32 / hmc_pick_group(int(nbeads)));
}

long long hmc_debug_bounds_violations(void) {
  cudaDeviceSynchronize();
  return hmc_bounds_violations_device();
}

int hmc_timer_start(hmc_handle h) {
  if (!h) return fail(h, "null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->tm0, h->stream));
  return 0;
}

int hmc_timer_stop(hmc_handle h, float *ms) {
  if (!h || !ms) return fail(h, "hmc_timer_stop: null");
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->tm1, h->stream));
  CK(cudaEventSynchronize(h->tm1));
  CK(cudaEventElapsedTime(ms, h->tm0, h->tm1));
  return 0;
}

unsigned hmc_wl_outer_iterations(const hmc_params *p) { return wl_outer_iterations(*p); }

} // extern "C"
