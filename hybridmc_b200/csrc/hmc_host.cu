// hmc_host.cu — host-side pieces of the path that run once per run (not per event):
// initial chain growth (init_pos) with the reference's RNG contract, and the FP64
// issue-rate microbenchmark that provides the roofline denominator.
#include "hmc_internal.h"

#include <cmath>
#include <random>
#include <vector>

namespace {

// Box::mindist, box.h:22-26
inline void mindist(double l, double inv_l, double &dx, double &dy, double &dz) {
  dx -= std::round(dx * inv_l) * l;
  dy -= std::round(dy * inv_l) * l;
  dz -= std::round(dz * inv_l) * l;
}

bool has_bond(const uint32_t (*b)[2], unsigned n, unsigned i, unsigned j) {
  for (unsigned k = 0; k < n; k++) {
    unsigned a = b[k][0], c = b[k][1];
    if (a > c) std::swap(a, c);
    if (a == i && c == j) return true;
  }
  return false;
}

// unit_sphere, hardspheres.cc:19-34 (Marsaglia 1972)
void unit_sphere(std::mt19937 &mt, double &x, double &y, double &z) {
  std::uniform_real_distribution<> uni(-1.0, 1.0);
  double v, w, s;
  do {
    v = uni(mt);
    w = uni(mt);
    s = v * v + w * w;
  } while (!(s < 1));
  x = 2 * v * std::sqrt(1 - v * v - w * w);
  y = 2 * w * std::sqrt(1 - v * v - w * w);
  z = 1 - 2 * (v * v + w * w);
}

// FP64 issue-rate probes: 8 independent dependent chains per thread.
__global__ void fp64_fma_probe(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000000001, c = 1e-12;
  for (int i = 0; i < iters; i++) {
    a0 = __fma_rn(a0, m, c);
    a1 = __fma_rn(a1, m, c);
    a2 = __fma_rn(a2, m, c);
    a3 = __fma_rn(a3, m, c);
    a4 = __fma_rn(a4, m, c);
    a5 = __fma_rn(a5, m, c);
    a6 = __fma_rn(a6, m, c);
    a7 = __fma_rn(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void fp64_muladd_probe(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000000001, c = 1e-12;
  for (int i = 0; i < iters; i++) {
    a0 = __dadd_rn(__dmul_rn(a0, m), c);
    a1 = __dadd_rn(__dmul_rn(a1, m), c);
    a2 = __dadd_rn(__dmul_rn(a2, m), c);
    a3 = __dadd_rn(__dmul_rn(a3, m), c);
    a4 = __dadd_rn(__dmul_rn(a4, m), c);
    a5 = __dadd_rn(__dmul_rn(a5, m), c);
    a6 = __dadd_rn(__dmul_rn(a6, m), c);
    a7 = __dadd_rn(__dmul_rn(a7, m), c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

} // namespace

extern "C" {

// init_pos (hardspheres.cc:37-138) for one chain with the given engine.
// Returns 0, or -1 when `tries` placements of one bead fail (the reference throws
// std::logic_error there).
int hmc_init_pos_chain(const hmc_params *p, const uint32_t *seeds, unsigned n_seeds, double *pos) {
  std::seed_seq seq(seeds, seeds + n_seeds); // main.cc:379-380
  std::mt19937 mt(seq);
  return hmc_init_pos_engine(p, &mt, pos);
}

int hmc_init_pos_engine(const hmc_params *p, void *engine, double *pos) {
  std::mt19937 &mt = *static_cast<std::mt19937 *>(engine);
  const unsigned nb = p->nbeads;
  const double l = p->length, inv_l = 1.0 / l;
  const double nnear_min2 = p->nnear_min * p->nnear_min, nnear_max2 = p->nnear_max * p->nnear_max;
  const double rh2 = p->rh * p->rh, rc2 = p->rc * p->rc;
  std::uniform_real_distribution<> uniform_near(0.0, 1.0);
  if (nb < 1) return 0;
  pos[0] = pos[1] = pos[2] = 0.0;
  uint64_t tries = 0;
  unsigned i = 1;
  const double near_min3 = std::pow(p->near_min, 3.0), near_max3 = std::pow(p->near_max, 3.0);
  while (i < nb) {
    if (!(tries++ < p->tries)) return -1;
    double ux, uy, uz;
    unit_sphere(mt, ux, uy, uz);
    const double val = uniform_near(mt);
    const double near = std::cbrt(near_min3 + (val * (near_max3 - near_min3)));
    const double x = near * ux + pos[3 * (i - 1)];
    const double y = near * uy + pos[3 * (i - 1) + 1];
    const double z = near * uz + pos[3 * (i - 1) + 2];
    if (i >= 2) {
      double dx = x - pos[3 * (i - 2)], dy = y - pos[3 * (i - 2) + 1], dz = z - pos[3 * (i - 2) + 2];
      mindist(l, inv_l, dx, dy, dz);
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (d2 < nnear_min2 || d2 > nnear_max2) continue;
    }
    bool overlap = false;
    for (unsigned j = 0; j < i; j++) {
      double dx = x - pos[3 * j], dy = y - pos[3 * j + 1], dz = z - pos[3 * j + 2];
      mindist(l, inv_l, dx, dy, dz);
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (i >= j + 3 && d2 < rh2) {
        overlap = true;
        break;
      }
      if (has_bond(p->transient_bonds, p->n_transient, j, i) && d2 < rc2) {
        overlap = true;
        break;
      }
    }
    if (overlap) continue;
    pos[3 * i] = x;
    pos[3 * i + 1] = y;
    pos[3 * i + 2] = z;
    i++;
    tries = 0;
  }
  return 0;
}

// Measured FP64 issue rates of this GPU (TFLOP/s, FMA = 2 flops; mul+add = 2 flops
// issued as two instructions — the ceiling of a -fmad=false kernel).
int hmc_fp64_peak(int device, double *fma_tflops, double *muladd_tflops) {
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int block = 256, grid = sms * 8, iters = 1 << 15;
  double *out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * size_t(block) * grid) != cudaSuccess) return -2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best[2] = {0, 0};
  for (int which = 0; which < 2; which++) {
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      if (which == 0) fp64_fma_probe<<<grid, block>>>(out, iters);
      else fp64_muladd_probe<<<grid, block>>>(out, iters);
      cudaEventRecord(e1);
      if (cudaEventSynchronize(e1) != cudaSuccess) {
        cudaFree(out);
        return -3;
      }
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      const double flops = 2.0 * 8.0 * double(iters) * double(block) * double(grid);
      const double tf = flops / (double(ms) * 1e-3) * 1e-12;
      if (rep > 0 && tf > best[which]) best[which] = tf;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (fma_tflops) *fma_tflops = best[0];
  if (muladd_tflops) *muladd_tflops = best[1];
  return 0;
}

} // extern "C"
