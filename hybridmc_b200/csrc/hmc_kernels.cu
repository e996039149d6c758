// hmc_kernels.cu — sm_100a FP64 ensemble kernel for HybridMC's event-driven hybrid
// Monte Carlo (reference hot path: src/main.cc:45-305 and the hardspheres.cc /
// crankshaft.cc functions it calls).
//
// Design (B200-first, not a translation of the reference's heap + cell lists):
//   * one GROUP of G lanes (G = 8, 16 or 32, a sub-warp) owns one replica for the
//     whole launch; the 32/G groups of a warp run in LOCKSTEP (every branch that
//     contains a shuffle or sync is warp-uniform, finished groups are predicated off),
//     because partial-mask __syncwarp/__shfl_sync compile to a slow MATCH/REDUX/BRA.DIV
//     sequence on sm_100a.  Warps pull batches of replicas from a global work counter,
//     so there is no barrier between warps (event counts per step vary by 10-70 %).
//   * the replica lives in shared memory as bead records (x,y,z,vx,vy,vz,clock, packed
//     3-D cell index) and an EVENT TABLE that holds, per bead pair, its next predicted
//     collision with the event type latched at prediction time, and per bead its next
//     cell crossing.  It replaces the reference's std::priority_queue + per-bead
//     collision counters (src/event.h:15-116).  The table is stored in two parts:
//       - DENSE slots for the events that always exist: one cell-wall time per bead and
//         one time per local pair (k,k+1), (k,k+2) — 3 nb - 3 slots, inside / next to
//         the bead records;
//       - a compact LIST of the nonlocal pairs that currently have an event (time +
//         key = pair and latched type).  Of the nb^2/2 nonlocal pairs only ~0.5 nb are
//         live at any time (measured: 21 of 946 for crambin, max 70 in 4e5 events), so
//         the list is scanned in one or two lane strides where a dense pair table
//         needs nb^2/2G.
//     Semantics:
//       - executing a collision (i,j) deletes every list entry touching i or j, rewrites
//         their local-pair slots and appends the new predictions that have an event
//         (= the counter bump that invalidates queued events, hardspheres.cc:1118-1119)
//       - executing a cell crossing of i only min-merges predictions for the beads
//         in the 9 newly adjacent cells and replaces i's cell slot; all other entries
//         touching i are kept (no counter bump, hardspheres.cc:1622-1645)
//       - the next event is a lane-parallel argmin over dense slots + list and a
//         shuffle reduce; ties go to the lowest (cells, then pairs in (i,j) order) key.
//     SURVEY.md §7a shows this is bit-identical to the heap; it does none of the
//     87-90 % stale-pop work of the reference.
//   * predictions are made exactly when and for whom the reference makes them
//     (27-cell neighbourhood at a collision, 9 new cells at a crossing; with
//     ncell >= 4 cell adjacency is a per-pair predicate on the cell indices, so no
//     per-cell lists are needed), lane-parallel over partner beads.
//   * all FP64 arithmetic keeps the reference's IEEE operation order; the file is
//     compiled with -fmad=false so nothing is contracted into FMAs, and uses IEEE
//     sqrt / division and round-half-away (round()), so event times are
//     bit-identical to the CPU reference.
//   * per-replica counter-based Philox4x32-10 RNG (production) or a host-injected
//     stream (reference-exact mode).
// No tensor cores: nothing here is a dense contraction.
#include "hmc_internal.h"

#include <math_constants.h>

// register budget: HMC_BLOCK threads x HMC_MIN_BLOCKS CTAs per SM
#ifndef HMC_BLOCK
#define HMC_BLOCK 128
#endif
#ifndef HMC_MIN_BLOCKS
#define HMC_MIN_BLOCKS 4
#endif

// The once-per-interval pieces (initialize_system, sampling, crankshaft) are inlined into
// the kernel; -DHMC_OUTLINE_COLD keeps them out of line (measured on B200: 2-4 % slower,
// the crankshaft block 3x slower — profiles/README.md).
#ifdef HMC_OUTLINE_COLD
#define HMC_COLD __noinline__
#else
#define HMC_COLD __forceinline__
#endif

// -DHMC_PF_UNROLL=n: unroll factor of the pre-filter loop (experiment switch)
#define HMC_STR_(x) #x
#define HMC_PRAGMA_(x) _Pragma(HMC_STR_(x))
#ifdef HMC_PF_UNROLL
#define HMC_PF_UNROLL_PRAGMA HMC_PRAGMA_(unroll HMC_PF_UNROLL)
#else
#define HMC_PF_UNROLL_PRAGMA
#endif

// -DHMC_DEBUG_BOUNDS: every index into the shared-memory replica image (bead records, the
// event list, the partner list) is checked; a violation is counted (hmc_debug_bounds_violations)
// and redirected to element 0.  The substitute for compute-sanitizer where that is not
// available; tests/test_gpu_debug_builds.py runs the shipped configurations under it.
#ifdef HMC_DEBUG_BOUNDS
__device__ unsigned long long g_bounds_violations = 0ull;
__device__ __forceinline__ int hmc_bchk(int i, int n) {
  if (unsigned(i) >= unsigned(n)) {
    atomicAdd(&g_bounds_violations, 1ull);
    return 0;
  }
  return i;
}
#define HMC_IDX(i, n) hmc_bchk(int(i), int(n))
#else
#define HMC_IDX(i, n) (i)
#endif

namespace {

// ----------------------------------------------------------------- helpers --

// Shared-memory image of one replica.
//   beads: array-of-structs, BSTRIDE doubles per bead:
//            x y z vx vy vz clock [cell | wall] t_cell t_near
//          one address computation per bead and 128-bit loads fetch the whole record; a
//          stride of 80 B keeps every quarter-warp of such loads conflict-free
//          (20 k mod 32 covers all banks for 8 consecutive k).
//          word 0 of slot 7: packed cell x | y<<8 | z<<16 | wall<<24 (the wall latched
//          with the bead's cell-crossing time t_cell); word 1: one-hot image of the cell
//          for the adjacency test.  t_near: event of the pair (k,k+1).
//          (The latched type of a local-well event is not stored: at the event the pair
//          sits on the inner or on the outer wall of its well, which tells them apart.)
//   tnn:   event time of the pair (k,k+2).
//   le:    the list of live nonlocal predictions, 16 B each: time, key, pad; slots >= the
//          list length hold +inf.  key = 1<<24 | lo<<16 | hi<<8 | type; a cell event has
//          key k<<16.
// The arrays sit at launch-uniform byte offsets (KArgs::off_*) from the replica's base.
constexpr int BSTRIDE = 10;
enum { F_X = 0, F_Y = 1, F_Z = 2, F_VX = 3, F_VY = 4, F_VZ = 5, F_T = 6, F_CELL = 7, F_TC = 8, F_TN = 9 };
constexpr unsigned PAIR_KEY = 0x01000000u;
constexpr int PINFO_STRIDE = 64; // pairinfo is a [nb][64] matrix per transition

struct Smem {
  double *b;   // [nb][BSTRIDE]
  double *tnn; // [nb]
  uint4 *le;   // [list_cap]
  uint8_t *plist; // [2 nb] + scratch: partners that survive the single-precision pre-filter
  int nb;
#ifdef HMC_DEBUG_BOUNDS
  int cap; // list capacity, for the index checks
#endif
};

__device__ __forceinline__ Smem carve(unsigned char *base, const KArgs &a) {
  Smem s;
  s.b = reinterpret_cast<double *>(base);
  s.tnn = reinterpret_cast<double *>(base + a.off_tnn);
  s.le = reinterpret_cast<uint4 *>(base + a.off_list);
  s.plist = base + a.off_plist;
  s.nb = a.nb;
#ifdef HMC_DEBUG_BOUNDS
  s.cap = a.list_cap;
#endif
  return s;
}

__device__ __forceinline__ uint4 le_make(double t, unsigned key) {
  return make_uint4(unsigned(__double2loint(t)), unsigned(__double2hiint(t)), key, 0u);
}
__device__ __forceinline__ double le_time(const uint4 &v) {
  return __hiloint2double(int(v.y), int(v.x));
}

#define BX(s, k) ((s).b[BSTRIDE * HMC_IDX(k, (s).nb) + F_X])
#define BY(s, k) ((s).b[BSTRIDE * HMC_IDX(k, (s).nb) + F_Y])
#define BZ(s, k) ((s).b[BSTRIDE * HMC_IDX(k, (s).nb) + F_Z])
#define BVX(s, k) ((s).b[BSTRIDE * HMC_IDX(k, (s).nb) + F_VX])
#define BVY(s, k) ((s).b[BSTRIDE * HMC_IDX(k, (s).nb) + F_VY])
#define BVZ(s, k) ((s).b[BSTRIDE * HMC_IDX(k, (s).nb) + F_VZ])
#define BT(s, k) ((s).b[BSTRIDE * HMC_IDX(k, (s).nb) + F_T])
// packed cell index x | y << 8 | z << 16, and in the top byte the wall latched with the
// bead's cell-crossing time
#define BCELL(s, k) (reinterpret_cast<unsigned *>((s).b + BSTRIDE * HMC_IDX(k, (s).nb) + F_CELL)[0])
#define BWALLB(s, k) (reinterpret_cast<uint8_t *>((s).b + BSTRIDE * HMC_IDX(k, (s).nb) + F_CELL)[3])
// one-hot image of the cell, 10 bits per axis (ncell <= 10): 1 << x | 1 << (10 + y) | 1 << (20 + z)
#define BOWN(s, k) (reinterpret_cast<unsigned *>((s).b + BSTRIDE * HMC_IDX(k, (s).nb) + F_CELL)[1])

__device__ __forceinline__ unsigned own_mask(unsigned cx, unsigned cy, unsigned cz) {
  return (1u << cx) | (1u << (10u + cy)) | (1u << (20u + cz));
}
// cells within one step (periodic) of the given cell, per axis, in the layout of own_mask;
// a bead k sits in the 27-neighbourhood iff popc(nbr & own_k) == 3
__device__ __forceinline__ unsigned nbr_mask(const KArgs &a, unsigned cx, unsigned cy, unsigned cz) {
  return unsigned(a.nbr_tab[cx]) | (unsigned(a.nbr_tab[cy]) << 10) | (unsigned(a.nbr_tab[cz]) << 20);
}

// per-replica register state (uniform across the lanes of a group).  Only what the event
// loop touches at every event lives here; the bond constants of the transition are read
// through `tr` where a bond pair is involved, counters of rare events go straight to the
// global accumulators.  (The loop-invariant state of a replica used to take ~60 of the
// 128 registers; the compiler paid for it by rematerialising addresses and special
// registers all over the event loop.)
struct Rep {
  int r;                 // replica index
  int trans;             // its transition
  unsigned long long config;
  const TransDev *tr;
  const uint8_t *pinfo;  // pairinfo matrix of this replica's transition
  const unsigned long long *bmask; // bondmask row
  unsigned err;
  unsigned n_coll, n_cell; // events since the last flush (one MD interval)
  int nl;                // length of the nonlocal event list
  unsigned flags;        // REP_LIVE | REP_TRACING
};
enum { REP_LIVE = 1u,   // not a clone filling an idle group of the warp: has global side effects
       REP_TRACING = 2u };
__device__ __forceinline__ bool rep_live(const Rep &R) { return (R.flags & REP_LIVE) != 0u; }
// RNG identity: replica index in the whole (possibly multi-rank) ensemble
__device__ __forceinline__ unsigned rep_rng_id(const KArgs &a, const Rep &R) {
  return unsigned(R.trans) * a.rng_total + a.rng_first +
         unsigned(R.r - R.trans * a.replicas_per_trans);
}
__device__ __forceinline__ double *rep_sbias(const KArgs &a, const Rep &R) {
  return a.s_bias + size_t(R.r) * a.nstates;
}

// "does any replica group of this warp satisfy p" for a GROUP-UNIFORM predicate: with one
// group per warp (G == 32) that is p itself and no vote instruction is needed.
__device__ __forceinline__ float2 f2(float x) { return make_float2(x, x); }

template <int G> __device__ __forceinline__ bool warp_any(bool p) {
  if constexpr (G == 32) {
    return p;
  } else {
    return __any_sync(0xffffffffu, p);
  }
}

template <int G> __device__ __forceinline__ unsigned group_mask(unsigned wlane) {
  if constexpr (G == 32) {
    return 0xffffffffu;
  } else {
    return ((1u << G) - 1u) << (wlane & ~unsigned(G - 1));
  }
}


// std::round (half away from zero), exact for every finite x: x - trunc(x) is exact,
// so the comparison with 0.5 sees the true fraction.  5 instructions instead of the
// libdevice sequence.
__device__ __forceinline__ double round_away(double x) {
  const double r = trunc(x);
  const double f = x - r;
  return (fabs(f) >= 0.5) ? r + copysign(1.0, x) : r;
}

// d - round(d/l)*l (Box::mindist, box.h:22-26), general form: used for cell walls,
// where the unwrapped coordinate may be many boxes away from the wall.
__device__ __forceinline__ double mindist1(const KArgs &a, double d) {
  return d - round_away(d * a.inv_l) * a.l;
}

// The same for a difference between two beads of the chain.  When the host has
// checked that the fully stretched chain is shorter than 1.5 boxes (a.short_chain; true
// for every shipped configuration) |d/l| < 1.5, so round() is 0 or +-1, the product is
// +-0 or +-l exactly, and one compare + select + copysign replaces trunc/round/multiply.
// Bit-identical to the general form; the branch is uniform across the launch.
__device__ __forceinline__ double mindist_pair(const KArgs &a, double d) {
  const double t = d * a.inv_l;
  if (a.short_chain) return d - copysign(fabs(t) >= 0.5 ? a.l : 0.0, t);
  return d - round_away(t) * a.l;
}
// the same with the (launch-uniform) choice made at compile time: the event path calls it
// three times per prediction
template <bool SC> __device__ __forceinline__ double mindist_pair_t(const KArgs &a, double d) {
  const double t = d * a.inv_l;
  if (SC) return d - copysign(fabs(t) >= 0.5 ? a.l : 0.0, t);
  return d - round_away(t) * a.l;
}

// Box::mindist, box.h:22-26
__device__ __forceinline__ void mindist3(const KArgs &a, double &dx, double &dy, double &dz) {
  dx = mindist_pair(a, dx);
  dy = mindist_pair(a, dy);
  dz = mindist_pair(a, dz);
}
template <bool SC>
__device__ __forceinline__ void mindist3_t(const KArgs &a, double &dx, double &dy, double &dz) {
  dx = mindist_pair_t<SC>(a, dx);
  dy = mindist_pair_t<SC>(a, dy);
  dz = mindist_pair_t<SC>(a, dz);
}

struct Bead {
  double x, y, z, vx, vy, vz, t;
};

__device__ __forceinline__ Bead load_bead(const Smem &s, int k) {
  const double2 *r = reinterpret_cast<const double2 *>(s.b + BSTRIDE * k);
  const double2 a0 = r[0], a1 = r[1], a2 = r[2];
  Bead b;
  b.x = a0.x;
  b.y = a0.y;
  b.z = a1.x;
  b.vx = a1.y;
  b.vy = a2.x;
  b.vz = a2.y;
  b.t = BT(s, k);
  return b;
}

// x y z vx vy vz t of one bead record as three 128-bit + one 64-bit store, all under one
// predicate (st.shared with a guard predicate instead of a branch around the stores)
__device__ __forceinline__ void store_bead(double *rec, const Bead &b, double t, bool p) {
  const unsigned addr = unsigned(__cvta_generic_to_shared(rec));
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %0, 0;\n\t"
      "@p st.shared.v2.f64 [%1], {%2, %3};\n\t"
      "@p st.shared.v2.f64 [%1+16], {%4, %5};\n\t"
      "@p st.shared.v2.f64 [%1+32], {%6, %7};\n\t"
      "@p st.shared.f64 [%1+48], %8;\n\t}"
      :
      : "r"(int(p)), "r"(addr), "d"(b.x), "d"(b.y), "d"(b.z), "d"(b.vx), "d"(b.vy), "d"(b.vz),
        "d"(t)
      : "memory");
}
__device__ __forceinline__ void sts_pred_u32(unsigned *ptr, unsigned v, bool p) {
  const unsigned addr = unsigned(__cvta_generic_to_shared(ptr));
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %0, 0;\n\t@p st.shared.u32 [%1], %2;\n\t}"
               :
               : "r"(int(p)), "r"(addr), "r"(v)
               : "memory");
}
__device__ __forceinline__ void sts_pred_v4(uint4 *ptr, const uint4 &v, bool p) {
  const unsigned addr = unsigned(__cvta_generic_to_shared(ptr));
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %0, 0;\n\t@p st.shared.v4.u32 [%1], {%2, %3, %4, %5};\n\t}"
      :
      : "r"(int(p)), "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
      : "memory");
}
__device__ __forceinline__ void sts_pred_f64(double *ptr, double v, bool p) {
  const unsigned addr = unsigned(__cvta_generic_to_shared(ptr));
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %0, 0;\n\t@p st.shared.f64 [%1], %2;\n\t}"
               :
               : "r"(int(p)), "r"(addr), "d"(v)
               : "memory");
}
__device__ __forceinline__ void sts_pred_u8(uint8_t *ptr, unsigned v, bool p) {
  const unsigned addr = unsigned(__cvta_generic_to_shared(ptr));
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %0, 0;\n\t@p st.shared.u8 [%1], %2;\n\t}"
               :
               : "r"(int(p)), "r"(addr), "r"(v)
               : "memory");
}

// lane bookkeeping of a replica group inside its warp
template <int G> struct GroupLanes {
  unsigned shift; // first lane of the group
  unsigned bits;  // G ones
  unsigned below; // ones below this lane (group-relative)
  __device__ __forceinline__ GroupLanes(int lane, unsigned wlane) { // wlane: lane in the warp
    shift = wlane & ~unsigned(G - 1);
    bits = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    below = (1u << lane) - 1u;
  }
  // the group's slice of a full-warp ballot
  __device__ __forceinline__ unsigned ballot(bool p) const {
    return (__ballot_sync(0xffffffffu, p) >> shift) & bits;
  }
};

// largest value of a group-uniform quantity over the groups of the warp (trip counts)
template <int G> __device__ __forceinline__ int warp_max(int v) {
  if constexpr (G == 32) {
    return v;
  } else {
    return int(__reduce_max_sync(0xffffffffu, unsigned(v)));
  }
}

// delta_tpv (hardspheres.cc:372-392) followed by the type logic of if_nearest_bond /
// if_nnearest_bond (:863-923) or if_coll (:425-476).  Returns the latched event type,
// or 0xff when the reference queues nothing.
//
// Every pair needs exactly ONE root of |dr + dv tau|^2 = sigma^2.  Which one (inner
// root at the lower radius, or outer root at the upper radius) is decided with the
// reference's comparisons only (rv < 0, rv^2 > vv (rr - sigma^2)); then one sqrt and
// one division are evaluated with the operand values the reference's chosen branch
// uses, so the result is bit-identical.  The body is written with selects only, so
// that all lanes of a warp (and both replica groups in it) stay converged through
// the sqrt / division sequences.
//
// Beads are passed as indices (p, q) with p_is_lo telling which one is lower: the
// reference starts from (i,j) = (lo,hi) and swaps iff times[lo] < times[hi], so "i" is
// the bead with the later clock and, on a tie, the lower index.
template <bool SC>
__device__ __forceinline__ unsigned predict_pair(const KArgs &a, const Rep &R, const Smem &s,
                                                 int p, int q, bool p_is_lo, int d,
                                                 unsigned info, bool want, double &tout) {
  // choose the roles by INDEX and load each bead once (no 64-bit selects)
  const double tp = BT(s, p), tq = BT(s, q);
  const bool sw = p_is_lo ? (tp < tq) : !(tq < tp); // true: "i" is q
  const int bi_ = sw ? q : p, bj_ = sw ? p : q;
  const double ti = sw ? tq : tp, tj = sw ? tp : tq;
  // three 128-bit loads per bead
  const double2 *ri = reinterpret_cast<const double2 *>(s.b + BSTRIDE * bi_);
  const double2 *rj = reinterpret_cast<const double2 *>(s.b + BSTRIDE * bj_);
  const double2 i0 = ri[0], i1 = ri[1], i2 = ri[2];
  const double2 j0 = rj[0], j1 = rj[1], j2 = rj[2];
  const double xi = i0.x, yi = i0.y, zi = i1.x, vxi = i1.y, vyi = i2.x, vzi = i2.y;
  const double xj = j0.x, yj = j0.y, zj = j1.x, vxj = j1.y, vyj = j2.x, vzj = j2.y;
  const double dt = ti - tj;
  double dx = xj + vxj * dt - xi;
  double dy = yj + vyj * dt - yi;
  double dz = zj + vzj * dt - zi;
  mindist3_t<SC>(a, dx, dy, dz);
  const double dvx = vxj - vxi, dvy = vyj - vyi, dvz = vzj - vzi;
  const double rv = dx * dvx + dy * dvy + dz * dvz;
  const double vv = dvx * dvx + dvy * dvy + dvz * dvz;
  const double rr = dx * dx + dy * dy + dz * dz;
  const double rv2 = rv * rv;
  const bool approaching = rv < 0;
  const bool local = d <= 2, near = d == 1;
  bool inner, outer_ok;
  double s2;
  unsigned ty;
  // (warp-uniform branch: every lane of the warp calls predict_pair together)
  if (!__any_sync(0xffffffffu, info != 0u)) {
    // no bonded / bondable pair among the 32 pairs of this pass — the common case:
    // local wells and plain hard cores only
    const int cls = min(d, 3) - 1; // 0 nearest, 1 next-nearest, 2 nonlocal
    const double s_in = a.s_in_tab[cls];
    inner = approaching && (rv2 > vv * (rr - s_in)); // t_until_inner_coll :217-240
    s2 = inner ? s_in : a.s_out_tab[cls];
    outer_ok = local;
    ty = unsigned(2 * cls) + ((inner || !local) ? 0u : 1u);
  } else {
    // nonlocal bond bookkeeping (info == 0 for local pairs)
    const unsigned tb = info & 0x7fu;
    const bool perm = (info & 0x80u) != 0;
    const unsigned long long tmask = tb ? (1ull << (tb - 1)) : 0ull;
    const bool bonded_t = (R.config & tmask) != 0;
    const bool nonbonded_t = (~R.config & tmask) != 0;
    const TransDev &T = *R.tr;
    // get_rc2_inner / get_rc2_outer, hardspheres.cc:394-419
    const double r2i = (perm && T.has_p_rc) ? T.p_rc2 : T.rc2;
    double r2o = (perm && T.has_p_rc) ? T.p_rc2 : T.rc2;
    r2o = (T.has_stair && nonbonded_t) ? T.stair2 : r2o;
    const bool capcand = tmask && (unsigned(__popcll(R.config)) < T.max_nbonds) && nonbonded_t;
    // first inner candidate: well minimum (local) / bond radius (capture) / hard core
    const double s_in1 = local ? (near ? a.near_min2 : a.nnear_min2) : (capcand ? r2i : a.rh2);
    const bool hit1 = approaching && (rv2 > vv * (rr - s_in1)); // t_until_inner_coll :217-240
    // a failed capture falls through to the hard core (if_coll :457-467)
    const bool hit2 = !local && capcand && !hit1 && approaching && (rv2 > vv * (rr - a.rh2));
    inner = hit1 || hit2;
    const double s_out = local ? (near ? a.near_max2 : a.nnear_max2) : r2o;
    s2 = hit1 ? s_in1 : (hit2 ? a.rh2 : s_out);
    outer_ok = local || bonded_t || perm || (T.has_stair && nonbonded_t);
    if (local) ty = (near ? HMC_EV_MIN_NEAREST : HMC_EV_MIN_NNEAREST) + (inner ? 0u : 1u);
    else ty = hit1 ? (capcand ? HMC_EV_MAX_NONLOCAL_INNER : HMC_EV_MIN_NONLOCAL_INNER)
                   : (hit2 ? HMC_EV_MIN_NONLOCAL_INNER : HMC_EV_MAX_NONLOCAL_OUTER);
  }
  // Lanes whose result is not needed (pair not predicted by the reference, or no event)
  // get harmless operands: a negative radicand or a NaN/zero dividend would send the
  // whole warp through the slow-path subroutines of sqrt / division.
  const bool valid = want && (inner || outer_ok);
  const double rad1 = vv * (rr - s2);
  // (the harmless operand is a kernel argument, not a literal: with a literal the compiler
  // turns sqrt(valid ? x : 1.0) into valid ? sqrt(x) : 1.0 and the unused lanes' zero or
  // negative radicands take the slow path after all)
  const double root = sqrt(valid ? rv2 - rad1 : a.one);
  const double num = inner ? (-rv - root) : (-rv + root); // :237 / :253
  tout = ti + (valid ? num : 1.0) / (valid ? vv : 1.0);
  return valid ? ty : 0xffu;
}

// t_until_cell + if_cell, hardspheres.cc:560-693.  Returns false when the bead is
// at rest (no event queued).
__device__ __forceinline__ bool predict_cell(const KArgs &a, const Bead &b, int cx, int cy, int cz,
                                             double &tout, unsigned &wall) {
  // general min-image: the unwrapped coordinate may be many boxes from the wall
  const double dx = mindist1(a, a.lcell * double(cx) - b.x);
  const double dy = mindist1(a, a.lcell * double(cy) - b.y);
  const double dz = mindist1(a, a.lcell * double(cz) - b.z);
  bool cr = false;
  double dt = 0.0;
  if (b.vx > 0) {
    dt = (dx + a.lcell) / b.vx;
    cr = true;
    wall = 0;
  } else if (b.vx < 0) {
    dt = dx / b.vx;
    cr = true;
    wall = 1;
  }
  if (b.vy > 0) {
    const double c = (dy + a.lcell) / b.vy;
    if (!cr || c < dt) {
      cr = true;
      dt = c;
      wall = 2;
    }
  } else if (b.vy < 0) {
    const double c = dy / b.vy;
    if (!cr || c < dt) {
      cr = true;
      dt = c;
      wall = 3;
    }
  }
  if (b.vz > 0) {
    const double c = (dz + a.lcell) / b.vz;
    if (!cr || c < dt) {
      cr = true;
      dt = c;
      wall = 4;
    }
  } else if (b.vz < 0) {
    const double c = dz / b.vz;
    if (!cr || c < dt) {
      cr = true;
      dt = c;
      wall = 5;
    }
  }
  tout = b.t + dt;
  return cr;
}

// Cell adjacency within the 27-neighbourhood (periodic; valid for ncell >= 4) for packed
// cells (x | y<<8 | z<<16), all three axes at once with byte-SIMD:
// per-axis |ca - cb| must be 0, 1 or nc-1.  Returns a byte mask (0xff per adjacent axis).
__device__ __forceinline__ unsigned adj_mask(unsigned ca, unsigned cb, unsigned ncm1x3) {
  const unsigned ad = __vabsdiffu4(ca, cb);
  return __vcmpleu4(ad, 0x00010101u) | __vcmpeq4(ad, ncm1x3);
}

// -------------------------------------------------------------------- RNG --

// Philox4x32-10 (Salmon et al., SC'11), counter-based: the stream of a replica
// does not depend on scheduling.
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ double u53(unsigned hi, unsigned lo) {
  const unsigned long long v = ((unsigned long long)hi << 32 | lo) >> 11;
  return (double(v) + 0.5) * (1.0 / 9007199254740992.0); // (0,1)
}

// libstdc++ generate_canonical<double,53> on a 32-bit engine (bits/random.tcc:3349)
__device__ __forceinline__ double canonical_words(unsigned w0, unsigned w1) {
  double sum = double(w0) + double(w1) * 4294967296.0;
  double r = sum / 18446744073709551616.0;
  if (r >= 1.0) r = 0.99999999999999988898;
  return r;
}

struct WordStream { // crankshaft word source: injected tape or Philox
  const uint32_t *tape;
  unsigned n, c;
  bool underrun;
  bool philox;
  uint2 key;
  unsigned c1, c2, c3; // counter words 1..3
  uint4 buf;
  unsigned buf_blk;
  unsigned moves_done;
};

__device__ __forceinline__ unsigned ws_next(WordStream &w) {
  if (w.philox) {
    const unsigned blk = w.c >> 2;
    if (blk != w.buf_blk) {
      w.buf = philox4x32(make_uint4(blk, w.c1, w.c2, w.c3), w.key);
      w.buf_blk = blk;
    }
    const unsigned q = w.c & 3u;
    w.c++;
    return q == 0 ? w.buf.x : q == 1 ? w.buf.y : q == 2 ? w.buf.z : w.buf.w;
  }
  if (w.c >= w.n) {
    w.underrun = true;
    w.c++;
    return 0u;
  }
  return w.tape[w.c++];
}

// libstdc++ uniform_int_distribution on a 32-bit engine: Lemire's method
// (bits/uniform_int_dist.h:255-281)
__device__ __forceinline__ unsigned ws_uniform_int(WordStream &w, unsigned lo, unsigned hi) {
  const unsigned range = hi - lo + 1u;
  unsigned long long prod = (unsigned long long)ws_next(w) * range;
  unsigned low = unsigned(prod);
  if (low < range) {
    const unsigned thr = (0u - range) % range;
    while (low < thr && !w.underrun) {
      prod = (unsigned long long)ws_next(w) * range;
      low = unsigned(prod);
    }
  }
  return unsigned(prod >> 32) + lo;
}

// ------------------------------------------------------------ trace / misc --

__device__ __forceinline__ void trace_event(const KArgs &a, const Rep &R, int lane, unsigned type,
                                            unsigned i, unsigned j, double t) {
  if (R.flags == (REP_LIVE | REP_TRACING) && lane == 0) {
    const unsigned n = *a.trace_n;
    if (n < a.trace_cap) {
      hmc_event e;
      e.t = t;
      e.type = type;
      e.i = i;
      e.j = j;
      e.pad = 0;
      a.trace[n] = e;
    }
    *a.trace_n = n + 1;
  }
}

// total_k_E + compute_hamiltonian, hardspheres.cc:1063-1082 (serial bead order)
__device__ __forceinline__ double hamiltonian(const KArgs &a, const Rep &R, const Smem &s) {
  double ke = 0.0;
  const double hm = 0.5 * a.m;
  for (int i = 0; i < a.nb; i++) {
    const double vx2 = BVX(s, i) * BVX(s, i);
    const double vy2 = BVY(s, i) * BVY(s, i);
    const double vz2 = BVZ(s, i) * BVZ(s, i);
    ke += hm * (vx2 + vy2 + vz2);
  }
  return ke + rep_sbias(a, R)[R.config];
}

// ------------------------------------------------------------ event table --

// what an out-of-line step piece changes in the replica registers
struct StepOut {
  unsigned err;
  int nl;
};

// Cell crossing: keep the earlier of the queued and the new event of a pair
// (add_events_for_bead_after_crossing, hardspheres.cc:740-856: the older event stays
// queued, so the earlier of the two fires).  Rare — a crossing predicts ~4 partners and one
// in ten has an event — and therefore out of line.  Executed by the whole warp.
template <int G>
__device__ __noinline__ StepOut list_merge(const KArgs &a, const Smem &s_in, int nl, unsigned err,
                                           bool mrg, double tt, unsigned key, int lane,
                                           unsigned wlane) {
  const Smem s = s_in;
  const GroupLanes<G> gl(lane, wlane);
  unsigned pend = gl.ballot(mrg);
  while (__any_sync(0xffffffffu, pend != 0u)) {
    const bool act = pend != 0u;
    const int src = int(gl.shift) + (act ? __ffs(int(pend)) - 1 : 0);
    const double ct = __shfl_sync(0xffffffffu, tt, src);
    const unsigned ck = __shfl_sync(0xffffffffu, key, src);
    bool found = false;
    const int nmax = warp_max<G>(nl);
    for (int base = 0; base < nmax; base += G) {
      const int e = base + lane;
      const uint4 v = s.le[HMC_IDX(e, a.list_cap + 32)]; // (whole lane strides: reads past nl)
      const bool m = act && e < nl && ((v.z ^ ck) & 0xffffff00u) == 0u;
      if (m && ct < le_time(v)) s.le[HMC_IDX(e, a.list_cap)] = le_make(ct, ck);
      found |= m;
    }
    const unsigned fb = gl.ballot(found);
    if (act && fb == 0u) {
      if (nl < a.list_cap) {
        if (lane == 0) s.le[HMC_IDX(nl, a.list_cap)] = le_make(ct, ck);
        nl++;
      } else {
        err |= HMC_ERR_LIST_OVERFLOW;
      }
    }
    pend &= pend - 1u;
    __syncwarp();
  }
  StepOut out;
  out.err = err;
  out.nl = nl;
  return out;
}

// Store one prediction the reference made for the pair (b,k) (`on`; pty == 0xff: the
// reference queues nothing).  Local wells go to their dense slot; a nonlocal event is
// appended to the list after a collision (every older entry of the two beads is dropped
// by the sweep of next_event that ends the event) or min-merged into it at a cell crossing.  Executed by the
// whole warp.
template <int G>
__device__ __forceinline__ void commit_prediction(const KArgs &a, Rep &R, const Smem &s, bool on,
                                                  int b, int k, unsigned pty, double tt,
                                                  bool is_cell, int lane,
                                                  const GroupLanes<G> &gl) {
  const int lo = min(b, k), hi = max(b, k), d = hi - lo;
  const bool have = on && pty != 0xffu;
  // local wells: collisions only (a crossing predicts nonlocal partners only), always an event
  double *dense = d == 1 ? s.b + BSTRIDE * lo + F_TN : s.tnn + lo;
  if (have && d < 3) *dense = tt;
  const bool nl = have && d >= 3;
  const unsigned key = PAIR_KEY | (unsigned(lo) << 16) | (unsigned(hi) << 8) | pty;
  // collision: append behind the surviving entries
  const bool app = nl && !is_cell;
  const unsigned gb = gl.ballot(app);
  const int dst = R.nl + __popc(gb & gl.below);
  const int nnew = R.nl + __popc(gb);
  if (app && dst < a.list_cap) s.le[HMC_IDX(dst, a.list_cap)] = le_make(tt, key);
  if (nnew > a.list_cap) R.err |= HMC_ERR_LIST_OVERFLOW;
  R.nl = min(nnew, a.list_cap);
  const bool mrg = nl && is_cell;
  if (__any_sync(0xffffffffu, mrg)) {
    const Smem sc = s; // (a copy: the address of s itself must not escape)
    const StepOut mo = list_merge<G>(a, sc, R.nl, R.err, mrg, tt, key, lane, gl.shift + unsigned(lane));
    R.nl = mo.nl;
    R.err = mo.err;
  }
}

// --------------------------------------------------- step initialisation --

// check_bond, hardspheres.cc:980-1004
__device__ __forceinline__ bool check_bond_tol(double min2, double max2, double d2) {
  const double tol = 1000000 * 2.220446049250313e-16;
  return d2 > min2 * (1 - tol) && d2 < max2 * (1 + tol);
}

// min-image squared distance of beads i,j with the operand order pos[i]-pos[j]
// P points at the first component of a coordinate triple inside the bead records
// (positions: s.b + F_X, crankshaft trial coordinates: s.b + F_VX), stride BSTRIDE.
__device__ __forceinline__ double dist2_ij(const KArgs &a, const double *P, int i, int j) {
  const double *pi = P + BSTRIDE * i, *pj = P + BSTRIDE * j;
  double dx = pi[0] - pj[0], dy = pi[1] - pj[1], dz = pi[2] - pj[2];
  mindist3(a, dx, dy, dz);
  return dx * dx + dy * dy + dz * dz;
}

// check_local_dist (hardspheres.cc:1006-1037, tolerance) / check_local_dist_if_crankshaft
// (crankshaft.cc:88-123, strict) and check_nonlocal_dist (crankshaft.cc:125-171) over all
// pairs, lane-parallel.  Returns bit0 = local ok, bit1 = nonlocal ok (group-uniform).
template <int G>
__device__ __forceinline__ unsigned check_constraints(const KArgs &a, const Rep &R, const double *P,
                                                      bool strict, int lane, unsigned gmask) {
  bool lok = true, nok = true;
  // bond constants of the transition, read once (not per pair)
  const double perm_r2 = R.tr->has_p_rc ? R.tr->p_rc2 : R.tr->rc2;
  const bool has_stair = R.tr->has_stair != 0u;
  const double stair2 = R.tr->stair2;
  for (int p = lane; p < a.npairs; p += G) {
    const uchar2 ij = __ldg(&a.pair_ij[p]);
    const int i = ij.x, j = ij.y, d = j - i;
    if (d <= 2) {
      // reference operand order here is pos[j]-pos[i]; dist2 is sign-symmetric
      // only after squaring, so keep the order
      const double *pi = P + BSTRIDE * i, *pj = P + BSTRIDE * j;
      double dx = pj[0] - pi[0], dy = pj[1] - pi[1], dz = pj[2] - pi[2];
      mindist3(a, dx, dy, dz);
      const double d2 = dx * dx + dy * dy + dz * dz;
      const double mn = d == 1 ? a.near_min2 : a.nnear_min2;
      const double mx = d == 1 ? a.near_max2 : a.nnear_max2;
      const bool ok = strict ? (d2 > mn && d2 < mx) : check_bond_tol(mn, mx, d2);
      lok = lok && ok;
    } else {
      const double d2 = dist2_ij(a, P, i, j);
      const unsigned info = __ldg(&R.pinfo[i * PINFO_STRIDE + j]);
      bool ok = d2 > a.rh2;
      if (info & 0x80u) ok = ok && (d2 < perm_r2);
      if (has_stair && (info & 0x7fu)) ok = ok && (d2 < stair2);
      nok = nok && ok;
    }
  }
  const unsigned lb = __ballot_sync(0xffffffffu, !lok) & gmask;
  const unsigned nbad = __ballot_sync(0xffffffffu, !nok) & gmask;
  return (lb ? 0u : 1u) | (nbad ? 0u : 2u);
}

// initialize_system, main.cc:45-95: constraint checks, init_cells, velocities
// (+ COM shift), and the full prediction fill of the event table.
// Executed by the whole warp; only groups with `on` change their state (the others are in
// the middle of a step).  Once per MD interval: kept out of line, the event loop's
// instruction footprint is what the instruction cache has to hold.
// Out-of-line pieces take the replica registers and the shared-memory pointers by const
// reference (the caller's copy) and work on local copies; what they change comes back by
// value — a reference into the caller's frame would turn every access into a local-memory
// load.
template <int G, bool SC>
__device__ HMC_COLD StepOut init_step(const KArgs &a, const Rep &R_in, const Smem &s_in,
                                          const double *normals, unsigned long long draw, int lane,
                                          unsigned wlane, unsigned gmask, bool on) {
  Rep R = R_in;
  const Smem s = s_in;
  const int nb = a.nb;
  const unsigned ok = check_constraints<G>(a, R, s.b + F_X, false, lane, gmask);
  if (on && !(ok & 1u)) R.err |= HMC_ERR_LOCAL_OVERLAP;
  if (on && !(ok & 2u)) R.err |= HMC_ERR_NONLOCAL_OVERLAP;

  // init_cells (hardspheres.cc:334-369) + raw velocity draws (init_vel :197-209)
  const unsigned rng_id = rep_rng_id(a, R);
  for (int k = lane; k < nb && on; k += G) {
    double x = BX(s, k), y = BY(s, k), z = BZ(s, k);
    x -= floor(x * a.inv_l) * a.l; // Box::minpos
    y -= floor(y * a.inv_l) * a.l;
    z -= floor(z * a.inv_l) * a.l;
    int ix = int(a.inv_lcell * x), iy = int(a.inv_lcell * y), iz = int(a.inv_lcell * z);
    const int hi = a.ncell - 1;
    ix = min(max(ix, 0), hi);
    iy = min(max(iy, 0), hi);
    iz = min(max(iz, 0), hi);
    BCELL(s, k) = unsigned(ix) | (unsigned(iy) << 8) | (unsigned(iz) << 16);
    BOWN(s, k) = own_mask(min(unsigned(ix), 9u), min(unsigned(iy), 9u), min(unsigned(iz), 9u)); // (used iff ncell <= 10)
    double nx, ny, nz;
    if (a.rng_mode == HMC_RNG_INJECTED) {
      nx = normals[3 * k + 0];
      ny = normals[3 * k + 1];
      nz = normals[3 * k + 2];
    } else {
      // Maxwell-Boltzmann by Box-Muller from two Philox blocks per bead;
      // counter = (2k+c, draw_lo, draw_hi, replica)
      const uint2 key = make_uint2(a.seed0, a.seed1);
      const uint4 c0 = philox4x32(
          make_uint4(2u * k, unsigned(draw), unsigned(draw >> 32), rng_id), key);
      const uint4 c1 = philox4x32(
          make_uint4(2u * k + 1u, unsigned(draw), unsigned(draw >> 32), rng_id), key);
      const double r0 = sqrt(-2.0 * log(u53(c0.x, c0.y)));
      double sn, cs;
      sincospi(2.0 * u53(c0.z, c0.w), &sn, &cs);
      const double r1 = sqrt(-2.0 * log(u53(c1.x, c1.y)));
      double sn1, cs1;
      sincospi(2.0 * u53(c1.z, c1.w), &sn1, &cs1);
      nx = r0 * cs * a.vsigma;
      ny = r0 * sn * a.vsigma;
      nz = r1 * cs1 * a.vsigma;
    }
    BVX(s, k) = nx;
    BVY(s, k) = ny;
    BVZ(s, k) = nz;
  }
  __syncwarp();

  // shift_vel_CM_to_0 (hardspheres.cc:173-194): serial sums in bead order
  double tx = 0.0, ty = 0.0, tz = 0.0;
  for (int i = 0; i < nb; i++) {
    tx += BVX(s, i);
    ty += BVY(s, i);
    tz += BVZ(s, i);
  }
  const double cmx = tx / double(nb), cmy = ty / double(nb), cmz = tz / double(nb);
  __syncwarp();
  for (int k = lane; k < nb && on; k += G) {
    BVX(s, k) -= cmx;
    BVY(s, k) -= cmy;
    BVZ(s, k) -= cmz;
  }
  __syncwarp();

  // empty event list
  R.nl = on ? 0 : R.nl;
  // init_cell_events (hardspheres.cc:696-704)
  for (int k = lane; k < nb && on; k += G) {
    const Bead b = load_bead(s, k);
    double t;
    unsigned wall = 0;
    const unsigned ck = BCELL(s, k);
    const bool cr = predict_cell(a, b, int(ck & 0xffu), int((ck >> 8) & 0xffu), int((ck >> 16) & 0xffu), t, wall);
    s.b[BSTRIDE * k + F_TC] = cr ? t : CUDART_INF;
    BWALLB(s, k) = uint8_t(wall);
    // the last beads have no (k,k+1) / (k,k+2) partner
    if (k >= nb - 1) s.b[BSTRIDE * k + F_TN] = CUDART_INF;
    if (k >= nb - 2) s.tnn[HMC_IDX(k, nb)] = CUDART_INF;
  }
  __syncwarp();
  // init_nearest/nnearest_bond_events (:885-937) and add_events_for_all_beads (:540-555)
  const GroupLanes<G> gl(lane, wlane);
  for (int base = 0; base < a.npairs; base += G) {
    const int p = min(base + lane, a.npairs - 1); // every lane calls predict_pair
    const bool mine = base + lane < a.npairs;
    const uchar2 ij = __ldg(&a.pair_ij[p]);
    const int i = ij.x, j = ij.y, d = j - i;
    bool doit = mine;
    if (d >= 3)
      doit = mine && (adj_mask(BCELL(s, i), BCELL(s, j), unsigned(a.ncell - 1) * 0x00010101u) &
                      0x00ffffffu) == 0x00ffffffu;
    double tt;
    const unsigned ty = predict_pair<SC>(a, R, s, i, j, true, d,
                                         __ldg(&R.pinfo[i * PINFO_STRIDE + j]), doit && on, tt);
    commit_prediction<G>(a, R, s, doit && on, i, j, ty, tt, false, lane, gl);
  }
  __syncwarp();
  StepOut out;
  out.err = R.err;
  out.nl = R.nl;
  return out;
}

// --------------------------------------------------------- the event loop --

// Next event of the replica: lane-strided minimum over the dense slots (cell walls, local
// wells) and the nonlocal list + shuffle reduce.  Ties go to the lowest key, i.e. cell
// events by bead first, then pairs in (i,j) order.  Replaces EventQueue::top()
// (event.h:105-116).
// The same sweep over the list finishes a collision of (bi,bj) (`coll`): that collision
// invalidated every event the two beads had queued (the counter bump,
// hardspheres.cc:1118-1119), i.e. the list entries below n_old that touch bi or bj — the
// entries from n_old on are the predictions just made.  They are dropped and the rest is
// compacted in place, keeping its order.  Executed by the whole warp.
template <int G>
__device__ __forceinline__ void next_event(const Smem &s, Rep &R, bool coll, int bi, int bj,
                                           int n_old, int lane, const GroupLanes<G> &gl,
                                           double &bt, unsigned &bk) {
  const int nb = s.nb;
  double ct = CUDART_INF, pt = CUDART_INF;
  unsigned ck = 0u, pk = 0xffffffffu;
  // dense slots, one bead per lane; group-uniform trip count, tail lanes re-read the last
  // bead (same times, same keys).  Cells and pairs keep separate minima: within each the
  // keys come in ascending order, so a strict '<' keeps the lowest key of a tie.
  for (int base = 0; base < nb; base += G) {
    const int k = min(base + lane, nb - 1);
    const double2 d = *reinterpret_cast<const double2 *>(s.b + BSTRIDE * k + F_TC);
    const double t2 = s.tnn[HMC_IDX(k, s.nb)];
    const unsigned kb = unsigned(k) * 0x00010100u; // k << 16 | k << 8
    const bool l0 = d.x < ct;
    ct = l0 ? d.x : ct;
    ck = l0 ? unsigned(k) << 16 : ck;
    const bool l1 = d.y < pt;
    pt = l1 ? d.y : pt;
    pk = l1 ? kb + (PAIR_KEY | 0x100u) : pk; // (k, k+1)
    const bool l2 = t2 < pt;
    pt = l2 ? t2 : pt;
    pk = l2 ? kb + (PAIR_KEY | 0x200u) : pk; // (k, k+2)
  }
  // nonlocal list: unordered, so the full tie rule
  const int nl = R.nl;
  const int nmax = warp_max<G>(nl);
  const unsigned pi = unsigned(bi) * 0x00010100u, pj = unsigned(bj) * 0x00010100u;
  int nk = 0;
  for (int base = 0; base < nmax; base += G) {
    const int e = base + lane;
    const uint4 v = s.le[HMC_IDX(e, s.cap + 32)];
    const unsigned x = v.z & 0x00ffff00u; // lo, hi
    const bool touch = ((__vcmpeq4(x, pi) | __vcmpeq4(x, pj)) & 0x00ffff00u) != 0u;
    const bool keep = e < nl && !(coll && e < n_old && touch);
    const unsigned gb = gl.ballot(keep);
    const int dst = nk + __popc(gb & gl.below); // dst <= e: never ahead of the reads
    sts_pred_v4(s.le + HMC_IDX(dst, s.cap + 32), v, keep && coll);    // (without a collision nothing moves)
    nk += __popc(gb);
    const double t = le_time(v);
    const bool lt = keep && ((t < pt) | ((t == pt) & (v.z < pk)));
    pt = lt ? t : pt;
    pk = lt ? v.z : pk;
  }
  R.nl = nk;
  const bool tp = pt < ct; // a cell event wins a tie against any pair
  bt = tp ? pt : ct;
  bk = tp ? pk : ck;
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) {
    const double ot = __shfl_xor_sync(0xffffffffu, bt, off);
    const unsigned ok = __shfl_xor_sync(0xffffffffu, bk, off);
    const bool take = (ot < bt) | ((ot == bt) & (ok < bk)); // (no short-circuit branches)
    bt = take ? ot : bt;
    bk = take ? ok : bk;
  }
}

// One event of run_step's loop (main.cc:107-132) for every group of the warp that has
// one (`has`); the others are predicated off.
//
// The body is one instruction path for collisions and cell crossings alike (selects
// instead of branches): the replica groups that share a warp execute it together, and
// a crossing simply discards the collision arithmetic it did not need.
template <int G, bool SC, bool TP>
__device__ __forceinline__ void event_body(const KArgs &a, Rep &R, const Smem &s, bool has,
                                           double &bt_io, unsigned &bk_io, int lane,
                                           const GroupLanes<G> &gl) {
  const int nb = a.nb, nc = a.ncell;
  {
    const double bt = bt_io;
    const unsigned bk = has ? bk_io : 0u;
    const int n_old = R.nl; // list entries that predate this event

    const double t = bt;
    const bool is_cell = bk < PAIR_KEY;
    const int nbead = is_cell ? 1 : 2; // beads whose slots are re-predicted
    int bi = int((bk >> 16) & 0xffu);
    int bj = is_cell ? bi : int((bk >> 8) & 0xffu);
    // (opaque to the compiler: it would otherwise re-derive the two indices from the key at
    // every use in the partner loops instead of keeping them in registers)
    asm volatile("" : "+r"(bi), "+r"(bj));
    const int pidx = bi * PINFO_STRIDE + bj;
    const int dij = bj - bi;
    // the wall latched at the prediction of a crossing (t_until_cell :589-666)
    const unsigned wall = is_cell ? unsigned(BWALLB(s, bi)) : 0u;

    Bead I = load_bead(s, bi), J = load_bead(s, bj);
    const unsigned ci0 = BCELL(s, bi);
    int cxn = int(ci0 & 0xffu), cyn = int((ci0 >> 8) & 0xffu), czn = int((ci0 >> 16) & 0xffu);
    // BeadCellEvent (hardspheres.cc:1622-1645): new cell from the wall latched at
    // prediction (t_until_cell :589-666)
    if (is_cell) {
      // (c +- 1) mod nc by compare/select: an integer '%' is a ~20-instruction sequence
      const int step = (wall & 1u) ? -1 : 1;
      int moved = (wall < 2 ? cxn : wall < 4 ? cyn : czn) + step;
      moved = moved < 0 ? nc - 1 : (moved >= nc ? 0 : moved);
      cxn = wall < 2 ? moved : cxn;
      cyn = (wall >= 2 && wall < 4) ? moved : cyn;
      czn = (wall >= 4 && wall < 6) ? moved : czn;
    }
    // update_pos (:1040-1047) for both beads (the same bead twice for a crossing)
    const double dti = t - I.t, dtj = t - J.t;
    I.x += I.vx * dti;
    I.y += I.vy * dti;
    I.z += I.vz * dti;
    J.x += J.vx * dtj;
    J.y += J.vy * dtj;
    J.z += J.vz * dtj;
    // collision events, hardspheres.cc:1106-1620: final_vel (:1050-1061)
    double dx = J.x - I.x, dy = J.y - I.y, dz = J.z - I.z;
    mindist3_t<SC>(a, dx, dy, dz);
    // latched event type: in the key for a nonlocal pair; a local-well event happens ON the
    // inner or the outer wall of the well (radii 15 % apart), so the separation at the
    // event tells which one was latched; for a crossing: the wall
    const int cls = min(dij, 2) - 1;
    const bool on_inner = dx * dx + dy * dy + dz * dz < a.s_mid_tab[max(cls, 0)];
    const unsigned ty = is_cell ? wall
                                : (dij >= 3 ? (bk & 0xffu) : unsigned(2 * cls) + (on_inner ? 0u : 1u));
    // radius of the wall that was hit
    double sigma2 = a.sigma2_tab[min(ty, 4u)];
    // bond events (capture at the bond radius / the outer wall of a bond, a staircase or a
    // permanent bond) are rare: their constants are read from the transition's table inside
    // a warp-uniform branch
    const bool bondev = has && !is_cell && ty >= HMC_EV_MAX_NONLOCAL_INNER;
    unsigned long long tmask = 0ull;
    bool capture = false, has_dS = false;
    double dS = 0.0;
    if (__any_sync(0xffffffffu, bondev)) {
      const unsigned info = bondev ? unsigned(__ldg(&R.pinfo[pidx])) : 0u;
      const unsigned tb = info & 0x7fu;
      const bool perm = (info & 0x80u) != 0;
      tmask = tb ? (1ull << (tb - 1)) : 0ull;
      const TransDev &T = *R.tr;
      const double r2p = (perm && T.has_p_rc) ? T.p_rc2 : T.rc2;
      double s2b = r2p;                                   // MaxNonlocalInner :1462-1466
      if (ty == HMC_EV_MAX_NONLOCAL_OUTER)                // :1548-1552
        s2b = (T.has_stair && (~R.config & tmask)) ? T.stair2 : r2p;
      sigma2 = bondev ? s2b : sigma2;
      // entropy step of bond events: s_of_inner_event :299-313 / s_of_outer_event :317-331
      capture = bondev && ty == HMC_EV_MAX_NONLOCAL_INNER;
      const bool escape = bondev && ty == HMC_EV_MAX_NONLOCAL_OUTER && tmask && (R.config & tmask);
      has_dS = capture || escape;
      if (has_dS) {
        const double *sb = rep_sbias(a, R);
        dS = sb[unsigned(R.config ^ tmask)] - sb[unsigned(R.config)];
      }
    }
    sigma2 = (is_cell || !has) ? 1.0 : sigma2;
    // v_after_coll (:258-295)
    const double dvx = J.vx - I.vx, dvy = J.vy - I.vy, dvz = J.vz - I.vz;
    const double mu = 0.5 * a.m;
    // (a crossing or an idle group divides 1 by 1: a zero dividend takes the slow path)
    const bool coll = has && !is_cell;
    double rv = (coll ? dx * dvx + dy * dvy + dz * dvz : 1.0) / sigma2;
    const double rv2 = rv * rv;
    bool transmitted = false;
    if (has_dS) {
      const double dSn = 2 * dS / (sigma2 * mu);
      if (rv2 > dSn) {
        rv -= copysign(sqrt(rv2 - dSn), rv);
        transmitted = true;
      }
    }
    rv = transmitted ? rv : rv + rv;
    const double px = rv * dx, py = rv * dy, pz = rv * dz;
    if (!is_cell) {
      I.vx += px / 2;
      I.vy += py / 2;
      I.vz += pz / 2;
      J.vx -= px / 2;
      J.vy -= py / 2;
      J.vz -= pz / 2;
    }
    if (transmitted && has) { // (only inside a bond event: rare)
      R.config ^= tmask;
      // count_bond.formed / broken (:1470-1474, :1556-1560)
      if (lane == 0 && rep_live(R))
        atomicAdd(&a.acc[size_t(a.n_trans) * a.nstates + (capture ? 0 : a.n_trans) + R.trans], 1ull);
    }
    R.n_cell += (has && is_cell) ? 1u : 0u;
    R.n_coll += (has && !is_cell) ? 1u : 0u;
    __syncwarp();
    // write the 1-2 event beads back: lane 0 stores bead bi (and its cell), lane 1 bead bj,
    // as predicated 128-bit stores (no divergent branch in the lockstep path)
    store_bead(s.b + BSTRIDE * bi, I, t, has && lane == 0);
    if (has && lane == 0) { // (the wall byte is rewritten with the new cell time below)
      BCELL(s, bi) = unsigned(cxn) | (unsigned(cyn) << 8) | (unsigned(czn) << 16);
      BOWN(s, bi) = own_mask(min(unsigned(cxn), 9u), min(unsigned(cyn), 9u), min(unsigned(czn), 9u));
    }
    store_bead(s.b + BSTRIDE * bj, J, t, has && lane == 1 && !is_cell);
    if (has) trace_event(a, R, lane, is_cell ? unsigned(HMC_EV_BEAD_CELL) : ty, bi, is_cell ? wall : bj, t);
    __syncwarp();

    // ---- re-prediction ----
    // collision: the local wells of bi and bj always, nonlocal partners only inside the
    // 27-cell neighbourhood (add_events_for_one_bead :507-536)
    // crossing: only partners in the 9 newly adjacent cells, min-merged
    // (add_events_for_bead_after_crossing :740-856); plus if_cell for each bead.

    // if_cell (:678-693) for the 1 or 2 beads: one lane per (bead, axis) evaluates its
    // wall time (one division each, in parallel), then the reference's x,y,z order with
    // strict '<' is applied to the three candidates.
    {
      const int qb = lane >= 3 ? 1 : 0, ax = lane - 3 * qb; // lanes 0-2: bead bi, 3-5: bead bj
      const bool on = lane < 6;
      const int b = qb ? bj : bi;
      double *rec = s.b + BSTRIDE * b;
      const int axc = on ? ax : 0;
      const double v_ld = rec[F_VX + axc], p_ld = rec[F_X + axc]; // (unconditional loads)
      const unsigned c_ld = BCELL(s, b);
      const double v = on ? v_ld : 1.0;
      const double pp = on ? p_ld : 0.0;
      const double cc = on ? double((c_ld >> (8 * axc)) & 0xffu) : 0.0;
      const double dd = mindist1(a, a.lcell * cc - pp);
      const double numer = v > 0 ? dd + a.lcell : dd;
      const double cand = numer / v;
      // earliest wall of each bead: fold z into y, then that into x, inside the lane
      // triplets (strict '<': on a tie the earlier axis stays, as the reference's x,y,z
      // scan does); an axis at rest competes with +inf
      double ct = (on && v != 0.0) ? cand : CUDART_INF;
      int cw = 2 * axc + (v < 0 ? 1 : 0); // wall code: 2 axis + (negative side)
#pragma unroll
      for (int pass = 1; pass >= 0; pass--) {
        const double ot = __shfl_down_sync(0xffffffffu, ct, 1, G);
        const int ow = __shfl_down_sync(0xffffffffu, cw, 1, G);
        const bool take = (ax == pass) & on & (ot < ct);
        ct = take ? ot : ct;
        cw = take ? ow : cw;
      }
      // lane 0 now holds bead bi's wall, lane 3 bead bj's
      const bool mine = has && (lane == 0 || (lane == 3 && !is_cell));
      if (mine) {
        rec[F_TC] = rec[F_T] + ct;
        reinterpret_cast<uint8_t *>(rec + F_CELL)[3] = uint8_t(cw);
      }
    }

    // partner predictions, both beads in one flattened lane loop
    {
      const unsigned axis = wall >> 1;
      const unsigned cb0 = BCELL(s, bi), cb1 = BCELL(s, bj);
      const unsigned ncm1x3 = unsigned(nc - 1) * 0x00010101u;
      // the cell coordinate that became adjacent through a crossing
      int far = int((cb0 >> (8 * axis)) & 0xffu) + ((wall & 1u) ? -1 : 1);
      far += far < 0 ? nc : 0;
      far -= far >= nc ? nc : 0;
      const unsigned others = 0x00ffffffu & ~(0xffu << (8 * axis)); // the two other axes
      // ncell <= 10: which partners the reference predicts, as one mask per event bead
      // (the 27-neighbourhood of its cell; for a crossing the 9 cells that became adjacent)
      unsigned probe0 = 0u, probe1 = 0u;
      if (a.cell_masks) {
        probe0 = nbr_mask(a, cb0 & 0xffu, (cb0 >> 8) & 0xffu, (cb0 >> 16) & 0xffu);
        probe1 = nbr_mask(a, cb1 & 0xffu, (cb1 >> 8) & 0xffu, (cb1 >> 16) & 0xffu);
        if (is_cell) probe0 = (probe0 & ~(0x3ffu << (10u * axis))) | ((1u << far) << (10u * axis));
      }
      if (!TP) {
        // short chains: a single pass over all 2 nb candidates (the pre-filter below costs
        // more than the few predictions it would save)
        const int nitem = nbead * nb;
        // warp-uniform trip count: 2*nb unless every group of the warp has a crossing
        const int nloop = warp_any<G>(has && !is_cell) ? 2 * nb : nb;
        for (int base = 0; base < nloop; base += G) {
          const int idx = base + lane;
          const bool second = idx >= nb;
          const int b = second ? bj : bi;
          int k = second ? idx - nb : idx;
          const bool in_range = idx < nitem;
          k = in_range ? k : b;
          const int d = abs(b - k);
          const int pq = b * PINFO_STRIDE + k;
          // which (b,k) the reference predicts at this event: local wells at a collision,
          // nonlocal partners in the probed cells
          bool want;
          if (a.cell_masks) {
            want = __popc((second ? probe1 : probe0) & BOWN(s, k)) == 3;
          } else {
            const unsigned cb = second ? cb1 : cb0, ck = BCELL(s, k);
            const unsigned am = adj_mask(cb, ck, ncm1x3);
            want = is_cell ? (int((ck >> (8 * axis)) & 0xffu) == far && (am & others) == others)
                           : (am & 0x00ffffffu) == 0x00ffffffu;
          }
          // (the pair (bi,bj) itself is predicted once, from bi's side)
          const bool active = has && in_range && k != b && !(second && k == bi);
          const bool doit = active && (d < 3 ? !is_cell : want);
          double tt;
          const unsigned pty =
              predict_pair<SC>(a, R, s, b, k, b < k, d, __ldg(&R.pinfo[pq]), doit, tt);
          commit_prediction<G>(a, R, s, doit, b, k, pty, tt, is_cell, lane, gl);
        }
      } else {
        // Pass 1 — which (b,k) does the reference predict, and can the pair collide at all?
        // Most candidates are plain nonlocal hard-core pairs (info == 0, |b-k| >= 3) whose
        // only possible event is an inner collision at rh: approaching and
        // rv^2 > vv (rr - rh^2) (if_coll :425-476).  A single-precision evaluation of the
        // same quantities with rigorous error margins rejects the pairs that certainly fail
        // that test (receding, or missing by more than the margin).  Everything else — local
        // wells, bond pairs, hits and near misses — goes to the exact double-precision
        // prediction of pass 2, compacted so that its lanes are all busy.  A rejected pair
        // has no event in exact arithmetic either, so the table is bit-identical to
        // predicting every candidate.
        // One lane per partner k; it is tested against both event beads, sharing its loads.
        // The two tests of a lane (partner k against bi and against bj) are the same
        // arithmetic on two data sets: they run as packed float2 operations (FFMA2 / FADD2
        // / FMUL2 of sm_100), bi in .x and bj in .y.
        // event-bead velocities (negated) and their 1-norms
        const float2 nux = make_float2(-float(I.vx), -float(J.vx));
        const float2 nuy = make_float2(-float(I.vy), -float(J.vy));
        const float2 nuz = make_float2(-float(I.vz), -float(J.vz));
        const float2 U = make_float2(fabsf(nux.x) + fabsf(nuy.x) + fabsf(nuz.x),
                                     fabsf(nux.y) + fabsf(nuy.y) + fabsf(nuz.y));
        const float2 nox = make_float2(0.0f, -float(J.x - I.x)), noy = make_float2(0.0f, -float(J.y - I.y)),
                     noz = make_float2(0.0f, -float(J.z - I.z));
        // lengths: |x_k(t_k) - x_b(t)| <= stretched chain + |v_k| dt; one more |v_k| dt for the
        // advance and one box for the wrap
        const unsigned long long bm0 = __ldg(&R.bmask[bi]), bm1 = __ldg(&R.bmask[bj]);
        int cnt = 0;
        HMC_PF_UNROLL_PRAGMA
        for (int base = 0; base < nb; base += G) {
          const int kk = base + lane;
          const bool kin = kk < nb;
          const int k = kin ? kk : 0;
          const uint2 ckw = *reinterpret_cast<const uint2 *>(s.b + BSTRIDE * k + F_CELL);
          const unsigned ck = ckw.x, own_k = ckw.y;
          const double2 *rk = reinterpret_cast<const double2 *>(s.b + BSTRIDE * k);
          const double2 k0 = rk[0], k1 = rk[1], k2 = rk[2];
          const float dtf = float(t - BT(s, k));
          const float wx = float(k1.y), wy = float(k2.x), wz = float(k2.y);
          const float W = fabsf(wx) + fabsf(wy) + fabsf(wz);
          // magnitudes that bound every intermediate: X for lengths, V for speeds
          const float X = fmaf(W + W, dtf, a.pf_xc);
          const float2 XV = __fmul2_rn(make_float2(X, X), __fadd2_rn(make_float2(W, W), U));
          const float far_min = 1e-5f * X * X;
          // separation at the partner's clock: subtract in double (unwrapped coordinates
          // may be large), then everything else in single precision
          const float2 dt2 = make_float2(dtf, dtf);
          const float2 wx2 = make_float2(wx, wx), wy2 = make_float2(wy, wy), wz2 = make_float2(wz, wz);
          // (x_k - x_j = (x_k - x_i) - (x_j - x_i): one conversion per axis, the offset of
          // bead bj from bead bi is event-uniform and already in single precision)
          float2 dx = __ffma2_rn(wx2, dt2, __fadd2_rn(f2(float(k0.x - I.x)), nox));
          float2 dy = __ffma2_rn(wy2, dt2, __fadd2_rn(f2(float(k0.y - I.y)), noy));
          float2 dz = __ffma2_rn(wz2, dt2, __fadd2_rn(f2(float(k1.x - I.z)), noz));
          // minimum image, d - l rint(d/l).  (A predicted pair sits in adjacent cells:
          // |separation| <= 2 lcell <= 0.4 boxes per axis with ncell >= 5, far from the
          // half-box where the rounding mode or a single-precision error could matter.)
          dx = __ffma2_rn(__fadd2_rn(__ffma2_rn(dx, a.pf_invl2, a.pf_magic), a.pf_nmagic), a.pf_nl2, dx);
          dy = __ffma2_rn(__fadd2_rn(__ffma2_rn(dy, a.pf_invl2, a.pf_magic), a.pf_nmagic), a.pf_nl2, dy);
          dz = __ffma2_rn(__fadd2_rn(__ffma2_rn(dz, a.pf_invl2, a.pf_magic), a.pf_nmagic), a.pf_nl2, dz);
          const float2 ex = __fadd2_rn(wx2, nux), ey = __fadd2_rn(wy2, nuy), ez = __fadd2_rn(wz2, nuz);
          const float2 rv = __ffma2_rn(dx, ex, __ffma2_rn(dy, ey, __fmul2_rn(dz, ez)));
          const float2 vv = __ffma2_rn(ex, ex, __ffma2_rn(ey, ey, __fmul2_rn(ez, ez)));
          const float2 gap = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __ffma2_rn(dz, dz, a.pf_nrh2)));
          // error bounds with unit roundoff u = 6e-8: |d rv| <= 12 u X V, |d rr| <= 17 u X^2,
          // |d (rv^2 - vv gap)| <= 50 u X^2 V^2; the margins are 14-33 times larger:
          // receding: rv > 1e-5 XV;  misses: rv^2 + 1e-4 (XV)^2 < vv gap
          const float2 rv_min = __fmul2_rn(XV, a.pf_c1e5);
          const float2 xvs = __fmul2_rn(XV, a.pf_c1e2);
          const float2 lhs = __ffma2_rn(xvs, xvs, __fmul2_rn(rv, rv));
          const float2 rhs = __fmul2_rn(vv, gap);
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const int b = q ? bj : bi;
            const int d = abs(b - k);
            // which (b,k) the reference predicts at this event: local wells at a collision,
            // nonlocal partners in the probed cells
            bool want;
            if (a.cell_masks) {
              want = __popc((q ? probe1 : probe0) & own_k) == 3;
            } else {
              const unsigned am = adj_mask(q ? cb1 : cb0, ck, ncm1x3);
              want = is_cell ? (int((ck >> (8 * axis)) & 0xffu) == far && (am & others) == others)
                             : (am & 0x00ffffffu) == 0x00ffffffu;
            }
            // (bj's side: collisions only, and the pair (bi,bj) itself is bi's)
            const bool active = has && kin && k != b && (q == 0 || (!is_cell && k != bi));
            const bool doit = active && (d < 3 ? !is_cell : want);
            const bool plain = (((q ? bm1 : bm0) >> k) & 1ull) == 0ull; // no bond bookkeeping
            const bool far_apart = (q ? gap.y : gap.x) > far_min;
            const bool receding = (q ? rv.y : rv.x) > (q ? rv_min.y : rv_min.x);
            const bool misses = (q ? lhs.y : lhs.x) < (q ? rhs.y : rhs.x);
            // (dtf < 0: the partner's clock is ahead of this event — only reachable after a
            // bond has left its well, a state the reference aborts on at the next step; such
            // pairs go to the exact path like everything else that is unusual)
            const bool reject = d >= 3 && plain && far_apart && (receding || misses) && dtf >= 0.0f;
#ifdef HMC_VERIFY_PREFILTER
            const bool keep = doit;
            const bool flagged = doit && reject; // predicted anyway; pass 2 checks it has no event
#else
            const bool keep = doit && !reject;
#endif
            const unsigned gbits = gl.ballot(keep);
            // (branch-free: a store that is not wanted goes to the scratch slot)
            const int dst = keep ? cnt + __popc(gbits & gl.below) : 2 * nb;
#ifdef HMC_VERIFY_PREFILTER
            s.plist[HMC_IDX(dst, 2 * nb + 1)] = uint8_t((k + q * nb) | (flagged ? 0x80 : 0));
#else
            s.plist[HMC_IDX(dst, 2 * nb + 1)] = uint8_t(k + q * nb);
#endif
            cnt += __popc(gbits);
          }
        }
        __syncwarp();

        // Pass 2 — exact prediction of the surviving pairs
        for (int base = 0; warp_any<G>(base < cnt); base += G) {
          const int e = base + lane;
          const bool on = e < cnt;
#ifdef HMC_VERIFY_PREFILTER
          const int raw = on ? int(s.plist[HMC_IDX(e, 2 * nb + 1)]) : 0;
          const int idx = raw & 0x7f;
#else
          const int idx = on ? int(s.plist[HMC_IDX(e, 2 * nb + 1)]) : 0;
#endif
          const bool second = idx >= nb;
          const int b = second ? bj : bi;
          const int k = on ? (second ? idx - nb : idx) : b;
          const int d = abs(b - k);
          const int pq = b * PINFO_STRIDE + k;
          double tt;
          unsigned pty = predict_pair<SC>(a, R, s, b, k, b < k, d, __ldg(&R.pinfo[pq]), on, tt);
#ifdef HMC_VERIFY_PREFILTER
          // the pre-filter rejected a real event: flag it and record the first offender
          // (bits 8-13 b, 14-19 k, 20 crossing) for profiles/verify_prefilter.py
          if (on && pty != 0xffu && (raw & 0x80) && !(R.err & 0x40000000u)) {
            R.err |= 0x40000000u | (unsigned(b) << 8) | (unsigned(k) << 14) | (is_cell ? 1u << 20 : 0u);
            if (R.flags == (REP_LIVE | REP_TRACING)) { // dump the pair's state as records of type >= 200
              unsigned n = *a.trace_n;
              auto put = [&](unsigned ty, double v) {
                if (n < a.trace_cap) {
                  hmc_event e;
                  e.t = v;
                  e.type = ty;
                  e.i = unsigned(b);
                  e.j = unsigned(k);
                  e.pad = 0;
                  a.trace[n] = e;
                }
                n++;
              };
              put(200, t); put(201, BT(s, b)); put(202, BT(s, k)); put(203, tt); put(204, double(pty));
              put(205, BX(s, b)); put(206, BY(s, b)); put(207, BZ(s, b));
              put(208, BVX(s, b)); put(209, BVY(s, b)); put(210, BVZ(s, b)); put(211, double(BCELL(s, b)));
              put(212, BX(s, k)); put(213, BY(s, k)); put(214, BZ(s, k));
              put(215, BVX(s, k)); put(216, BVY(s, k)); put(217, BVZ(s, k)); put(218, double(BCELL(s, k)));
              put(219, double(is_cell ? int(wall) : -1));
              *a.trace_n = n;
            }
          }
#endif
          commit_prediction<G>(a, R, s, on, b, k, pty, tt, is_cell, lane, gl);
        }
      }
      __syncwarp();
    }
    // drop what the collision invalidated and find the next event
    next_event<G>(s, R, coll, bi, bj, n_old, lane, gl, bt_io, bk_io);
    __syncwarp();
  }
}

// end of run_step: advance every bead to step_time (main.cc:135-138)
template <int G>
__device__ __forceinline__ void advance_all(const KArgs &a, const Smem &s, double step_time,
                                            int lane, bool on) {
  for (int k = lane; k < a.nb && on; k += G) {
    const double dt = step_time - BT(s, k);
    BX(s, k) += BVX(s, k) * dt;
    BY(s, k) += BVY(s, k) * dt;
    BZ(s, k) += BVZ(s, k) * dt;
    BT(s, k) = step_time;
  }
  __syncwarp();
}

// -------------------------------------------------------------- sampling --

__device__ __forceinline__ void record_config_int(const KArgs &a, const Rep &R, int lane) {
  if (a.cfg_cap && lane == 0 && rep_live(R)) {
    const unsigned n = a.cfg_n[R.r];
    if (n < a.cfg_cap) a.cfg_int[size_t(R.r) * a.cfg_cap + n] = R.config;
    a.cfg_n[R.r] = n + 1;
  }
}

// body of run_trajectory after run_step, main.cc:199-225 (whole warp executes; only
// groups with `on` sample; `step` may differ between the groups of a warp)
template <int G>
__device__ HMC_COLD unsigned sample_step(const KArgs &a, const Rep &R_in, const Smem &s_in,
                                             unsigned step, double e0, int lane, bool on) {
  Rep R = R_in;
  const Smem s = s_in;
  const int trans = R.r / a.replicas_per_trans;
  const bool out = on && rep_live(R); // global side effects
  // dist_between_nonlocal_beads (hardspheres.cc:1084-1104)
  if (a.dist_cap || a.hist_bins) {
    const unsigned nrow = (a.dist_cap && out) ? a.dist_n[R.r] : 0u;
    // the histograms are split by the state of the first transient bond, r_t < rc, the
    // split tools/mfpt.py applies to the dist rows (mfpt.py:104-125)
    int st_on = 0;
    if (a.hist_bins && R.tr->n_transient)
      st_on = dist2_ij(a, s.b + F_X, R.tr->tb_i[0], R.tr->tb_j[0]) < R.tr->rc2 ? 1 : 0;
    for (int k = lane; k < int(R.tr->n_nonlocal) && out; k += G) {
      const int i = R.tr->nl_i[k], j = R.tr->nl_j[k];
      const double d = sqrt(dist2_ij(a, s.b + F_X, i, j));
      if (a.dist_cap && nrow < a.dist_cap)
        a.dist[(size_t(R.r) * a.dist_cap + nrow) * a.dist_stride + k] = d;
      if (a.hist_bins) {
        int bin = int(d * a.hist_scale);
        bin = min(max(bin, 0), int(a.hist_bins) - 1);
        atomicAdd(&a.hist[((size_t(trans) * a.dist_stride + k) * 2 + st_on) * a.hist_bins + bin], 1ull);
      }
    }
    __syncwarp();
    if (a.dist_cap && lane == 0 && out) {
      if (nrow >= a.dist_cap) R.err |= HMC_ERR_SAMPLE_OVERFLOW;
      a.dist_n[R.r] = nrow + 1;
    }
  }
  const bool rec = on && (step % a.write_step == 0);
  if (rec) {
    if (lane == 0 && rep_live(R) && R.config < (unsigned long long)a.nstates) {
      atomicAdd(&a.acc[size_t(trans) * a.nstates + R.config], 1ull);
      if (a.acc_rep) a.acc_rep[size_t(R.r) * a.nstates + R.config] += 1ull; // (its only writer)
    }
    record_config_int(a, R, lane);
  }
  // unique_config: first recorded sample with config == 1 (main.cc:211-216)
  const bool take = rec && rep_live(R) && R.config == 1ull && !a.uniq_found[R.r];
  __syncwarp();
  if (take) {
    for (int k = lane; k < a.nb; k += G) {
      double *up = a.uniq_pos + (size_t(R.r) * a.nb + k) * 3;
      double *uv = a.uniq_vel + (size_t(R.r) * a.nb + k) * 3;
      up[0] = BX(s, k);
      up[1] = BY(s, k);
      up[2] = BZ(s, k);
      uv[0] = BVX(s, k);
      uv[1] = BVY(s, k);
      uv[2] = BVZ(s, k);
    }
    if (lane == 0) a.uniq_found[R.r] = 1;
  }
  __syncwarp();
  const double e1 = hamiltonian(a, R, s);
  const double ediff = fabs(1 - (e1 / e0));
  if (on && !(ediff < 1e-6)) R.err |= HMC_ERR_ENERGY;
  return R.err;
}

// ------------------------------------------------------------- crankshaft --

// crankshaft (crankshaft.cc:224-262): Rodrigues rotation of the chain tail about the
// bond (ind-1, ind), geometric checks, trial bitmask, Metropolis on s_bias.
// Trial coordinates use the velocity arrays as scratch (velocities are redrawn at the
// start of every MD step; they are saved to global memory before the first move).
struct CrankOut {
  unsigned long long config;
  unsigned err, words, moves_done;
};
template <int G>
__device__ HMC_COLD CrankOut crank_moves(const KArgs &a, const Rep &R_in, const Smem &s_in,
                                             const WordStream &ws_in, unsigned move_offset,
                                             unsigned n_moves, int lane, unsigned gmask, bool on) {
  Rep R = R_in;
  const Smem s = s_in;
  WordStream ws = ws_in;
  const int nb = a.nb;
  // dead: this group sits the moves out (it is not at a crankshaft point of its
  // schedule, or its tape ran out)
  bool dead = !on, stuck = false;
  unsigned long long n_att = 0, n_acc = 0;
  for (unsigned mv = 0; mv < n_moves; mv++) {
    const unsigned start = ws.c;
    unsigned attempts = 0;
    bool ok = false;
    // retry loop of crankshaft.cc:245-253; warp-uniform: groups that already hold a
    // valid trial are predicated off until every group of the warp has one
    for (;;) {
      // (the reference's loop is unbounded, crankshaft.cc:245-253: a rotation keeps every
      // nearest / next-nearest distance, so a chain with a local bond marginally outside
      // its well fails the strict check on every trial.  One such replica would hang the
      // whole persistent launch, so the search is capped and the replica flagged.)
      if (!ok && !dead && attempts >= (1u << 20)) {
        dead = true;
        stuck = true;
      }
      const bool need = !ok && !dead;
      if (!warp_any<G>(need)) break;
      unsigned ind = 1;
      double sn = 0.0, cs = 1.0;
      bool drew = false;
      if (need) {
        ind = ws_uniform_int(ws, 1u, unsigned(nb - 2));
        const unsigned cth = ws.c;
        const unsigned w0 = ws_next(ws), w1 = ws_next(ws);
        if (ws.underrun) {
          dead = true;
          ind = 1;
        } else {
          drew = true;
          if (ws.philox) {
            // uniform_real_distribution(0, 2*pi): canonical*(b-a)+a
            const double th = canonical_words(w0, w1) * (2 * 3.14159265358979323846 - 0.0) + 0.0;
            sincos(th, &sn, &cs);
          } else {
            sn = a.tape_sin[size_t(R.r) * a.tape_n + cth];
            cs = a.tape_cos[size_t(R.r) * a.tape_n + cth];
          }
          attempts++;
        }
      }
      // rodrigues_rotation, crankshaft.cc:11-73
      double dx = BX(s, ind) - BX(s, ind - 1);
      double dy = BY(s, ind) - BY(s, ind - 1);
      double dz = BZ(s, ind) - BZ(s, ind - 1);
      mindist3(a, dx, dy, dz);
      const double inv = 1 / sqrt(dx * dx + dy * dy + dz * dz);
      const double kx = dx * inv, ky = dy * inv, kz = dz * inv;
      const double cth1 = 1 - cs;
      const double Rxx = 1 + cth1 * (-kz * kz - ky * ky);
      const double Rxy = -sn * kz + cth1 * ky * kx;
      const double Rxz = sn * ky + cth1 * kz * kx;
      const double Ryx = sn * kz + cth1 * kx * ky;
      const double Ryy = 1 + cth1 * (-kz * kz - kx * kx);
      const double Ryz = -sn * kx + cth1 * kz * ky;
      const double Rzx = -sn * ky + cth1 * kx * kz;
      const double Rzy = sn * kx + cth1 * ky * kz;
      const double Rzz = 1 + cth1 * (-ky * ky - kx * kx);
      const double px = BX(s, ind), py = BY(s, ind), pz = BZ(s, ind);
      __syncwarp();
      if (drew) {
        for (int k = lane; k < nb; k += G) {
          if (k <= int(ind)) {
            BVX(s, k) = BX(s, k);
            BVY(s, k) = BY(s, k);
            BVZ(s, k) = BZ(s, k);
          } else {
            double ox = BX(s, k) - px, oy = BY(s, k) - py, oz = BZ(s, k) - pz;
            mindist3(a, ox, oy, oz);
            double nx = Rxx * ox + Rxy * oy + Rxz * oz;
            double ny = Ryx * ox + Ryy * oy + Ryz * oz;
            double nz = Rzx * ox + Rzy * oy + Rzz * oz;
            mindist3(a, nx, ny, nz);
            BVX(s, k) = px + nx;
            BVY(s, k) = py + ny;
            BVZ(s, k) = pz + nz;
          }
        }
      }
      __syncwarp();
      const bool good = check_constraints<G>(a, R, s.b + F_VX, true, lane, gmask) == 3u;
      if (drew) ok = good;
    }
    bool accept = false;
    unsigned long long trial = 0ull;
    if (dead) {
      if (ws.c != start) { // roll the unfinished move back (positions untouched)
        ws.c = start;
        R.err |= stuck ? HMC_ERR_CRANK_STUCK : HMC_ERR_TAPE_UNDERRUN;
      }
    } else {
      // config_int, crankshaft.cc:175-205: only transient-bond pairs can set bits
      for (unsigned k = 0; k < R.tr->n_transient; k++) {
        const int i = R.tr->tb_i[k], j = R.tr->tb_j[k];
        const double d2 = dist2_ij(a, s.b + F_VX, i, j);
        // first occurrence of a pair in the list owns it (std::find, config.h:34)
        bool first = true;
        for (unsigned q = 0; q < k; q++)
          if (R.tr->tb_i[q] == i && R.tr->tb_j[q] == j) first = false;
        if (first && !(d2 > R.tr->rc2)) trial ^= (1ull << k);
      }
      // accept_move, crankshaft.cc:207-222
      const double dS = rep_sbias(a, R)[trial] - rep_sbias(a, R)[R.config];
      accept = true;
      if (dS > 0) {
        const unsigned w0 = ws_next(ws), w1 = ws_next(ws);
        if (ws.underrun) {
          ws.c = start;
          R.err |= HMC_ERR_TAPE_UNDERRUN;
          dead = true;
          accept = false;
        } else {
          const double u = canonical_words(w0, w1) * (1.0 - 0.0) + 0.0;
          if (u > exp(-dS)) accept = false;
        }
      }
    }
    if (!dead) n_att += attempts;
    __syncwarp();
    if (accept) {
      for (int k = lane; k < nb; k += G) {
        BX(s, k) = BVX(s, k);
        BY(s, k) = BVY(s, k);
        BZ(s, k) = BVZ(s, k);
      }
      R.config = trial;
      n_acc++;
    }
    __syncwarp();
    // main.cc:232-236
    if (!dead) {
      ws.moves_done++;
      const unsigned i = move_offset + mv;
      for (unsigned j = 1; j < a.mc_write; j++)
        if (i == (j * a.mc_moves) / a.mc_write) record_config_int(a, R, lane);
    }
  }
  if (lane == 0 && rep_live(R)) {
    unsigned long long *ev = a.acc + size_t(a.n_trans) * a.nstates + 2 * size_t(a.n_trans);
    if (n_att) atomicAdd(&ev[2], n_att);
    if (n_acc) atomicAdd(&ev[3], n_acc);
  }
  CrankOut out;
  out.config = R.config;
  out.err = R.err;
  out.words = ws.c;
  out.moves_done = ws.moves_done;
  return out;
}

// -------------------------------------------------------------- the kernel --

__device__ __forceinline__ double wl_gamma(const KArgs &a, unsigned iter_wl) {
  return iter_wl == 0 ? a.gamma0 : 1.0 / double(iter_wl); // main.cc:283,302-303
}

template <int G, bool SC, bool TP>
__global__ void __launch_bounds__(HMC_BLOCK, HMC_MIN_BLOCKS) hmc_ensemble_kernel(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // (read once through an opaque asm: under register pressure the compiler otherwise
  // re-reads the special register — an S2R with its latency — at every use)
  unsigned tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int lane = int(tid & unsigned(G - 1));
  const unsigned wlane = tid & 31u;
  const unsigned gmask = group_mask<G>(wlane);
  const GroupLanes<G> gl(lane, wlane);
  const int group_in_cta = int(tid / unsigned(G));
  const int nb = a.nb;
  const Smem s = carve(smem_raw + size_t(group_in_cta) * a.smem_per_replica, a);

  constexpr int NG = 32 / G; // replica groups per warp, run in lockstep
  const int group_in_warp = int(wlane) / G;
  for (;;) {
    int r0 = 0;
    if (wlane == 0) r0 = int(atomicAdd(a.work_counter, unsigned(NG)));
    r0 = __shfl_sync(0xffffffffu, r0, 0);
    if (r0 >= a.n_replicas) break;
    // a group past the end of the ensemble runs a clone of the last replica with all
    // global side effects switched off, so the warp stays converged
    const bool live = r0 + group_in_warp < a.n_replicas;
    const int r = live ? r0 + group_in_warp : a.n_replicas - 1;

    Rep R;
    R.r = r;
    const int trans = r / a.replicas_per_trans;
    R.trans = trans;
    R.tr = a.trans + trans;
    R.pinfo = a.pairinfo + size_t(trans) * size_t(a.nb) * PINFO_STRIDE;
    R.bmask = a.bondmask + size_t(trans) * a.nb;
    R.config = a.config[r];
    R.err = 0;
    R.n_coll = R.n_cell = 0;
    R.nl = 0;
    R.flags = (live ? REP_LIVE : 0u) |
              ((a.trace_cap != 0 && a.trace_replica == r) ? REP_TRACING : 0u);

    // load the replica (bead clocks all equal t_now at a step boundary)
    for (int k = lane; k < nb; k += G) {
      const double *p = a.pos + (size_t(r) * nb + k) * 3;
      BX(s, k) = p[0];
      BY(s, k) = p[1];
      BZ(s, k) = p[2];
      const double *v = a.vel + (size_t(r) * nb + k) * 3;
      BVX(s, k) = v[0];
      BVY(s, k) = v[1];
      BVZ(s, k) = v[2];
      BT(s, k) = a.t_now;
    }
    __syncwarp();

    // crankshaft word source (built where it is used: it is not live across the event loop)
    auto make_ws = [&]() {
      WordStream ws;
      ws.tape = a.tape_words ? a.tape_words + size_t(r) * a.tape_n : nullptr;
      ws.n = a.tape_n;
      ws.c = 0;
      ws.underrun = false;
      ws.philox = (a.rng_mode == HMC_RNG_PHILOX);
      ws.key = make_uint2(a.seed0, a.seed1);
      ws.c1 = ws.c2 = 0;
      ws.c3 = rep_rng_id(a, R);
      ws.buf_blk = 0xffffffffu;
      ws.moves_done = 0;
      return ws;
    };
    // valid events of the finished interval -> global counters
    auto flush_events = [&]() {
      if (lane == 0 && live) {
        unsigned long long *ev = a.acc + size_t(a.n_trans) * a.nstates + 2 * size_t(a.n_trans);
        if (R.n_coll) atomicAdd(&ev[0], (unsigned long long)R.n_coll);
        if (R.n_cell) atomicAdd(&ev[1], (unsigned long long)R.n_cell);
      }
      R.n_coll = R.n_cell = 0;
    };

    bool vel_dirty = false; // velocity arrays hold crankshaft scratch
    auto save_vel = [&](bool on) {
      if (live && on) {
        for (int k = lane; k < nb; k += G) {
          double *v = a.vel + (size_t(r) * nb + k) * 3;
          v[0] = BVX(s, k);
          v[1] = BVY(s, k);
          v[2] = BVZ(s, k);
        }
      }
      __syncwarp();
    };

    if (a.op == HMC_OP_CRANK_ONLY) {
      const WordStream wc = make_ws();
      const Rep Rc = R;
      const Smem sc = s;
      const CrankOut co = crank_moves<G>(a, Rc, sc, wc, a.move_offset, a.count, lane, gmask, true);
      R.config = co.config;
      R.err = co.err;
      vel_dirty = true;
      if (lane == 0 && live && a.tape_consumed) {
        a.tape_consumed[r] = co.words;
        a.tape_consumed[a.n_replicas + r] = co.moves_done;
      }
    } else {
      // ---- the replica's schedule as a flat list of event steps e = 0 .. E-1 ----
      //   EQ / MAIN : one initialize_system per step (main.cc:154-163, 185-197)
      //   WL        : one initialize_system per trajectory of nsteps_wl steps
      //               (run_trajectory_wl, main.cc:249-276), then the +-gamma update
      // The groups of a warp walk this list independently: a group whose step has no
      // event left does its step-end / next-step work while the others keep processing
      // events (all of it executed warp-wide under predicates), so nobody waits for the
      // slowest replica of the warp at every step.
      const bool wl = a.phase == HMC_PHASE_WL;
      const unsigned per = wl ? a.nsteps_wl : (a.op == HMC_OP_RUN ? (a.phase == HMC_PHASE_EQ ? a.nsteps_eq : a.nsteps) : 1u);
      const unsigned nunits = wl && a.op == HMC_OP_RUN ? a.count * unsigned(a.nstates) : a.count;
      const unsigned E = nunits * per; // units: iterations (RUN), trajectories (WL), steps (MD_ONLY)
      // (a skipped replica's schedule is empty)
      unsigned e = (a.skip && a.skip[r]) ? E : 0u; // next / current event step of this group
      unsigned ninit = 0;              // initialize_system calls so far (RNG draw index)
      bool in_step = false;
      double step_time = 0.0, e0 = 0.0;
      unsigned cur_st = 0;
      double bt = CUDART_INF; // next event of the replica (carried from event to event)
      unsigned bk = 0u;
      for (;;) {
        // -- start the next step --
        const bool start = !in_step && e < E;
        if (warp_any<G>(start)) {
          const unsigned unit = e / per, sub = e - unit * per;
          unsigned st;
          if (wl) {
            const unsigned iter_wl = a.op == HMC_OP_RUN ? a.first + unit / unsigned(a.nstates)
                                                        : (a.first + unit) / unsigned(a.nstates);
            st = iter_wl * a.nsteps_wl + sub;
          } else {
            st = a.op == HMC_OP_RUN ? (a.first + unit) * per + sub : a.first + unit;
          }
          const bool do_init = start && (!wl || sub == 0);
          if (warp_any<G>(do_init)) {
            const double *nrm =
                (a.normals && do_init) ? a.normals + (size_t(r) * nunits + ninit) * size_t(nb) * 3
                                       : nullptr;
            // (the callee gets COPIES: if the address of R or s themselves escaped into a call
            // they could no longer live in registers during the event loop)
            const Rep Rc = R;
            const Smem sc = s;
            const StepOut so =
                init_step<G, SC>(a, Rc, sc, nrm, a.md_draw_base + ninit, lane, wlane, gmask, do_init);
            R.err = so.err;
            R.nl = so.nl;
            const double h0 = hamiltonian(a, R, s);
            if (do_init) {
              e0 = h0;
              ninit++;
            }
          }
          if (start) {
            cur_st = st;
            step_time = double(st) * (wl ? a.del_t_wl : a.del_t);
            in_step = true;
          }
          // first event of the step (for the groups in the middle of a step this recomputes
          // what they carry: their tables have not changed)
          next_event<G>(s, R, false, 0, 0, 0, lane, gl, bt, bk);
        }
        // -- next event of the current step --
        // a step of the shipped configurations has 5e2-1e4 events; a runaway replica
        // (non-finite state) is flagged and its step cut short instead of hanging
        const bool overrun = in_step && (R.n_coll + R.n_cell) >= a.max_events_per_step;
        if (overrun) R.err |= HMC_ERR_NAN;
        const bool has = in_step && (bt <= step_time) && !overrun;
        // -- finish the step --
        const bool finish = in_step && !has;
        if (warp_any<G>(finish)) {
          advance_all<G>(a, s, step_time, lane, finish);
          if (finish) flush_events();
          const unsigned unit = e / per, sub = e - unit * per;
          if (a.phase == HMC_PHASE_MAIN) {
            const Rep Rc = R;
            const Smem sc = s;
            R.err = sample_step<G>(a, Rc, sc, cur_st, e0, lane, finish);
          } else if (wl) {
            const double e1 = hamiltonian(a, R, s);
            if (finish && !(fabs(1 - (e1 / e0)) < 1e-6)) R.err |= HMC_ERR_ENERGY;
            // wang_landau, main.cc:288-304
            if (finish && sub + 1 == per && lane == 0 && live) {
              const unsigned iter_wl = a.op == HMC_OP_RUN ? a.first + unit / unsigned(a.nstates)
                                                          : (a.first + unit) / unsigned(a.nstates);
              const double g = wl_gamma(a, iter_wl);
              double *sb = rep_sbias(a, R);
              if (R.config == 0ull) sb[a.nstates - 1] -= g;
              else sb[a.nstates - 1] += g;
            }
            __syncwarp();
          }
          // crankshaft block after the last MD step of an iteration (main.cc:228-237)
          const bool crank = finish && a.op == HMC_OP_RUN && a.phase == HMC_PHASE_MAIN &&
                             a.mc_moves != 0 && sub + 1 == per;
          if (warp_any<G>(crank)) {
            save_vel(crank);
            WordStream wc = make_ws();
            if (crank) {
              const unsigned long long mcd = a.mc_draw_base + unit;
              wc.c1 = unsigned(mcd);
              wc.c2 = unsigned(mcd >> 32) | 0x80000000u;
              vel_dirty = true;
            }
            const Rep Rc = R;
            const Smem sc = s;
            const CrankOut co = crank_moves<G>(a, Rc, sc, wc, 0u, a.mc_moves, lane, gmask, crank);
            R.config = co.config;
            R.err = co.err;
          }
          if (finish) {
            in_step = false;
            e++;
            vel_dirty = crank; // a later init_step redraws the velocities
          }
        }
        if (!warp_any<G>(has)) {
          if (!warp_any<G>(in_step || e < E)) break;
          continue;
        }
        event_body<G, SC, TP>(a, R, s, has, bt, bk, lane, gl);
      }
    }

    // store the replica back
    __syncwarp();
    if (live) {
      for (int k = lane; k < nb; k += G) {
        double *p = a.pos + (size_t(r) * nb + k) * 3;
        p[0] = BX(s, k);
        p[1] = BY(s, k);
        p[2] = BZ(s, k);
      }
    }
    save_vel(!vel_dirty);
    flush_events();
    if (lane == 0 && live) {
      a.config[r] = R.config;
      if (R.err) atomicOr(&a.err[r], R.err);
    }
    __syncwarp();
  }
}

} // namespace

// ---------------------------------------------------------------- launcher --

int hmc_pick_group(int nb);

// bytes of a replica image besides the event list
static size_t smem_fixed(int nb) { return 8 * size_t(BSTRIDE) * nb + 8 * size_t(nb) + 2 * size_t(nb) + 1 + 32; }

int hmc_list_capacity(int nb, int group, int override_entries) {
  // Live nonlocal predictions per replica: ~0.5 nb on average; the list transiently also
  // holds the predictions of the running event.  Peaks measured on the CPU (2e5-event
  // probes): 1.5 nb for 46-50 beads, 2.4 nb for 20 beads with permanent bonds, with a slowly
  // decaying tail (20 beads overflowed 64 entries once in 1e9 events).  Wanted:
  // max(3.5 nb, 96); at least 3 nb; in between, whatever keeps 16 warps per SM resident.
  // An overflow raises HMC_ERR_LIST_OVERFLOW.  A multiple of 8.
  const int nonlocal = nb >= 4 ? (nb - 3) * (nb - 2) / 2 : 0;
  int want = (7 * nb + 1) / 2;
  if (want < 96) want = 96;
  int cap = want;
  const char *e = getenv("HMC_LIST_CAP"); // tuning / test knob
  const int forced = override_entries > 0 ? override_entries : (e ? atoi(e) : 0);
  if (forced > 0) { // hmc_set_list_capacity: compact states (many bonds formed) need more
    cap = want = forced;
  } else {
    const long long replicas = 16ll * (32 / group);               // per SM at 16 warps
    const long long budget = (233472ll - 4 * (1024 + 512)) / replicas; // bytes per replica
    const long long fit = (budget - (long long)smem_fixed(nb)) / 16;
    if (fit < cap) cap = int(fit > 3 * nb ? fit : 3 * nb);
    if (cap > want) cap = want;
  }
  // (every nonlocal pair live plus the predictions of one event: more cannot be in the list)
  if (cap > nonlocal + 2 * nb) cap = nonlocal + 2 * nb;
  if (cap < 8) cap = 8;
  return (cap + 7) & ~7;
}

// byte offsets of the arrays inside a replica image (see Smem)
void hmc_smem_layout(int nb, int group, int cap_override, uint32_t *off_tnn, uint32_t *off_list,
                     uint32_t *off_plist, uint32_t *total) {
  const size_t cap = size_t(hmc_list_capacity(nb, group, cap_override));
  const size_t tnn = 8 * size_t(BSTRIDE) * nb;   // bead records are 80 B: 16-byte aligned
  size_t list = tnn + 8 * size_t(nb);
  list = (list + 15) & ~size_t(15);               // 16-byte entries
  const size_t plist = list + 16 * cap;
  const size_t bytes = (plist + 2 * size_t(nb) + 1 + 15) & ~size_t(15);
  if (off_tnn) *off_tnn = uint32_t(tnn);
  if (off_list) *off_list = uint32_t(list);
  if (off_plist) *off_plist = uint32_t(plist);
  if (total) *total = uint32_t(bytes);
}

size_t hmc_smem_per_replica(int nb, int group) {
  uint32_t total = 0;
  hmc_smem_layout(nb, group, 0, nullptr, nullptr, nullptr, &total);
  return total;
}

int hmc_pick_group(int nb) {
  // lanes per replica: the per-event work that every lane of a group repeats (event
  // arithmetic, cell walls, the reduce of the argmin) is shared by the 32/G replicas of a
  // warp, so narrow groups win as long as the partner loops keep their lanes busy and the
  // replica images of 16 warps fit in shared memory: 8 lanes up to 24 beads, 16 beyond
  // (46 beads: 6.4 KB per replica, 32 replicas per SM).
  if (nb <= 24) return 8;
  return 16;
}

template <int G, bool SC, bool TP>
static cudaError_t launch_g(const KArgs &a, cudaStream_t stream, int *grid_out, int *block_out) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // CTA size: HMC_BLOCK threads, or half of it when shared memory (long chains) lets
  // more warps stay resident that way
  int block = HMC_BLOCK, per_sm = 0;
  size_t smem = 0;
  int best_warps = -1;
  for (int cand = HMC_BLOCK; cand >= 64 && cand >= G; cand /= 2) {
    // (+512: the list scans read whole lane strides, up to 31 entries past the capacity)
    const size_t sm = size_t(cand / G) * a.smem_per_replica + 512;
    cudaError_t e = cudaFuncSetAttribute(hmc_ensemble_kernel<G, SC, TP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm));
    if (e != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, hmc_ensemble_kernel<G, SC, TP>, cand, sm);
    if (e != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    if (n * cand / 32 > best_warps) {
      best_warps = n * cand / 32;
      block = cand;
      per_sm = n;
      smem = sm;
    }
  }
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(hmc_ensemble_kernel<G, SC, TP>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  // persistent grid: every resident warp slot, never more CTAs than work.  (Warps pull
  // batches of 32/G replicas and keep each for its whole schedule.  Measured: spreading the
  // batches over fewer, equally loaded warps is slower — a warp's pace hardly depends on how
  // many share its SM, so only the number of rounds counts.)
  const int groups_per_cta = block / G;
  long long need = (a.n_replicas + groups_per_cta - 1) / groups_per_cta;
  long long grid = (long long)sms * per_sm;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  hmc_ensemble_kernel<G, SC, TP><<<int(grid), block, smem, stream>>>(a);
  if (grid_out) *grid_out = int(grid);
  if (block_out) *block_out = block;
  return cudaGetLastError();
}

template <bool SC>
static cudaError_t launch_sc(const KArgs &a, int group, cudaStream_t stream, int *grid_out,
                             int *block_out) {
  // TP: partner re-prediction as pre-filter + compacted exact pass (chains >= 16 beads)
  if (!a.two_pass)
    return group == 8 ? launch_g<8, SC, false>(a, stream, grid_out, block_out)
                      : launch_g<16, SC, false>(a, stream, grid_out, block_out);
  switch (group) {
  case 8: return launch_g<8, SC, true>(a, stream, grid_out, block_out);
  case 32: return launch_g<32, SC, true>(a, stream, grid_out, block_out);
  default: return launch_g<16, SC, true>(a, stream, grid_out, block_out);
  }
}

cudaError_t hmc_launch(const KArgs &a, int group, cudaStream_t stream, int *grid_out,
                       int *block_out) {
  // SC: min-image fast path for chains shorter than 1.5 boxes, chosen per launch
  return a.short_chain ? launch_sc<true>(a, group, stream, grid_out, block_out)
                       : launch_sc<false>(a, group, stream, grid_out, block_out);
}

// violations counted by a -DHMC_DEBUG_BOUNDS build so far; -1: not such a build
long long hmc_bounds_violations_device() {
#ifdef HMC_DEBUG_BOUNDS
  unsigned long long v = 0;
  if (cudaMemcpyFromSymbol(&v, g_bounds_violations, sizeof v) != cudaSuccess) return -2;
  return (long long)v;
#else
  return -1;
#endif
}
