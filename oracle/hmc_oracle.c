/*
 * hmc_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see hmc_oracle.h).
 *
 * CPU restatement, in plain C, of the algorithm of margaritacolberg/hybridmc's
 * trajectory path.  Every function cites the reference file:line it follows.
 * The arithmetic keeps the reference's IEEE operation order (build with
 * -O2 -ffp-contract=off, no -ffast-math) so event traces are bit-identical to
 * the unmodified reference sources (oracle/_ref).
 *
 * Structure is deliberately different from the reference (one tagged event
 * record and one process routine instead of eight overloads, flat arrays
 * instead of vector-of-vectors, a hand-written binary heap that reproduces
 * libstdc++'s push_heap/pop_heap sift order so ties resolve identically).
 */
#include "hmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NBMAX HMC_MAX_BEADS

typedef struct {
  double t;
  uint64_t ni, nj;
  uint32_t i, j; /* for cell events j = wall */
  uint8_t type;
  uint8_t cxn, cyn, czn; /* target cell (cell events) */
} ev_t;

struct orc_sim {
  hmc_params p;
  unsigned nb, ncell, nstates;
  double l, inv_l, lcell, inv_lcell;
  double near_min2, near_max2, nnear_min2, nnear_max2, rh2, rc2, stair2, p_rc2;
  double pos[NBMAX][3], vel[NBMAX][3], times[NBMAX];
  uint64_t counter[NBMAX];
  double *s_bias;
  uint64_t config;
  /* bond masks per ordered pair i<j (NonlocalBonds::get_bond_mask, config.h:30-45) */
  uint64_t tmask[NBMAX][NBMAX], pmask[NBMAX][NBMAX], nlmask[NBMAX][NBMAX];
  /* cells (cells.h:13-30) */
  unsigned cx[NBMAX], cy[NBMAX], cz[NBMAX];
  uint8_t *cell_list; /* [ncell^3][nb] */
  uint8_t *cell_cnt;  /* [ncell^3]     */
  /* event queue */
  ev_t *heap;
  size_t hn, hcap;
  /* statistics */
  uint64_t *config_count;
  uint64_t formed, broken;
  uint64_t n_coll, n_cell, n_attempt, n_accept, n_pop, n_pred;
  /* samples */
  hmc_event *trace;
  unsigned trace_cap, trace_n;
  double *dist;
  unsigned dist_cap, dist_n;
  uint64_t *cfg_int;
  unsigned cfg_cap, cfg_n;
  int uniq_found;
  double uniq_pos[NBMAX][3], uniq_vel[NBMAX][3];
  unsigned last_moves_done;
};

/* ------------------------------------------------------------------ box.h -- */

/* Box::mindist, box.h:22-26 (std::round = half away from zero) */
static inline void mindist(const orc_sim *s, double *dx, double *dy, double *dz) {
  *dx -= round(*dx * s->inv_l) * s->l;
  *dy -= round(*dy * s->inv_l) * s->l;
  *dz -= round(*dz * s->inv_l) * s->l;
}

/* Box::minpos, box.h:16-20 */
static inline void minpos(const orc_sim *s, double *x, double *y, double *z) {
  *x -= floor(*x * s->inv_l) * s->l;
  *y -= floor(*y * s->inv_l) * s->l;
  *z -= floor(*z * s->inv_l) * s->l;
}

void orc_mindist(double l, double *d) {
  orc_sim b;
  b.l = l;
  b.inv_l = 1.0 / l;
  mindist(&b, &d[0], &d[1], &d[2]);
}
void orc_minpos(double l, double *x) {
  orc_sim b;
  b.l = l;
  b.inv_l = 1.0 / l;
  minpos(&b, &x[0], &x[1], &x[2]);
}

/* --------------------------------------------------- collision-time kernels -- */

/* t_until_inner_coll, hardspheres.cc:217-240 */
int orc_t_until_inner_coll(double x, double y, double z, double vx, double vy,
                           double vz, double sigma2, double *t) {
  const double rv = x * vx + y * vy + z * vz;
  if (!(rv < 0)) return 0;
  const double vv = vx * vx + vy * vy + vz * vz;
  const double rr = x * x + y * y + z * z;
  const double rad1 = vv * (rr - sigma2);
  const double rv2 = rv * rv;
  if (!(rv2 > rad1)) return 0;
  *t += (-rv - sqrt(rv2 - rad1)) / vv;
  return 1;
}

/* t_until_outer_coll, hardspheres.cc:244-254 */
void orc_t_until_outer_coll(double x, double y, double z, double vx, double vy,
                            double vz, double sigma2, double *t) {
  const double rv = x * vx + y * vy + z * vz;
  const double vv = vx * vx + vy * vy + vz * vz;
  const double rr = x * x + y * y + z * z;
  const double rad1 = vv * (rr - sigma2);
  const double rv2 = rv * rv;
  *t += (-rv + sqrt(rv2 - rad1)) / vv;
}

/* v_after_coll, hardspheres.cc:258-295; returns 1 on transmission */
int orc_v_after_coll(double x, double y, double z, double sigma2, double *vi,
                     double *vj, int has_dS, double dS, double m) {
  const double vx = vj[0] - vi[0], vy = vj[1] - vi[1], vz = vj[2] - vi[2];
  const double mu = 0.5 * m;
  double rv = (x * vx + y * vy + z * vz) / sigma2;
  const double rv2 = rv * rv;
  double dS_norm = 0.0;
  int transmitted;
  if (has_dS && rv2 > (dS_norm = 2 * dS / (sigma2 * mu))) {
    rv -= copysign(sqrt(rv2 - dS_norm), rv);
    transmitted = 1;
  } else {
    rv += rv;
    transmitted = 0;
  }
  const double px = rv * x, py = rv * y, pz = rv * z;
  vi[0] += px / 2;
  vi[1] += py / 2;
  vi[2] += pz / 2;
  vj[0] -= px / 2;
  vj[1] -= py / 2;
  vj[2] -= pz / 2;
  return transmitted;
}

/* get_rc2_inner, hardspheres.cc:394-403 */
static inline double rc2_inner(const orc_sim *s, uint64_t p_mask) {
  return (p_mask && s->p.has_p_rc) ? s->p_rc2 : s->rc2;
}
/* get_rc2_outer, hardspheres.cc:405-419 */
static inline double rc2_outer(const orc_sim *s, uint64_t t_mask, uint64_t p_mask) {
  double r = s->rc2;
  if (p_mask && s->p.has_p_rc) r = s->p_rc2;
  if (s->p.has_stair && (~s->config & t_mask)) r = s->stair2;
  return r;
}
double orc_rc2_inner(const hmc_params *p, int p_bond) {
  return (p_bond && p->has_p_rc) ? p->p_rc * p->p_rc : p->rc * p->rc;
}
double orc_rc2_outer(const hmc_params *p, int t_nonbonded, int p_bond) {
  double r = p->rc * p->rc;
  if (p_bond && p->has_p_rc) r = p->p_rc * p->p_rc;
  if (p->has_stair && t_nonbonded) r = p->stair * p->stair;
  return r;
}

/* ------------------------------------------------------------------- heap -- */
/* std::priority_queue<Event, vector, greater> (event.h:105-116) restated with
 * libstdc++'s __push_heap / __adjust_heap sift order (bits/stl_heap.h). */

static void heap_sift_up(ev_t *h, size_t hole, size_t top, ev_t v) {
  size_t parent = (hole - 1) / 2;
  while (hole > top && h[parent].t > v.t) {
    h[hole] = h[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  h[hole] = v;
}

static void heap_push(orc_sim *s, ev_t v) {
  if (s->hn == s->hcap) {
    s->hcap = s->hcap ? 2 * s->hcap : 1024;
    s->heap = (ev_t *)realloc(s->heap, s->hcap * sizeof(ev_t));
  }
  s->hn++;
  heap_sift_up(s->heap, s->hn - 1, 0, v);
}

static ev_t heap_pop(orc_sim *s) {
  ev_t *h = s->heap;
  const ev_t top = h[0];
  if (s->hn > 1) {
    const ev_t v = h[s->hn - 1];
    const size_t len = s->hn - 1;
    size_t hole = 0, child = 0;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (h[child].t > h[child - 1].t) child--;
      h[hole] = h[child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      h[hole] = h[child - 1];
      hole = child - 1;
    }
    heap_sift_up(h, hole, 0, v);
  }
  s->hn--;
  return top;
}

/* ------------------------------------------------------------------ cells -- */

static inline unsigned cell_index(const orc_sim *s, unsigned x, unsigned y, unsigned z) {
  return x + y * s->ncell + z * s->ncell * s->ncell;
}

/* init_cells, hardspheres.cc:334-369 */
static void init_cells(orc_sim *s) {
  const unsigned nc3 = s->ncell * s->ncell * s->ncell;
  memset(s->cell_cnt, 0, nc3);
  for (unsigned j = 0; j < s->nb; j++) {
    double x = s->pos[j][0], y = s->pos[j][1], z = s->pos[j][2];
    minpos(s, &x, &y, &z);
    int ix = (int)(s->inv_lcell * x), iy = (int)(s->inv_lcell * y),
        iz = (int)(s->inv_lcell * z);
    const int hi = (int)s->ncell - 1;
    ix = ix < 0 ? 0 : (ix > hi ? hi : ix);
    iy = iy < 0 ? 0 : (iy > hi ? hi : iy);
    iz = iz < 0 ? 0 : (iz > hi ? hi : iz);
    s->cx[j] = ix;
    s->cy[j] = iy;
    s->cz[j] = iz;
    const unsigned c = cell_index(s, ix, iy, iz);
    s->cell_list[c * s->nb + s->cell_cnt[c]++] = (uint8_t)j;
  }
}

/* move_to_new_cell, hardspheres.cc:707-736 (swap-erase keeps list order) */
static void move_to_new_cell(orc_sim *s, unsigned i, unsigned xn, unsigned yn, unsigned zn) {
  const unsigned oc = cell_index(s, s->cx[i], s->cy[i], s->cz[i]);
  const unsigned nc = cell_index(s, xn, yn, zn);
  uint8_t *ol = s->cell_list + oc * s->nb;
  unsigned k = 0;
  while (ol[k] != i) k++;
  s->cell_list[nc * s->nb + s->cell_cnt[nc]++] = (uint8_t)i;
  /* re-read: nc may equal oc only if ncell == 1, excluded by ncell >= 4 */
  ol[k] = ol[s->cell_cnt[oc] - 1];
  s->cell_cnt[oc]--;
  s->cx[i] = xn;
  s->cy[i] = yn;
  s->cz[i] = zn;
}

/* t_until_cell, hardspheres.cc:560-671 */
static int t_until_cell(const orc_sim *s, double xi, double yi, double zi, double vx,
                        double vy, double vz, unsigned ixo, unsigned iyo, unsigned izo,
                        double *dt, unsigned *wall, unsigned *xn, unsigned *yn,
                        unsigned *zn) {
  int crossing = 0;
  const unsigned nc = s->ncell;
  double dx = s->lcell * ixo - xi;
  double dy = s->lcell * iyo - yi;
  double dz = s->lcell * izo - zi;
  mindist(s, &dx, &dy, &dz);
  if (vx > 0) {
    const double c = (dx + s->lcell) / vx;
    if (!crossing || c < *dt) {
      crossing = 1, *dt = c, *wall = 0;
      *xn = (ixo + 1) % nc, *yn = iyo, *zn = izo;
    }
  } else if (vx < 0) {
    const double c = dx / vx;
    if (!crossing || c < *dt) {
      crossing = 1, *dt = c, *wall = 1;
      *xn = (ixo - 1 + nc) % nc, *yn = iyo, *zn = izo;
    }
  }
  if (vy > 0) {
    const double c = (dy + s->lcell) / vy;
    if (!crossing || c < *dt) {
      crossing = 1, *dt = c, *wall = 2;
      *xn = ixo, *yn = (iyo + 1) % nc, *zn = izo;
    }
  } else if (vy < 0) {
    const double c = dy / vy;
    if (!crossing || c < *dt) {
      crossing = 1, *dt = c, *wall = 3;
      *xn = ixo, *yn = (iyo - 1 + nc) % nc, *zn = izo;
    }
  }
  if (vz > 0) {
    const double c = (dz + s->lcell) / vz;
    if (!crossing || c < *dt) {
      crossing = 1, *dt = c, *wall = 4;
      *xn = ixo, *yn = iyo, *zn = (izo + 1) % nc;
    }
  } else if (vz < 0) {
    const double c = dz / vz;
    if (!crossing || c < *dt) {
      crossing = 1, *dt = c, *wall = 5;
      *xn = ixo, *yn = iyo, *zn = (izo - 1 + nc) % nc;
    }
  }
  return crossing;
}

int orc_t_until_cell(double x, double y, double z, double l, unsigned ncell,
                     double lcell, double vx, double vy, double vz, unsigned ixo,
                     unsigned iyo, unsigned izo, double *dt, unsigned *wall,
                     unsigned *cell_new) {
  orc_sim b;
  b.l = l;
  b.inv_l = 1.0 / l;
  b.ncell = ncell;
  b.lcell = lcell;
  return t_until_cell(&b, x, y, z, vx, vy, vz, ixo, iyo, izo, dt, wall, &cell_new[0],
                      &cell_new[1], &cell_new[2]);
}

/* if_cell, hardspheres.cc:678-693 */
static void if_cell(orc_sim *s, unsigned i) {
  double dt = 0.0;
  unsigned wall = 0, xn = 0, yn = 0, zn = 0;
  if (t_until_cell(s, s->pos[i][0], s->pos[i][1], s->pos[i][2], s->vel[i][0],
                   s->vel[i][1], s->vel[i][2], s->cx[i], s->cy[i], s->cz[i], &dt, &wall,
                   &xn, &yn, &zn)) {
    ev_t e;
    memset(&e, 0, sizeof e);
    e.t = s->times[i] + dt;
    e.type = HMC_EV_BEAD_CELL;
    e.i = i;
    e.j = wall;
    e.ni = s->counter[i];
    e.cxn = (uint8_t)xn, e.cyn = (uint8_t)yn, e.czn = (uint8_t)zn;
    heap_push(s, e);
  }
}

/* -------------------------------------------------------- pair prediction -- */

/* delta_tpv, hardspheres.cc:372-392 */
static void delta_tpv(const orc_sim *s, unsigned i, unsigned j, double *t, double *dx,
                      double *dy, double *dz, double *dvx, double *dvy, double *dvz) {
  if (s->times[i] < s->times[j]) {
    const unsigned k = i;
    i = j;
    j = k;
  }
  *t = s->times[i];
  const double dt = *t - s->times[j];
  *dx = s->pos[j][0] + s->vel[j][0] * dt - s->pos[i][0];
  *dy = s->pos[j][1] + s->vel[j][1] * dt - s->pos[i][1];
  *dz = s->pos[j][2] + s->vel[j][2] * dt - s->pos[i][2];
  mindist(s, dx, dy, dz);
  *dvx = s->vel[j][0] - s->vel[i][0];
  *dvy = s->vel[j][1] - s->vel[i][1];
  *dvz = s->vel[j][2] - s->vel[i][2];
}

static void push_pair(orc_sim *s, double t, unsigned type, unsigned i, unsigned j) {
  ev_t e;
  memset(&e, 0, sizeof e);
  e.t = t;
  e.type = (uint8_t)type;
  e.i = i;
  e.j = j;
  e.ni = s->counter[i];
  e.nj = s->counter[j];
  heap_push(s, e);
}

/* if_coll, hardspheres.cc:425-476: nonlocal pair -> at most one event, with the
 * priority bond capture (rc) > hard core (rh) > outer wall. */
static void if_coll(orc_sim *s, unsigned i, unsigned j) {
  if (i > j) {
    const unsigned k = i;
    i = j;
    j = k;
  }
  if (j == i || j == i + 1 || j == i + 2) return;
  double t, dx, dy, dz, dvx, dvy, dvz;
  delta_tpv(s, i, j, &t, &dx, &dy, &dz, &dvx, &dvy, &dvz);
  s->n_pred++;
  const uint64_t tm = s->tmask[i][j], pm = s->pmask[i][j];
  const double r2i = rc2_inner(s, pm);
  const double r2o = rc2_outer(s, tm, pm);
  const unsigned nbonds = (unsigned)__builtin_popcountll(s->config);
  if (tm && nbonds < s->p.max_nbonds && (~s->config & tm) &&
      orc_t_until_inner_coll(dx, dy, dz, dvx, dvy, dvz, r2i, &t)) {
    push_pair(s, t, HMC_EV_MAX_NONLOCAL_INNER, i, j);
  } else if (orc_t_until_inner_coll(dx, dy, dz, dvx, dvy, dvz, s->rh2, &t)) {
    push_pair(s, t, HMC_EV_MIN_NONLOCAL_INNER, i, j);
  } else if ((s->config & tm) || pm || (s->p.has_stair && (~s->config & tm))) {
    orc_t_until_outer_coll(dx, dy, dz, dvx, dvy, dvz, r2o, &t);
    push_pair(s, t, HMC_EV_MAX_NONLOCAL_OUTER, i, j);
  }
}

/* if_nearest_bond / if_nnearest_bond, hardspheres.cc:863-923 */
static void if_local(orc_sim *s, unsigned i, unsigned j, double min2, double max2,
                     unsigned type_min) {
  double t, dx, dy, dz, dvx, dvy, dvz;
  delta_tpv(s, i, j, &t, &dx, &dy, &dz, &dvx, &dvy, &dvz);
  s->n_pred++;
  if (orc_t_until_inner_coll(dx, dy, dz, dvx, dvy, dvz, min2, &t)) {
    push_pair(s, t, type_min, i, j);
  } else {
    orc_t_until_outer_coll(dx, dy, dz, dvx, dvy, dvz, max2, &t);
    push_pair(s, t, type_min + 1, i, j);
  }
}
static inline void if_near(orc_sim *s, unsigned i, unsigned j) {
  if_local(s, i, j, s->near_min2, s->near_max2, HMC_EV_MIN_NEAREST);
}
static inline void if_nnear(orc_sim *s, unsigned i, unsigned j) {
  if_local(s, i, j, s->nnear_min2, s->nnear_max2, HMC_EV_MIN_NNEAREST);
}

/* iterate_coll, hardspheres.cc:479-492 */
static void scan_cell(orc_sim *s, unsigned c, unsigned i) {
  const uint8_t *lst = s->cell_list + c * s->nb;
  const unsigned n = s->cell_cnt[c];
  for (unsigned k = 0; k < n; k++) if_coll(s, i, lst[k]);
}

/* add_events_for_one_bead, hardspheres.cc:507-536 (27 cells, z-y-x order) */
static void scan_27(orc_sim *s, unsigned i) {
  const unsigned nc = s->ncell;
  for (unsigned z = s->cz[i] - 1 + nc; z <= s->cz[i] + 1 + nc; z++)
    for (unsigned y = s->cy[i] - 1 + nc; y <= s->cy[i] + 1 + nc; y++)
      for (unsigned x = s->cx[i] - 1 + nc; x <= s->cx[i] + 1 + nc; x++)
        scan_cell(s, cell_index(s, x % nc, y % nc, z % nc), i);
}

/* add_events_for_bead_after_crossing, hardspheres.cc:740-856 (9 new cells) */
static void scan_9(orc_sim *s, unsigned i, unsigned wall) {
  const unsigned nc = s->ncell;
  const unsigned lo[3] = {s->cx[i] - 1 + nc, s->cy[i] - 1 + nc, s->cz[i] - 1 + nc};
  const unsigned hi[3] = {s->cx[i] + 1 + nc, s->cy[i] + 1 + nc, s->cz[i] + 1 + nc};
  const unsigned axis = wall / 2;
  const unsigned fixed = ((wall & 1) ? lo[axis] : hi[axis]) % nc;
  /* loop order of the reference: outer z (or y), inner y (or x) */
  if (axis == 0) {
    for (unsigned z = lo[2]; z <= hi[2]; z++)
      for (unsigned y = lo[1]; y <= hi[1]; y++)
        scan_cell(s, cell_index(s, fixed, y % nc, z % nc), i);
  } else if (axis == 1) {
    for (unsigned z = lo[2]; z <= hi[2]; z++)
      for (unsigned x = lo[0]; x <= hi[0]; x++)
        scan_cell(s, cell_index(s, x % nc, fixed, z % nc), i);
  } else {
    for (unsigned y = lo[1]; y <= hi[1]; y++)
      for (unsigned x = lo[0]; x <= hi[0]; x++)
        scan_cell(s, cell_index(s, x % nc, y % nc, fixed), i);
  }
}

/* ------------------------------------------------------- event processing -- */

/* update_pos, hardspheres.cc:1040-1047 */
static inline void update_pos(orc_sim *s, unsigned i, double t) {
  const double dt = t - s->times[i];
  s->pos[i][0] += s->vel[i][0] * dt;
  s->pos[i][1] += s->vel[i][1] * dt;
  s->pos[i][2] += s->vel[i][2] * dt;
  s->times[i] = t;
}

/* final_vel, hardspheres.cc:1050-1061 */
static int final_vel(orc_sim *s, double sigma2, unsigned i, unsigned j, int has_dS, double dS) {
  double dx = s->pos[j][0] - s->pos[i][0];
  double dy = s->pos[j][1] - s->pos[i][1];
  double dz = s->pos[j][2] - s->pos[i][2];
  mindist(s, &dx, &dy, &dz);
  return orc_v_after_coll(dx, dy, dz, sigma2, s->vel[i], s->vel[j], has_dS, dS, s->p.m);
}

static void trace_event(orc_sim *s, const ev_t *e) {
  if (s->trace_n < s->trace_cap) {
    hmc_event *o = &s->trace[s->trace_n];
    o->t = e->t;
    o->type = e->type;
    o->i = e->i;
    o->j = e->j;
    o->pad = 0;
  }
  if (s->trace_cap) s->trace_n++;
}

/* the eight process_event overloads, hardspheres.cc:1106-1645 */
static int process_event(orc_sim *s, const ev_t *e) {
  const unsigned nb = s->nb, i = e->i;
  if (e->type == HMC_EV_BEAD_CELL) {
    /* hardspheres.cc:1622-1645: no counter bump */
    if (s->counter[i] != e->ni) return 0;
    move_to_new_cell(s, i, e->cxn, e->cyn, e->czn);
    update_pos(s, i, e->t);
    scan_9(s, i, e->j);
    if_cell(s, i);
    s->n_cell++;
    trace_event(s, e);
    return 1;
  }
  const unsigned j = e->j;
  if (s->counter[i] != e->ni || s->counter[j] != e->nj) return 0;
  s->counter[i]++;
  s->counter[j]++;
  update_pos(s, i, e->t);
  update_pos(s, j, e->t);

  switch (e->type) {
  case HMC_EV_MIN_NEAREST: final_vel(s, s->near_min2, i, j, 0, 0.0); break;
  case HMC_EV_MAX_NEAREST: final_vel(s, s->near_max2, i, j, 0, 0.0); break;
  case HMC_EV_MIN_NNEAREST: final_vel(s, s->nnear_min2, i, j, 0, 0.0); break;
  case HMC_EV_MAX_NNEAREST: final_vel(s, s->nnear_max2, i, j, 0, 0.0); break;
  case HMC_EV_MIN_NONLOCAL_INNER: final_vel(s, s->rh2, i, j, 0, 0.0); break;
  case HMC_EV_MAX_NONLOCAL_INNER: {
    /* hardspheres.cc:1462-1477; s_of_inner_event :299-313 */
    const uint64_t tm = s->tmask[i][j], pm = s->pmask[i][j];
    const double r2 = rc2_inner(s, pm);
    const uint64_t nonbonded = s->config, bonded = s->config ^ tm;
    const double dS = s->s_bias[(unsigned)bonded] - s->s_bias[(unsigned)nonbonded];
    if (final_vel(s, r2, i, j, 1, dS)) {
      s->config ^= tm;
      s->formed++;
    }
    break;
  }
  case HMC_EV_MAX_NONLOCAL_OUTER: {
    /* hardspheres.cc:1548-1569; s_of_outer_event :317-331 */
    const uint64_t tm = s->tmask[i][j], pm = s->pmask[i][j];
    const double r2 = rc2_outer(s, tm, pm);
    int has_dS = 0;
    double dS = 0.0;
    if (tm && (s->config & tm)) {
      const uint64_t bonded = s->config, nonbonded = s->config ^ tm;
      dS = s->s_bias[(unsigned)nonbonded] - s->s_bias[(unsigned)bonded];
      has_dS = 1;
    }
    if (final_vel(s, r2, i, j, has_dS, dS)) {
      s->config ^= tm;
      s->broken++;
    }
    break;
  }
  default: break;
  }

  /* re-prediction of the local wells, in the reference's push order */
  if (e->type <= HMC_EV_MAX_NEAREST) { /* :1130-1158 */
    if (i > 1) if_nnear(s, i - 2, i);
    if (j > 1) if_nnear(s, j - 2, j);
    if (i > 0) if_near(s, i - 1, i);
    if_near(s, i, j);
    if (j + 1 < nb) if_near(s, j, j + 1);
    if (i + 2 < nb) if_nnear(s, i, i + 2);
    if (j + 2 < nb) if_nnear(s, j, j + 2);
  } else if (e->type <= HMC_EV_MAX_NNEAREST) { /* :1267-1291 */
    if (i > 1) if_nnear(s, i - 2, i);
    if (i > 0) if_near(s, i - 1, i);
    if_near(s, i, i + 1);
    if_nnear(s, i, j);
    if_near(s, j - 1, j);
    if (j + 1 < nb) if_near(s, j, j + 1);
    if (j + 2 < nb) if_nnear(s, j, j + 2);
  } else { /* :1393-1425 */
    const unsigned b2[2] = {i, j};
    for (int q = 0; q < 2; q++) {
      const unsigned b = b2[q];
      if (b > 1) if_nnear(s, b - 2, b);
      if (b > 0) if_near(s, b - 1, b);
      if (b + 1 < nb) if_near(s, b, b + 1);
      if (b + 2 < nb) if_nnear(s, b, b + 2);
    }
  }
  scan_27(s, i);
  scan_27(s, j);
  if_cell(s, i);
  if_cell(s, j);
  s->n_coll++;
  trace_event(s, e);
  return 1;
}

/* ------------------------------------------------------ constraint checks -- */

/* check_bond, hardspheres.cc:980-1004 (1e6*eps tolerance) */
static int check_bond(double min2, double max2, double dx, double dy, double dz) {
  const double tol = 1000000 * 2.220446049250313e-16;
  const double lo = min2 * (1 - tol), hi = max2 * (1 + tol);
  const double d2 = dx * dx + dy * dy + dz * dz;
  return (d2 > lo && d2 < hi) ? 1 : 0;
}
/* check_bond_if_crankshaft, crankshaft.cc:75-86 (no tolerance) */
static int check_bond_strict(double min2, double max2, double dx, double dy, double dz) {
  const double d2 = dx * dx + dy * dy + dz * dz;
  return (d2 > min2 && d2 < max2) ? 1 : 0;
}

/* check_local_dist, hardspheres.cc:1006-1037 / check_local_dist_if_crankshaft,
 * crankshaft.cc:88-123 */
static int check_local(const orc_sim *s, const double (*pos)[3], int strict) {
  for (unsigned i = 0; i + 1 < s->nb; i++) {
    double dx = pos[i + 1][0] - pos[i][0], dy = pos[i + 1][1] - pos[i][1],
           dz = pos[i + 1][2] - pos[i][2];
    mindist(s, &dx, &dy, &dz);
    if (!(strict ? check_bond_strict(s->near_min2, s->near_max2, dx, dy, dz)
                 : check_bond(s->near_min2, s->near_max2, dx, dy, dz)))
      return 0;
  }
  for (unsigned i = 0; i + 2 < s->nb; i++) {
    double dx = pos[i + 2][0] - pos[i][0], dy = pos[i + 2][1] - pos[i][1],
           dz = pos[i + 2][2] - pos[i][2];
    mindist(s, &dx, &dy, &dz);
    if (!(strict ? check_bond_strict(s->nnear_min2, s->nnear_max2, dx, dy, dz)
                 : check_bond(s->nnear_min2, s->nnear_max2, dx, dy, dz)))
      return 0;
  }
  return 1;
}

/* check_nonlocal_dist, crankshaft.cc:125-171 */
static int check_nonlocal(const orc_sim *s, const double (*pos)[3]) {
  for (unsigned i = 0; i + 3 < s->nb; i++) {
    for (unsigned j = i + 3; j < s->nb; j++) {
      double dx = pos[i][0] - pos[j][0], dy = pos[i][1] - pos[j][1],
             dz = pos[i][2] - pos[j][2];
      mindist(s, &dx, &dy, &dz);
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (!(d2 > s->rh2)) return 0;
      const uint64_t pm = s->pmask[i][j];
      if (pm && !(d2 < rc2_inner(s, pm))) return 0;
      const uint64_t tm = s->tmask[i][j];
      if (s->p.has_stair && tm && !(d2 < s->stair2)) return 0;
    }
  }
  return 1;
}

int orc_check_local_dist(const orc_sim *s) { return check_local(s, s->pos, 0); }
int orc_check_nonlocal_dist(const orc_sim *s) { return check_nonlocal(s, s->pos); }

/* ------------------------------------------------------------- velocities -- */

/* shift_vel_CM_to_0, hardspheres.cc:173-194 */
void orc_shift_vel_cm(double *vel, unsigned nb) {
  double tx = 0.0, ty = 0.0, tz = 0.0;
  for (unsigned i = 0; i < nb; i++) {
    tx += vel[3 * i + 0];
    ty += vel[3 * i + 1];
    tz += vel[3 * i + 2];
  }
  const double cx = tx / (double)nb, cy = ty / (double)nb, cz = tz / (double)nb;
  for (unsigned i = 0; i < nb; i++) {
    vel[3 * i + 0] -= cx;
    vel[3 * i + 1] -= cy;
    vel[3 * i + 2] -= cz;
  }
}

/* total_k_E + compute_hamiltonian, hardspheres.cc:1063-1082 */
static double hamiltonian(const orc_sim *s) {
  double ke = 0.0;
  for (unsigned i = 0; i < s->nb; i++) {
    const double vx2 = s->vel[i][0] * s->vel[i][0];
    const double vy2 = s->vel[i][1] * s->vel[i][1];
    const double vz2 = s->vel[i][2] * s->vel[i][2];
    ke += 0.5 * s->p.m * (vx2 + vy2 + vz2);
  }
  return ke + s->s_bias[s->config];
}

/* ------------------------------------------------------------------- steps -- */

/* initialize_system, main.cc:45-95 */
static int initialize_system(orc_sim *s, const double *normals) {
  int err = 0;
  if (!check_local(s, s->pos, 0)) err |= HMC_ERR_LOCAL_OVERLAP;
  if (!check_nonlocal(s, s->pos)) err |= HMC_ERR_NONLOCAL_OVERLAP;
  s->hn = 0;
  init_cells(s);
  memcpy(s->vel, normals, sizeof(double) * 3 * s->nb);
  orc_shift_vel_cm(&s->vel[0][0], s->nb);
  for (unsigned i = 0; i < s->nb; i++) if_cell(s, i);                 /* init_cell_events :696 */
  for (unsigned i = 0; i + 1 < s->nb; i++) if_near(s, i, i + 1);      /* :885 */
  for (unsigned i = 0; i + 2 < s->nb; i++) if_nnear(s, i, i + 2);     /* :926 */
  for (unsigned i = 0; i < s->nb; i++) scan_27(s, i);                 /* :540 */
  return err;
}

/* run_step, main.cc:97-149 */
static void run_step(orc_sim *s, double step_time) {
  while (s->hn) {
    if (s->heap[0].t > step_time) break;
    const ev_t e = heap_pop(s);
    s->n_pop++;
    process_event(s, &e);
  }
  for (unsigned i = 0; i < s->nb; i++) update_pos(s, i, step_time);
}

/* dist_between_nonlocal_beads, hardspheres.cc:1084-1104 */
static void sample_dist(orc_sim *s) {
  if (!s->dist_cap) return;
  if (s->dist_n >= s->dist_cap) {
    s->dist_n++;
    return;
  }
  double *row = s->dist + (size_t)s->dist_n * s->p.n_nonlocal;
  unsigned k = 0;
  for (unsigned i = 0; i + 3 < s->nb; i++)
    for (unsigned j = i + 3; j < s->nb; j++)
      if (s->nlmask[i][j]) {
        double dx = s->pos[i][0] - s->pos[j][0], dy = s->pos[i][1] - s->pos[j][1],
               dz = s->pos[i][2] - s->pos[j][2];
        mindist(s, &dx, &dy, &dz);
        if (k < s->p.n_nonlocal) row[k] = sqrt(dx * dx + dy * dy + dz * dz);
        k++;
      }
  s->dist_n++;
}

static void record_config_int(orc_sim *s) {
  if (s->cfg_cap) {
    if (s->cfg_n < s->cfg_cap) s->cfg_int[s->cfg_n] = s->config;
    s->cfg_n++;
  }
}

static double wl_gamma(const orc_sim *s, unsigned iter_wl) {
  /* main.cc:283,302-303 */
  return iter_wl == 0 ? s->p.gamma : 1.0 / (double)iter_wl;
}

int orc_md_step(orc_sim *s, int phase, unsigned step, const double *normals) {
  int err = initialize_system(s, normals);
  if (phase == HMC_PHASE_EQ) { /* run_trajectory_eq, main.cc:151-169 */
    run_step(s, step * s->p.del_t);
    return err;
  }
  const double e0 = hamiltonian(s);
  if (phase == HMC_PHASE_WL) { /* run_trajectory_wl + wang_landau, main.cc:249-305 */
    const unsigned iter_wl = step / s->nstates;
    for (unsigned st = iter_wl * s->p.nsteps_wl; st < (iter_wl + 1) * s->p.nsteps_wl; st++) {
      run_step(s, st * s->p.del_t_wl);
      const double ediff = fabs(1 - (hamiltonian(s) / e0));
      if (!(ediff < 1e-6)) err |= HMC_ERR_ENERGY;
    }
    const double g = wl_gamma(s, iter_wl);
    if (s->config == 0)
      s->s_bias[s->nstates - 1] -= g;
    else
      s->s_bias[s->nstates - 1] += g;
    return err;
  }
  /* run_trajectory body, main.cc:185-226 */
  run_step(s, step * s->p.del_t);
  sample_dist(s);
  if (step % s->p.write_step == 0) {
    if (s->config < s->nstates) s->config_count[s->config]++;
    record_config_int(s);
    if (s->config == 1 && !s->uniq_found) {
      s->uniq_found = 1;
      memcpy(s->uniq_pos, s->pos, sizeof s->uniq_pos);
      memcpy(s->uniq_vel, s->vel, sizeof s->uniq_vel);
    }
  }
  const double ediff = fabs(1 - (hamiltonian(s) / e0));
  if (!(ediff < 1e-6)) err |= HMC_ERR_ENERGY;
  return err;
}

/* ------------------------------------------------------------- crankshaft -- */

/* rodrigues_rotation, crankshaft.cc:11-73 (sin_t = sin(theta), cos_t = cos(theta)) */
static void rodrigues(const orc_sim *s, double sin_t, double cos_t, unsigned ind,
                      double (*trial)[3]) {
  const double (*pos)[3] = s->pos;
  double dx = pos[ind][0] - pos[ind - 1][0];
  double dy = pos[ind][1] - pos[ind - 1][1];
  double dz = pos[ind][2] - pos[ind - 1][2];
  mindist(s, &dx, &dy, &dz);
  const double inv = 1 / sqrt(dx * dx + dy * dy + dz * dz);
  const double kx = dx * inv, ky = dy * inv, kz = dz * inv;
  const double sn = sin_t, cs = 1 - cos_t;
  const double Rxx = 1 + cs * (-kz * kz - ky * ky);
  const double Rxy = -sn * kz + cs * ky * kx;
  const double Rxz = sn * ky + cs * kz * kx;
  const double Ryx = sn * kz + cs * kx * ky;
  const double Ryy = 1 + cs * (-kz * kz - kx * kx);
  const double Ryz = -sn * kx + cs * kz * ky;
  const double Rzx = -sn * ky + cs * kx * kz;
  const double Rzy = sn * kx + cs * ky * kz;
  const double Rzz = 1 + cs * (-ky * ky - kx * kx);
  for (unsigned i = 0; i <= ind; i++) {
    trial[i][0] = pos[i][0];
    trial[i][1] = pos[i][1];
    trial[i][2] = pos[i][2];
  }
  for (unsigned i = ind + 1; i < s->nb; i++) {
    double ox = pos[i][0] - pos[ind][0], oy = pos[i][1] - pos[ind][1],
           oz = pos[i][2] - pos[ind][2];
    mindist(s, &ox, &oy, &oz);
    double nx = Rxx * ox + Rxy * oy + Rxz * oz;
    double ny = Ryx * ox + Ryy * oy + Ryz * oz;
    double nz = Rzx * ox + Rzy * oy + Rzz * oz;
    mindist(s, &nx, &ny, &nz);
    trial[i][0] = pos[ind][0] + nx;
    trial[i][1] = pos[ind][1] + ny;
    trial[i][2] = pos[ind][2] + nz;
  }
}

void orc_rodrigues(const orc_sim *s, double sin_t, double cos_t, unsigned ind,
                   double *pos_trial) {
  rodrigues(s, sin_t, cos_t, ind, (double(*)[3])pos_trial);
}

/* config_int, crankshaft.cc:175-205 (all pairs i<j, dist2 <= rc2) */
static uint64_t config_from_pos(const orc_sim *s, const double (*pos)[3]) {
  uint64_t c = 0;
  for (unsigned i = 0; i < s->nb; i++)
    for (unsigned j = i + 1; j < s->nb; j++) {
      double dx = pos[i][0] - pos[j][0], dy = pos[i][1] - pos[j][1],
             dz = pos[i][2] - pos[j][2];
      mindist(s, &dx, &dy, &dz);
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (d2 > s->rc2) continue;
      c ^= s->tmask[i][j];
    }
  return c;
}
uint64_t orc_config_int(const orc_sim *s, const double *pos_trial) {
  return config_from_pos(s, (const double(*)[3])pos_trial);
}

/* libstdc++ generate_canonical<double,53> over a 32-bit engine
 * (bits/random.tcc:3349-3381): two words, low first. */
static inline double canonical_from_words(uint32_t w0, uint32_t w1) {
  double sum = 0.0, tmp = 1.0;
  sum += (double)w0 * tmp;
  tmp *= 4294967296.0;
  sum += (double)w1 * tmp;
  tmp *= 4294967296.0;
  double r = sum / tmp;
  if (r >= 1.0) r = nextafter(1.0, 0.0);
  return r;
}

typedef struct {
  const uint32_t *w;
  unsigned n, c;
  int underrun;
} tape_t;
static inline uint32_t tape_next(tape_t *t) {
  if (t->c >= t->n) {
    t->underrun = 1;
    t->c++;
    return 0;
  }
  return t->w[t->c++];
}

/* libstdc++ uniform_int_distribution<int>(a,b) on a 32-bit engine: Lemire's
 * method, bits/uniform_int_dist.h:255-281,305-335 */
static unsigned tape_uniform_int(tape_t *t, unsigned a, unsigned b) {
  const uint32_t range = b - a + 1;
  uint64_t prod = (uint64_t)tape_next(t) * (uint64_t)range;
  uint32_t low = (uint32_t)prod;
  if (low < range) {
    const uint32_t thr = (uint32_t)(-range) % range;
    while (low < thr && !t->underrun) {
      prod = (uint64_t)tape_next(t) * (uint64_t)range;
      low = (uint32_t)prod;
    }
  }
  return (unsigned)(prod >> 32) + a;
}

/* crankshaft, crankshaft.cc:224-262 + accept_move :207-222, driven by a tape */
int orc_crankshaft(orc_sim *s, unsigned move_offset, unsigned n_moves,
                   const uint32_t *words, const double *sin_tab, const double *cos_tab,
                   unsigned n_words, unsigned *cursor) {
  tape_t tp = {words, n_words, *cursor, 0};
  double trial[NBMAX][3];
  int err = 0;
  s->last_moves_done = 0;
  for (unsigned mv = 0; mv < n_moves; mv++) {
    const unsigned start = tp.c;
    unsigned attempts = 0;
    int ok = 0;
    do {
      const unsigned ind = tape_uniform_int(&tp, 1, s->nb - 2);
      const unsigned cth = tp.c; /* theta is formed from words cth, cth+1 */
      tape_next(&tp);
      tape_next(&tp);
      if (tp.underrun) break;
      rodrigues(s, sin_tab[cth], cos_tab[cth], ind, trial);
      attempts++;
      ok = check_local(s, (const double(*)[3])trial, 1) &&
           check_nonlocal(s, (const double(*)[3])trial);
    } while (!ok);
    if (tp.underrun) {
      tp.c = start;
      err |= HMC_ERR_TAPE_UNDERRUN;
      break;
    }
    const uint64_t trial_cfg = config_from_pos(s, (const double(*)[3])trial);
    const double dS = s->s_bias[trial_cfg] - s->s_bias[s->config];
    int accept = 1;
    if (dS > 0) {
      const uint32_t w0 = tape_next(&tp), w1 = tape_next(&tp);
      if (tp.underrun) {
        /* roll the whole move back; positions were not touched yet */
        tp.c = start;
        err |= HMC_ERR_TAPE_UNDERRUN;
        break;
      }
      const double u = canonical_from_words(w0, w1) * (1.0 - 0.0) + 0.0;
      if (u > exp(-dS)) accept = 0;
    }
    s->n_attempt += attempts;
    if (accept) {
      memcpy(s->pos, trial, sizeof(double) * 3 * s->nb);
      s->config = trial_cfg;
      s->n_accept++;
    }
    s->last_moves_done = mv + 1;
    /* main.cc:232-236 */
    const unsigned i = move_offset + mv;
    for (unsigned j = 1; j < s->p.mc_write; j++)
      if (i == (j * s->p.mc_moves) / s->p.mc_write) record_config_int(s);
  }
  *cursor = tp.c;
  return err;
}

/* -------------------------------------------------------------- lifecycle -- */

static uint64_t mask_of(const uint32_t (*bonds)[2], unsigned n, unsigned i, unsigned j) {
  for (unsigned k = 0; k < n; k++) {
    unsigned a = bonds[k][0], b = bonds[k][1];
    if (a > b) { /* NonlocalBonds ctor, config.cc:6-13 */
      const unsigned q = a;
      a = b;
      b = q;
    }
    if (a == i && b == j) return (uint64_t)1 << k;
  }
  return 0;
}

orc_sim *orc_create(const hmc_params *p) {
  orc_sim *s = (orc_sim *)calloc(1, sizeof *s);
  s->p = *p;
  s->nb = p->nbeads;
  s->ncell = p->ncell;
  s->nstates = 1u << p->n_transient;
  s->l = p->length;
  s->inv_l = 1.0 / p->length;
  s->lcell = p->length / p->ncell;
  s->inv_lcell = 1 / s->lcell;
  s->near_min2 = p->near_min * p->near_min;
  s->near_max2 = p->near_max * p->near_max;
  s->nnear_min2 = p->nnear_min * p->nnear_min;
  s->nnear_max2 = p->nnear_max * p->nnear_max;
  s->rh2 = p->rh * p->rh;
  s->rc2 = p->rc * p->rc;
  s->stair2 = p->has_stair ? p->stair * p->stair : 0.0;
  s->p_rc2 = p->has_p_rc ? p->p_rc * p->p_rc : 0.0;
  for (unsigned i = 0; i < s->nb; i++)
    for (unsigned j = i + 1; j < s->nb; j++) {
      s->tmask[i][j] = mask_of(p->transient_bonds, p->n_transient, i, j);
      s->pmask[i][j] = mask_of(p->permanent_bonds, p->n_permanent, i, j);
      s->nlmask[i][j] = mask_of(p->nonlocal_bonds, p->n_nonlocal, i, j);
    }
  const unsigned nc3 = s->ncell * s->ncell * s->ncell;
  s->cell_list = (uint8_t *)calloc((size_t)nc3 * s->nb, 1);
  s->cell_cnt = (uint8_t *)calloc(nc3, 1);
  s->s_bias = (double *)calloc(s->nstates, sizeof(double));
  s->config_count = (uint64_t *)calloc(s->nstates, sizeof(uint64_t));
  return s;
}

void orc_destroy(orc_sim *s) {
  if (!s) return;
  free(s->cell_list);
  free(s->cell_cnt);
  free(s->s_bias);
  free(s->config_count);
  free(s->heap);
  free(s->trace);
  free(s->dist);
  free(s->cfg_int);
  free(s);
}

void orc_set_pos(orc_sim *s, const double *pos) { memcpy(s->pos, pos, sizeof(double) * 3 * s->nb); }
void orc_get_pos(const orc_sim *s, double *pos) { memcpy(pos, s->pos, sizeof(double) * 3 * s->nb); }
void orc_get_vel(const orc_sim *s, double *vel) { memcpy(vel, s->vel, sizeof(double) * 3 * s->nb); }
void orc_set_config(orc_sim *s, uint64_t c) { s->config = c; }
uint64_t orc_get_config(const orc_sim *s) { return s->config; }

/* init_update_config, hardspheres.cc:144-171 */
void orc_init_config_from_pos(orc_sim *s) {
  s->config = 0;
  for (unsigned i = 0; i + 3 < s->nb; i++)
    for (unsigned j = i + 3; j < s->nb; j++) {
      double dx = s->pos[j][0] - s->pos[i][0], dy = s->pos[j][1] - s->pos[i][1],
             dz = s->pos[j][2] - s->pos[i][2];
      mindist(s, &dx, &dy, &dz);
      const double d2 = dx * dx + dy * dy + dz * dz;
      const uint64_t tm = s->tmask[i][j];
      if (tm && d2 < s->rc2) s->config ^= tm;
    }
}

void orc_set_s_bias(orc_sim *s, const double *sb, int n) {
  memcpy(s->s_bias, sb, sizeof(double) * ((unsigned)n < s->nstates ? (unsigned)n : s->nstates));
}
void orc_get_s_bias(const orc_sim *s, double *sb, int n) {
  memcpy(sb, s->s_bias, sizeof(double) * ((unsigned)n < s->nstates ? (unsigned)n : s->nstates));
}
/* init_s, entropy.cc:14-20 */
void orc_init_s(orc_sim *s) {
  for (unsigned i = 0; i < s->nstates; i++)
    s->s_bias[i] = 3 * (s->p.n_transient - (unsigned)__builtin_popcount(i));
}
void orc_reset_clocks(orc_sim *s) {
  for (unsigned i = 0; i < s->nb; i++) {
    s->times[i] = 0.0;
    s->counter[i] = 0;
  }
}
void orc_reset_counts(orc_sim *s) {
  memset(s->config_count, 0, sizeof(uint64_t) * s->nstates);
  s->dist_n = 0;
  s->cfg_n = 0;
}

void orc_enable_trace(orc_sim *s, unsigned cap) {
  free(s->trace);
  s->trace = cap ? (hmc_event *)calloc(cap, sizeof(hmc_event)) : NULL;
  s->trace_cap = cap;
  s->trace_n = 0;
}
unsigned orc_get_trace(const orc_sim *s, hmc_event *out, unsigned max_events) {
  unsigned n = s->trace_n < s->trace_cap ? s->trace_n : s->trace_cap;
  if (n > max_events) n = max_events;
  if (out) memcpy(out, s->trace, n * sizeof(hmc_event));
  return s->trace_n;
}
void orc_enable_samples(orc_sim *s, unsigned dist_rows, unsigned cfg_cap) {
  free(s->dist);
  free(s->cfg_int);
  s->dist = dist_rows ? (double *)calloc((size_t)dist_rows * (s->p.n_nonlocal ? s->p.n_nonlocal : 1), sizeof(double)) : NULL;
  s->cfg_int = cfg_cap ? (uint64_t *)calloc(cfg_cap, sizeof(uint64_t)) : NULL;
  s->dist_cap = dist_rows;
  s->cfg_cap = cfg_cap;
  s->dist_n = s->cfg_n = 0;
}
unsigned orc_get_dist(const orc_sim *s, double *dist) {
  const unsigned n = s->dist_n < s->dist_cap ? s->dist_n : s->dist_cap;
  if (dist) memcpy(dist, s->dist, sizeof(double) * (size_t)n * s->p.n_nonlocal);
  return s->dist_n;
}
unsigned orc_get_config_int(const orc_sim *s, uint64_t *out) {
  const unsigned n = s->cfg_n < s->cfg_cap ? s->cfg_n : s->cfg_cap;
  if (out) memcpy(out, s->cfg_int, sizeof(uint64_t) * n);
  return s->cfg_n;
}
void orc_get_counts(const orc_sim *s, uint64_t *cc, int nstates, uint64_t *formed,
                    uint64_t *broken) {
  if (cc) memcpy(cc, s->config_count, sizeof(uint64_t) * ((unsigned)nstates < s->nstates ? (unsigned)nstates : s->nstates));
  if (formed) *formed = s->formed;
  if (broken) *broken = s->broken;
}
void orc_get_event_counts(const orc_sim *s, uint64_t ev[6]) {
  ev[0] = s->n_coll;
  ev[1] = s->n_cell;
  ev[2] = s->n_attempt;
  ev[3] = s->n_accept;
  ev[4] = s->n_pop;
  ev[5] = s->n_pred;
}
unsigned orc_last_moves_done(const orc_sim *s) { return s->last_moves_done; }
int orc_get_unique_config(const orc_sim *s, double *pos, double *vel) {
  if (s->uniq_found) {
    if (pos) memcpy(pos, s->uniq_pos, sizeof(double) * 3 * s->nb);
    if (vel) memcpy(vel, s->uniq_vel, sizeof(double) * 3 * s->nb);
  }
  return s->uniq_found;
}

/* ---------------------------------------------------------------- entropy -- */

/* compute_norm, entropy.cc:24-34 */
static double compute_norm(const uint64_t *cc, int n) {
  double norm = 0.0;
  for (int i = 0; i < n; i++) norm += cc[i];
  norm /= (double)n;
  return norm;
}

/* compute_entropy, entropy.cc:36-69 */
void orc_compute_entropy(const uint64_t *cc, double *sb, int n, double pos_scale,
                         double neg_scale) {
  const double norm = compute_norm(cc, n);
  const int nat = n - 1;
  const double native_count = cc[nat];
  double s_native;
  if (native_count > 0)
    s_native = sb[nat] + pos_scale * log(native_count / norm);
  else
    s_native = sb[nat] - neg_scale * log(norm);
  for (int i = 0; i < n; i++) {
    double si;
    if (cc[i] > 0)
      si = sb[i] + pos_scale * log(cc[i] / norm);
    else
      si = sb[i] - neg_scale * log(norm);
    sb[i] = si - s_native;
  }
}

/* compute_g_val, entropy.cc:72-89 */
double orc_compute_g_val(const uint64_t *cc, int n) {
  const double norm = compute_norm(cc, n);
  double g = 0.0;
  for (int i = 0; i < n; i++)
    if (cc[i] > 0) g += cc[i] * log(cc[i] / norm);
  g *= 2.0;
  return g;
}

/* chi-squared cdf: regularised lower incomplete gamma P(a,x) — published series /
 * Lentz continued fraction (Boost.Math is the reference's dependency here,
 * entropy.cc:11,98-104; unpinned version, restated from the textbook forms). */
static double gamma_p(double a, double x) {
  if (x <= 0) return 0.0;
  const double gln = lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int k = 0; k < 100000; k++) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    return sum * exp(-x + a * log(x) - gln);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 100000; i++) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 1e-16) break;
  }
  return 1.0 - exp(-x + a * log(x) - gln) * h;
}

/* g_test, entropy.cc:92-116 */
int orc_g_test(const uint64_t *cc, int n, double sig, double *g_val, double *g_crit,
               double *p_val) {
  const double dof = n - 1;
  const double g = orc_compute_g_val(cc, n);
  const double pv = 1.0 - gamma_p(0.5 * dof, 0.5 * g);
  double lo = 0.0, hi = 1.0;
  while (1.0 - gamma_p(0.5 * dof, 0.5 * hi) > sig) hi *= 2.0;
  for (int k = 0; k < 200; k++) {
    const double mid = 0.5 * (lo + hi);
    if (1.0 - gamma_p(0.5 * dof, 0.5 * mid) > sig)
      lo = mid;
    else
      hi = mid;
  }
  const double crit = 0.5 * (lo + hi);
  if (g_val) *g_val = g;
  if (g_crit) *g_crit = crit;
  if (p_val) *p_val = pv;
  return (g < crit) ? 1 : 0;
}

/* ------------------------------------------------------ std::mt19937 et al -- */

/* std::seed_seq::generate ([rand.util.seedseq]) for n = 624 words, followed by
 * mersenne_twister_engine::seed(Sseq&) (all-zero guard). */
void orc_mt_seed_seq(orc_mt19937 *g, const uint32_t *v, unsigned s) {
  const unsigned n = 624, t = 11, p = (n - t) / 2, q = p + t;
  const unsigned m = (s + 1 > n) ? s + 1 : n;
  uint32_t *b = g->mt;
  for (unsigned k = 0; k < n; k++) b[k] = 0x8b8b8b8bu;
  for (unsigned k = 0; k < m; k++) {
    uint32_t a = b[k % n] ^ b[(k + p) % n] ^ b[(k + n - 1) % n];
    a ^= a >> 27;
    const uint32_t r1 = 1664525u * a;
    uint32_t r2 = r1;
    if (k == 0)
      r2 += s;
    else if (k <= s)
      r2 += (k % n) + v[k - 1];
    else
      r2 += (k % n);
    b[(k + p) % n] += r1;
    b[(k + q) % n] += r2;
    b[k % n] = r2;
  }
  for (unsigned k = m; k < m + n; k++) {
    uint32_t a = b[k % n] + b[(k + p) % n] + b[(k + n - 1) % n];
    a ^= a >> 27;
    const uint32_t r3 = 1566083941u * a;
    const uint32_t r4 = r3 - (k % n);
    b[(k + p) % n] ^= r3;
    b[(k + q) % n] ^= r4;
    b[k % n] = r4;
  }
  int zero = (b[0] & 0x80000000u) == 0;
  for (unsigned k = 1; zero && k < n; k++)
    if (b[k] != 0) zero = 0;
  if (zero) b[0] = 0x80000000u;
  g->idx = 624;
}

void orc_mt_seed(orc_mt19937 *g, uint32_t seed) {
  g->mt[0] = seed;
  for (unsigned i = 1; i < 624; i++)
    g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + i;
  g->idx = 624;
}

uint32_t orc_mt_next(orc_mt19937 *g) {
  if (g->idx >= 624) {
    uint32_t *mt = g->mt;
    for (unsigned k = 0; k < 624; k++) {
      const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
      mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

void orc_mt_discard(orc_mt19937 *g, unsigned n) {
  while (n--) orc_mt_next(g);
}

double orc_mt_canonical(orc_mt19937 *g) {
  const uint32_t w0 = orc_mt_next(g);
  const uint32_t w1 = orc_mt_next(g);
  return canonical_from_words(w0, w1);
}

/* std::normal_distribution<double>::operator() (Marsaglia polar,
 * bits/random.tcc:1811-1844), fresh distribution object as in init_vel
 * (hardspheres.cc:203) */
void orc_mt_normals(orc_mt19937 *g, double sigma, double *out, unsigned n) {
  int saved_ok = 0;
  double saved = 0.0;
  for (unsigned k = 0; k < n; k++) {
    double ret;
    if (saved_ok) {
      saved_ok = 0;
      ret = saved;
    } else {
      double x, y, r2;
      do {
        x = 2.0 * orc_mt_canonical(g) - 1.0;
        y = 2.0 * orc_mt_canonical(g) - 1.0;
        r2 = x * x + y * y;
      } while (r2 > 1.0 || r2 == 0.0);
      const double mult = sqrt(-2 * log(r2) / r2);
      saved = x * mult;
      saved_ok = 1;
      ret = y * mult;
    }
    out[k] = ret * sigma + 0.0;
  }
}

void orc_mt_tape(const orc_mt19937 *g, uint32_t *words, double *sin_tab, double *cos_tab,
                 unsigned n) {
  orc_mt19937 c = *g;
  for (unsigned k = 0; k < n; k++) words[k] = orc_mt_next(&c);
  for (unsigned k = 0; k < n; k++) {
    if (k + 1 < n) {
      /* std::uniform_real_distribution<>(0, 2*M_PI): canonical*(b-a)+a */
      const double th = canonical_from_words(words[k], words[k + 1]) * (2 * M_PI - 0.0) + 0.0;
      sin_tab[k] = sin(th);
      cos_tab[k] = cos(th);
    } else {
      sin_tab[k] = 0.0;
      cos_tab[k] = 1.0;
    }
  }
}

/* unit_sphere, hardspheres.cc:19-34 */
static void unit_sphere(orc_mt19937 *g, double *x, double *y, double *z) {
  double v, w, s;
  do {
    v = orc_mt_canonical(g) * (1.0 - -1.0) + -1.0;
    w = orc_mt_canonical(g) * (1.0 - -1.0) + -1.0;
    s = v * v + w * w;
  } while (!(s < 1));
  *x = 2 * v * sqrt(1 - v * v - w * w);
  *y = 2 * w * sqrt(1 - v * v - w * w);
  *z = 1 - 2 * (v * v + w * w);
}

/* init_pos, hardspheres.cc:37-138 */
int orc_init_pos(const hmc_params *p, orc_mt19937 *g, double *pos_out) {
  orc_sim *s = orc_create(p);
  double (*pos)[3] = (double(*)[3])pos_out;
  const unsigned nb = p->nbeads;
  int rc = 0;
  if (nb < 1) {
    orc_destroy(s);
    return 0;
  }
  pos[0][0] = pos[0][1] = pos[0][2] = 0.0;
  uint64_t tries = 0;
  unsigned i = 1;
  const double min3 = pow(p->near_min, 3.0), max3 = pow(p->near_max, 3.0);
  while (i < nb) {
    if (!(tries++ < p->tries)) {
      rc = -1;
      break;
    }
    double ux, uy, uz;
    unit_sphere(g, &ux, &uy, &uz);
    const double val = orc_mt_canonical(g) * (1.0 - 0.0) + 0.0;
    const double near = cbrt(min3 + (val * (max3 - min3)));
    const double x = near * ux + pos[i - 1][0];
    const double y = near * uy + pos[i - 1][1];
    const double z = near * uz + pos[i - 1][2];
    if (i >= 2) {
      double dx = x - pos[i - 2][0], dy = y - pos[i - 2][1], dz = z - pos[i - 2][2];
      mindist(s, &dx, &dy, &dz);
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (d2 < s->nnear_min2 || d2 > s->nnear_max2) continue;
    }
    int overlap = 0;
    for (unsigned j = 0; j < i; j++) {
      double dx = x - pos[j][0], dy = y - pos[j][1], dz = z - pos[j][2];
      mindist(s, &dx, &dy, &dz);
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (i >= j + 3 && d2 < s->rh2) {
        overlap = 1;
        break;
      }
      if (s->tmask[j][i] && d2 < s->rc2) {
        overlap = 1;
        break;
      }
    }
    if (overlap) continue;
    pos[i][0] = x;
    pos[i][1] = y;
    pos[i][2] = z;
    i++;
    tries = 0;
  }
  orc_destroy(s);
  return rc;
}

/* ------------------------------------------------ whole program (main.cc) -- */

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int orc_run_main(const hmc_params *p, const double *pos_in, unsigned max_gtest_iters,
                 unsigned dist_rows_cap, double *dist_out, unsigned *n_dist_rows,
                 orc_main_result *res) {
  memset(res, 0, sizeof *res);
  orc_sim *s = orc_create(p);
  const unsigned nb = p->nbeads, ns = s->nstates;
  if (ns > 16) {
    orc_destroy(s);
    return -1;
  }
  orc_mt19937 g;
  orc_mt_seed_seq(&g, p->seeds, p->n_seeds); /* main.cc:379-380 */
  const double vsig = sqrt(p->temp / p->m);
  double *normals = (double *)malloc(sizeof(double) * 3 * nb);
  int err = 0;

  orc_reset_clocks(s);
  /* initialize_pos, main.cc:26-43 */
  if (pos_in) {
    orc_set_pos(s, pos_in);
    orc_init_config_from_pos(s);
  } else if (orc_init_pos(p, &g, &s->pos[0][0]) != 0) {
    err = -2;
  }
  orc_init_s(s);
  const double t0 = now_s();
  /* equilibration, main.cc:422-431 (update_config_eq starts from the value
   * initialize_pos left: 0 for init_pos, computed for input files) */
  for (unsigned iter = 0; !err && iter < p->total_iter_eq; iter++)
    for (unsigned step = iter * p->nsteps_eq; step < (iter + 1) * p->nsteps_eq; step++) {
      orc_mt_normals(&g, vsig, normals, 3 * nb);
      err |= orc_md_step(s, HMC_PHASE_EQ, step, normals);
    }
  /* Wang-Landau, main.cc:434-441 */
  orc_init_config_from_pos(s);
  orc_reset_clocks(s);
  {
    double gamma = p->gamma;
    unsigned iter_wl = 0;
    while (!err && gamma > p->gamma_f) {
      for (unsigned i = 0; i < ns; i++) {
        orc_mt_normals(&g, vsig, normals, 3 * nb);
        err |= orc_md_step(s, HMC_PHASE_WL, iter_wl * ns + i, normals);
      }
      iter_wl += 1;
      gamma = 1.0 / (double)iter_wl;
    }
  }
  /* the G-test loop, main.cc:451-491 */
  if (dist_rows_cap) orc_enable_samples(s, dist_rows_cap, 0);
  const unsigned cap = max_gtest_iters ? max_gtest_iters : p->max_g_test_count;
  const unsigned tape_n = 2048;
  uint32_t *words = (uint32_t *)malloc(sizeof(uint32_t) * tape_n);
  double *st = (double *)malloc(sizeof(double) * tape_n);
  double *ct = (double *)malloc(sizeof(double) * tape_n);
  unsigned gcount = 0;
  int accepted = 0;
  do {
    orc_reset_clocks(s);
    memset(s->config_count, 0, sizeof(uint64_t) * ns);
    for (unsigned iter = 0; !err && iter < p->total_iter; iter++) {
      for (unsigned step = iter * p->nsteps; step < (iter + 1) * p->nsteps; step++) {
        orc_mt_normals(&g, vsig, normals, 3 * nb);
        err |= orc_md_step(s, HMC_PHASE_MAIN, step, normals);
      }
      unsigned done = 0;
      while (!err && done < p->mc_moves) {
        /* tape = the next words of the engine; refilled from the roll-back
         * point if a move runs past its end */
        orc_mt_tape(&g, words, st, ct, tape_n);
        unsigned cursor = 0;
        const int e2 = orc_crankshaft(s, done, p->mc_moves - done, words, st, ct, tape_n, &cursor);
        if ((e2 & HMC_ERR_TAPE_UNDERRUN) && s->last_moves_done == 0) {
          err |= e2; /* a single move needs more than the whole tape */
          break;
        }
        orc_mt_discard(&g, cursor);
        done += s->last_moves_done;
      }
    }
    orc_compute_entropy(s->config_count, s->s_bias, (int)ns, p->pos_scale, p->neg_scale);
    gcount++;
    accepted = orc_g_test(s->config_count, (int)ns, p->sig_level, NULL, NULL, NULL);
  } while (!err && gcount < cap && !accepted);
  res->seconds_md = now_s() - t0;
  for (unsigned i = 0; i < ns; i++) {
    res->s_bias[i] = s->s_bias[i];
    res->config_count[i] = s->config_count[i];
  }
  res->collisions = s->n_coll;
  res->cell_crossings = s->n_cell;
  res->heap_pops = s->n_pop;
  res->predictions = s->n_pred;
  res->gtest_iters = gcount;
  res->accepted = accepted;
  res->err = err;
  if (dist_out && n_dist_rows) *n_dist_rows = orc_get_dist(s, dist_out);
  free(words);
  free(st);
  free(ct);
  free(normals);
  orc_destroy(s);
  return err ? -1 : 0;
}

/* The oracle's replay of oracle/ref_driver.cc:ref_bench_main_phase (bench.py's CPU leg):
 * the same call sequence, bit-identical trajectory, and the event counters the reference
 * does not have.  collisions / cell_crossings count the timed part only. */
int orc_bench_main_phase(const hmc_params *p, const double *pos_in, unsigned eq_iters,
                         unsigned rounds, unsigned iters_per_round, double *seconds,
                         double *pos_final, uint64_t *config_final, uint64_t *collisions,
                         uint64_t *cell_crossings) {
  orc_sim *s = orc_create(p);
  const unsigned nb = p->nbeads, ns = s->nstates;
  orc_mt19937 g;
  orc_mt_seed_seq(&g, p->seeds, p->n_seeds);
  const double vsig = sqrt(p->temp / p->m);
  double *normals = (double *)malloc(sizeof(double) * 3 * nb);
  int err = 0;
  orc_reset_clocks(s);
  orc_set_pos(s, pos_in);
  orc_init_config_from_pos(s);
  orc_init_s(s);
  for (unsigned iter = 0; !err && iter < eq_iters; iter++)
    for (unsigned step = iter * p->nsteps_eq; step < (iter + 1) * p->nsteps_eq; step++) {
      orc_mt_normals(&g, vsig, normals, 3 * nb);
      err |= orc_md_step(s, HMC_PHASE_EQ, step, normals);
    }
  orc_init_config_from_pos(s);
  const unsigned tape_n = 2048;
  uint32_t *words = (uint32_t *)malloc(sizeof(uint32_t) * tape_n);
  double *st = (double *)malloc(sizeof(double) * tape_n);
  double *ct = (double *)malloc(sizeof(double) * tape_n);
  const uint64_t c0 = s->n_coll, x0 = s->n_cell;
  const double t0 = now_s();
  for (unsigned r = 0; !err && r < rounds; r++) {
    orc_reset_clocks(s);
    memset(s->config_count, 0, sizeof(uint64_t) * ns);
    for (unsigned iter = 0; !err && iter < iters_per_round; iter++) {
      for (unsigned step = iter * p->nsteps; step < (iter + 1) * p->nsteps; step++) {
        orc_mt_normals(&g, vsig, normals, 3 * nb);
        err |= orc_md_step(s, HMC_PHASE_MAIN, step, normals);
      }
      unsigned done = 0;
      while (!err && done < p->mc_moves) {
        orc_mt_tape(&g, words, st, ct, tape_n);
        unsigned cursor = 0;
        const int e2 = orc_crankshaft(s, done, p->mc_moves - done, words, st, ct, tape_n, &cursor);
        if ((e2 & HMC_ERR_TAPE_UNDERRUN) && s->last_moves_done == 0) {
          err |= e2;
          break;
        }
        orc_mt_discard(&g, cursor);
        done += s->last_moves_done;
      }
    }
    orc_compute_entropy(s->config_count, s->s_bias, (int)ns, p->pos_scale, p->neg_scale);
  }
  *seconds = now_s() - t0;
  *collisions = s->n_coll - c0;
  *cell_crossings = s->n_cell - x0;
  memcpy(pos_final, s->pos, sizeof(double) * 3 * nb);
  *config_final = orc_get_config(s);
  free(words);
  free(st);
  free(ct);
  free(normals);
  orc_destroy(s);
  return err ? -1 : 0;
}
