"""CPU: the partner pre-filter of the CUDA kernel (hybridmc_b200/csrc/hmc_kernels.cu, pass 1
of the partner re-prediction) restated in numpy float32 and checked against the reference's
double-precision predicate (if_coll, hardspheres.cc:425-476: approaching and
rv^2 > vv (rr - rh^2)) on states produced by the oracle: a pair the filter drops must never
have an event in exact arithmetic, and the filter must drop most pairs (otherwise it buys
nothing).  The GPU-side proof is the verification build (profiles/verify_prefilter.py)."""
import numpy as np
import pytest

from common import OracleRun, transition_params

f32 = np.float32


def prefilter_rejects(dx0, w, u, dt, L, rh2, chain_max):
    """float32 restatement of the kernel's test for one pair; dx0 = x_k(t_k) - x_b(t)"""
    fl, invl, frh2 = f32(L), f32(1.0 / L), f32(rh2)
    xc = f32(f32(chain_max) + fl)
    d0, w, u, dtf = dx0.astype(f32), w.astype(f32), u.astype(f32), f32(dt)
    W, U = f32(np.abs(w).sum(dtype=f32)), f32(np.abs(u).sum(dtype=f32))
    X = f32(f32(W + W) * dtf + xc)
    XV = f32(X * f32(W + U))
    dx = (w * dtf + d0).astype(f32)
    dx = (dx - np.rint(dx * invl).astype(f32) * fl).astype(f32)
    ex = (w - u).astype(f32)
    rv, vv = f32(dx @ ex), f32(ex @ ex)
    gap = f32(f32(dx @ dx) - frh2)
    far = gap > f32(1e-5) * X * X
    receding = rv > f32(1e-5) * XV
    xvs = f32(f32(1e-2) * XV)
    misses = f32(rv * rv + xvs * xvs) < f32(vv * gap)
    return bool(far and (receding or misses) and dtf >= 0)


@pytest.mark.parametrize("name", ["test", "crambin"])
def test_dropped_pairs_never_have_an_event(name):
    p = transition_params(name, 0)
    o = OracleRun(p)
    o.init_pos()
    o.sim.init_s()
    o.sim.init_config_from_pos()
    o.sim.reset_clocks()
    L, rh2, nb = p.length, p.rh ** 2, p.nbeads
    chain_max = (nb - 1) * p.near_max
    rng = np.random.default_rng(1)
    tested = dropped = 0
    for step in range(60 if name == "test" else 25):
        o.sim.md_step(0, step, o.normals(1)[0])
        P, V = o.sim.get_pos(), o.sim.get_vel()
        for _ in range(6):
            b = int(rng.integers(0, nb))
            ub = V[b] + rng.normal(size=3) * 0.5          # the event bead just changed velocity
            dt = rng.uniform(0, p.del_t, nb) * (rng.random(nb) < 0.7)  # partners' clocks lag
            Xk = P - V * dt[:, None]
            for k in range(nb):
                if abs(k - b) < 3:
                    continue
                # exact predicate, operand order of delta_tpv (hardspheres.cc:372-392)
                dx = Xk[k] + V[k] * dt[k] - P[b]
                dx = dx - np.round(dx / L) * L
                if np.any(np.abs(dx) > 2 * L / p.ncell):
                    continue  # not in adjacent cells: the kernel never tests such a pair
                dv = V[k] - ub
                rvv, vv, rr = dx @ dv, dv @ dv, dx @ dx
                event = rvv < 0 and rvv * rvv > vv * (rr - rh2)
                rej = prefilter_rejects(Xk[k] - P[b], V[k], ub, dt[k], L, rh2, chain_max)
                tested += 1
                dropped += rej
                assert not (rej and event), (b, k, dt[k], rvv, vv, rr)
    assert tested > 2000 and dropped > 0.5 * tested, (tested, dropped)
