"""Algorithm-independent cross-checks of the oracle and of the model loader. TEST INFRASTRUCTURE ONLY.

The oracle (abr_oracle.cc) and the engine both consume the loader's flattened model and both follow
MJX's algorithms (CRBA + RNE + Cholesky, Newton with MJX's line search). A shared mistake in either
place is invisible to oracle-vs-engine parity tests, so this module restates the same physics along
DIFFERENT routes, sharing nothing but the raw tree (masses, inertias, frames, joint axes):

* `setconst_numeric`   mj_setConst's constants (`dof_invweight0`, `body_invweight0`, `stat.meaninertia`,
                       `body_subtreemass`) from NUMERICALLY differentiated kinematics and a
                       kinetic-energy mass matrix, instead of the loader's analytic Jacobians;
* `aba`                forward dynamics by Featherstone's articulated-body algorithm (O(n), no mass matrix),
                       against the oracle's CRBA + RNE + L'DL `qacc_smooth`;
* `constraint_minimum` the constrained acceleration as the minimiser of the primal convex cost (App. A.8) found by a
                       generic quasi-Newton method from scipy, against the oracle's converged Newton solve.

Everything is float64 numpy on one world; none of it is on any product path.
"""
from __future__ import annotations

import numpy as np

JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3


# ----------------------------------------------------------------------------- small rotation helpers (own copies)
def _qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def _qmat(q):
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q)
    w, v = q[0], q[1:]
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + 2 * w * K + 2 * K @ K  # Rodrigues form (the loader uses the element-wise formula)


def _axis_angle(axis, ang):
    return np.concatenate([[np.cos(ang / 2)], np.asarray(axis) * np.sin(ang / 2)])


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


# ----------------------------------------------------------------------------- plain kinematics from the raw tree
def kinematics(m, qpos):
    """World poses of every body frame and inertial frame (App. A.2), straight from the raw tree."""
    nb = m.nbody
    xpos, xrot = np.zeros((nb, 3)), np.tile(np.eye(3), (nb, 1, 1))
    xquat = np.zeros((nb, 4))
    xquat[0, 0] = 1
    for b in range(1, nb):
        p = m.body_parentid[b]
        pos = xpos[p] + xrot[p] @ m.body_pos[b]
        quat = _qmul(xquat[p], m.body_quat[b])
        for j in range(m.body_jntadr[b], m.body_jntadr[b] + m.body_jntnum[b]) if m.body_jntnum[b] else ():
            a = m.jnt_qposadr[j]
            t = m.jnt_type[j]
            if t == JNT_FREE:
                pos = np.array(qpos[a:a + 3], dtype=np.float64)
                quat = np.array(qpos[a + 3:a + 7], dtype=np.float64)
                quat /= np.linalg.norm(quat)
            elif t == JNT_HINGE:
                R = _qmat(quat)
                anchor = pos + R @ m.jnt_pos[j]
                quat = _qmul(quat, _axis_angle(m.jnt_axis[j], qpos[a] - m.qpos0[a]))
                pos = anchor - _qmat(quat) @ m.jnt_pos[j]
            elif t == JNT_SLIDE:
                pos = pos + _qmat(quat) @ m.jnt_axis[j] * (qpos[a] - m.qpos0[a])
            else:
                raise NotImplementedError
        xpos[b], xquat[b], xrot[b] = pos, quat, _qmat(quat)
    xipos = np.array([xpos[b] + xrot[b] @ m.body_ipos[b] for b in range(nb)])
    xirot = np.array([xrot[b] @ _qmat(m.body_iquat[b]) for b in range(nb)])
    return xpos, xrot, xipos, xirot


def integrate_pos(m, qpos, dq):
    """qpos displaced by the tangent vector dq (free joint: world linear, body-local angular; App. A.9)."""
    q = np.array(qpos, dtype=np.float64)
    for j in range(m.njnt):
        a, d, t = m.jnt_qposadr[j], m.jnt_dofadr[j], m.jnt_type[j]
        if t == JNT_FREE:
            q[a:a + 3] += dq[d:d + 3]
            w = dq[d + 3:d + 6]
            n = np.linalg.norm(w)
            if n > 0:
                q[a + 3:a + 7] = _qmul(q[a + 3:a + 7], _axis_angle(w / n, n))
                q[a + 3:a + 7] /= np.linalg.norm(q[a + 3:a + 7])
        else:
            q[a] += dq[d]
    return q


def numeric_body_jacobians(m, qpos, eps=1e-6):
    """6 x nv Jacobian (translation of the inertial-frame origin, then rotation) of every body by central differences."""
    nb, nv = m.nbody, m.nv
    J = np.zeros((nb, 6, nv))
    for i in range(nv):
        e = np.zeros(nv)
        e[i] = eps
        _, rp, ip, _ = kinematics(m, integrate_pos(m, qpos, e))
        _, rm, im, _ = kinematics(m, integrate_pos(m, qpos, -e))
        J[:, :3, i] = (ip - im) / (2 * eps)
        for b in range(nb):
            dR = (rp[b] - rm[b]) / (2 * eps)
            W = dR @ (0.5 * (rp[b] + rm[b])).T  # [w]x = dR R'
            J[b, 3:, i] = [W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]]
            J[b, 3:, i] *= 0.5
    return J


def mass_matrix_energy(m, qpos, J=None):
    """M from the kinetic energy T = 1/2 sum_b (m v'v + w' I w): M = sum_b J_b' diag(m, I_b) J_b + armature."""
    J = numeric_body_jacobians(m, qpos) if J is None else J
    _, _, _, xirot = kinematics(m, qpos)
    M = np.diag(np.asarray(m.dof_armature, dtype=np.float64)).copy()
    for b in range(1, m.nbody):
        Iw = xirot[b] @ np.diag(m.body_inertia[b]) @ xirot[b].T
        M += m.body_mass[b] * J[b, :3].T @ J[b, :3] + J[b, 3:].T @ Iw @ J[b, 3:]
    return M


def setconst_numeric(m):
    """mj_setConst's constants at qpos0 by the numeric route. Returns a dict with the loader's field names."""
    J = numeric_body_jacobians(m, m.qpos0)
    M = mass_matrix_energy(m, m.qpos0, J)
    # unit responses M x = e_i by conjugate-direction-free plain elimination (the loader calls numpy's inverse)
    Minv = np.linalg.solve(M, np.eye(m.nv)) if m.nv else np.zeros((0, 0))
    dinv = np.diag(Minv).copy()
    for j in range(m.njnt):
        if m.jnt_type[j] == JNT_FREE:
            d = m.jnt_dofadr[j]
            dinv[d:d + 3] = dinv[d:d + 3].mean()
            dinv[d + 3:d + 6] = dinv[d + 3:d + 6].mean()
    biw = np.zeros((m.nbody, 2))
    moving = np.zeros(m.nbody, dtype=bool)
    for b in range(1, m.nbody):
        moving[b] = moving[m.body_parentid[b]] or m.body_jntnum[b] > 0
        if moving[b]:
            A = J[b] @ Minv @ J[b].T
            biw[b] = [max(1e-15, np.trace(A[:3, :3]) / 3), max(1e-15, np.trace(A[3:, 3:]) / 3)]
    sub = np.asarray(m.body_mass, dtype=np.float64).copy()
    for b in range(m.nbody - 1, 0, -1):
        sub[m.body_parentid[b]] += sub[b]
    return dict(qM0=M, dof_invweight0=dinv, body_invweight0=biw, meaninertia=float(np.mean(np.diag(M))) if m.nv else 1.0,
                body_subtreemass=sub)


# ----------------------------------------------------------------------------- Featherstone ABA (world coordinates)
def aba(m, qpos, qvel, tau, gravity):
    """qacc with M(q) qacc + c(q, v) = tau by the articulated-body algorithm. Spatial vectors are (angular, linear)
    about the WORLD origin; armature adds to the joint-space articulated inertia (rotor model of MuJoCo)."""
    nb, nv = m.nbody, m.nv
    xpos, xrot, xipos, xirot = kinematics(m, qpos)

    def xf_inertia(b):
        Ic = xirot[b] @ np.diag(m.body_inertia[b]) @ xirot[b].T
        c, ms = xipos[b], m.body_mass[b]
        C = _skew(c)
        I6 = np.zeros((6, 6))
        I6[:3, :3] = Ic - ms * C @ C
        I6[:3, 3:] = ms * C
        I6[3:, :3] = -ms * C
        I6[3:, 3:] = ms * np.eye(3)
        return I6

    def crm(v):  # motion cross product
        X = np.zeros((6, 6))
        X[:3, :3] = _skew(v[:3]); X[3:, :3] = _skew(v[3:]); X[3:, 3:] = _skew(v[:3])
        return X

    # joint motion subspaces in world coordinates, per dof
    S = np.zeros((nv, 6))
    for j in range(m.njnt):
        b, d, t = m.jnt_bodyid[j], m.jnt_dofadr[j], m.jnt_type[j]
        if t == JNT_FREE:
            for k in range(3):
                S[d + k, 3 + k] = 1.0  # world-frame translation
                ax = xrot[b][:, k]     # body-local rotation axes through the body origin
                S[d + 3 + k, :3] = ax
                S[d + 3 + k, 3:] = np.cross(xpos[b], ax)
        else:
            # the joint frame is the body frame BEFORE this joint's own motion for the anchor; for a single joint per
            # body the anchor is fixed in the parent, so use the parent-side construction
            p = m.body_parentid[b]
            Rp = xrot[p] @ _qmat(m.body_quat[b])
            anchor = xpos[p] + xrot[p] @ m.body_pos[b] + Rp @ m.jnt_pos[j]
            ax = Rp @ m.jnt_axis[j]
            if t == JNT_HINGE:
                S[d, :3] = ax
                S[d, 3:] = np.cross(anchor, ax)
            else:
                S[d, 3:] = ax
    if any(m.body_jntnum[b] > 1 for b in range(nb)):
        raise NotImplementedError("aba cross-check: one joint per body")
    dofs_of = [list(range(m.body_dofadr[b], m.body_dofadr[b] + m.body_dofnum[b])) if m.body_dofnum[b] else [] for b in range(nb)]
    v = np.zeros((nb, 6))
    cbias = np.zeros((nb, 6))
    IA = np.zeros((nb, 6, 6))
    pA = np.zeros((nb, 6))
    for b in range(1, nb):
        p = m.body_parentid[b]
        v[b] = v[p]
        ds = dofs_of[b]
        if ds and m.jnt_type[m.dof_jntid[ds[0]]] == JNT_FREE:
            # translations along world-fixed axes carry no velocity-product term; the three rotation axes are fixed in the
            # body itself, so their derivative is (body velocity) x S_k, and the angular-angular part cancels in the sum
            v[b] = v[b] + S[ds[:3]].T @ np.asarray(qvel)[ds[:3]]
            srot = S[ds[3:]].T @ np.asarray(qvel)[ds[3:]]
            cbias[b] = crm(v[b]) @ srot
            v[b] = v[b] + srot
        else:
            for d in ds:  # an axis fixed in the parent: derivative = (velocity before this joint) x S
                cbias[b] += crm(v[b]) @ S[d] * qvel[d]
                v[b] = v[b] + S[d] * qvel[d]
        IA[b] = xf_inertia(b)
        pA[b] = -crm(v[b]).T @ (IA[b] @ v[b])  # v x* (I v)
    U, Dinv, u = {}, {}, {}
    for b in range(nb - 1, 0, -1):
        ds = dofs_of[b]
        if ds:
            Sb = S[ds].T  # 6 x k
            Ub = IA[b] @ Sb
            Di = np.linalg.inv(Sb.T @ Ub + np.diag(np.asarray(m.dof_armature, dtype=np.float64)[ds]))
            ub = np.asarray(tau, dtype=np.float64)[ds] - Sb.T @ pA[b]
            U[b], Dinv[b], u[b] = Ub, Di, ub
            Ia = IA[b] - Ub @ Di @ Ub.T
            pa = pA[b] + Ia @ cbias[b] + Ub @ Di @ ub
        else:
            Ia, pa = IA[b], pA[b] + IA[b] @ cbias[b]
        p = m.body_parentid[b]
        if p > 0:
            IA[p] += Ia
            pA[p] += pa
    a = np.zeros((nb, 6))
    a[0, 3:] = -np.asarray(gravity, dtype=np.float64)
    qacc = np.zeros(nv)
    for b in range(1, nb):
        p = m.body_parentid[b]
        ab = a[p] + cbias[b]
        ds = dofs_of[b]
        if ds:
            qdd = Dinv[b] @ (u[b] - U[b].T @ ab)
            qacc[ds] = qdd
            ab = ab + S[ds].T @ qdd
        a[b] = ab
    return qacc


# ----------------------------------------------------------------------------- constrained acceleration, generic minimiser
def constraint_minimum(M, qacc_smooth, J, D, aref, ne, x0=None):
    """argmin_a 1/2 (a - a_s)' M (a - a_s) + 1/2 sum_i D_i r_i^2 [i < ne or r_i < 0], r = J a - aref (App. A.8), by scipy's
    L-BFGS-B on the analytic gradient, polished with a few exact Newton steps on the final active set."""
    from scipy.optimize import minimize

    M, J, D, aref = (np.asarray(x, dtype=np.float64) for x in (M, J, D, aref))
    a_s = np.asarray(qacc_smooth, dtype=np.float64)
    eq = np.arange(len(D)) < ne
    scale = 1.0 / max(1e-12, np.mean(np.diag(M)))

    def fg(a):
        r = J @ a - aref
        act = eq | (r < 0)
        e = a - a_s
        f = 0.5 * e @ M @ e + 0.5 * np.sum(D * r * r * act)
        g = M @ e + J.T @ (D * r * act)
        return f * scale, g * scale

    a0 = a_s.copy() if x0 is None else np.asarray(x0, dtype=np.float64)
    res = minimize(fg, a0, jac=True, method="L-BFGS-B", options=dict(maxiter=5000, ftol=1e-16, gtol=1e-12, maxcor=50))
    a = res.x
    for _ in range(20):  # piecewise-quadratic cost: Newton on the current active set lands on the minimiser once the set is right
        r = J @ a - aref
        act = eq | (r < 0)
        H = M + J.T @ (J * (D * act)[:, None])
        g = M @ (a - a_s) + J.T @ (D * r * act)
        step = np.linalg.solve(H, g)
        if np.abs(step).max() < 1e-13 * max(1.0, np.abs(a).max()):
            break
        # backtrack on the true cost
        f0 = fg(a)[0]
        t = 1.0
        while t > 1e-6 and fg(a - t * step)[0] > f0:
            t *= 0.5
        a = a - t * step
    return a


# ----------------------------------------------------------------------------------------------------------------------
# Convex collision along routes that share nothing with collision_convex's separating axes / face clipping: the minimum
# translation that separates two polytopes is the distance from the origin to the boundary of their Minkowski difference
# (scipy's qhull), and the distance from a point / a segment to a polytope is a small constrained least-squares problem.
def geom_world_vertices(m, g, xpos, xrot):
    """World-frame hull vertices of convex geom g (box / mesh) for body frames (xpos, xrot) from `kinematics`."""
    b = int(m.geom_bodyid[g])
    R = xrot[b] @ _qmat(m.geom_quat[g])
    p = xpos[b] + xrot[b] @ m.geom_pos[g]
    v = m.vert[m.geom_vertadr[g]:m.geom_vertadr[g] + m.geom_vertnum[g]]
    return p + v @ R.T


def polytope_penetration(VA, VB):
    """(depth, normal): the smallest translation of B along `normal` (from A to B) that separates the two hulls; depth < 0 = the gap
    along the best separating facet direction when they do not touch. Through the facets of hull(B - A)."""
    from scipy.spatial import ConvexHull

    D = (VB[None, :, :] - VA[:, None, :]).reshape(-1, 3)  # Minkowski difference B - A: contains the origin iff the hulls overlap
    eq = ConvexHull(D).equations  # n.x + d <= 0 inside
    # distance from the origin to each facet plane along its outward normal: -d (positive when the origin is inside)
    # facet normal n: max_B n.b - min_A n.a = -d is smallest there, i.e. B lies on the -n side of A: the direction from A to B is -n
    k = int(np.argmin(-eq[:, 3]))
    return float(-eq[k, 3]), -eq[k, :3].copy()


def point_polytope_distance(p, V):
    """Euclidean distance from point p to hull(V) (0 inside) and the closest point, by SLSQP over convex weights."""
    from scipy.optimize import minimize

    n = len(V)
    f = lambda w: float(np.sum((w @ V - p) ** 2))
    g = lambda w: 2.0 * V @ (w @ V - p)
    best = None
    for w0 in (np.full(n, 1.0 / n), np.eye(n)[int(np.argmin(np.sum((V - p) ** 2, axis=1)))]):
        r = minimize(f, w0, jac=g, bounds=[(0, 1)] * n, constraints=[{"type": "eq", "fun": lambda w: w.sum() - 1, "jac": lambda w: np.ones(n)}],
                     method="SLSQP", options={"ftol": 1e-16, "maxiter": 500})
        if best is None or r.fun < best.fun:
            best = r
    return float(np.sqrt(max(best.fun, 0.0))), best.x @ V


def segment_polytope_distance(a, b, V):
    """Distance between the segment (a, b) and hull(V): SLSQP over (t, convex weights)."""
    from scipy.optimize import minimize

    n = len(V)

    def f(z):
        return float(np.sum((a + z[0] * (b - a) - z[1:] @ V) ** 2))

    def g(z):
        e = a + z[0] * (b - a) - z[1:] @ V
        return np.concatenate([[2.0 * e @ (b - a)], -2.0 * V @ e])

    best = None
    for t0 in (0.0, 0.5, 1.0):
        pt = a + t0 * (b - a)
        w0 = np.eye(n)[int(np.argmin(np.sum((V - pt) ** 2, axis=1)))]
        r = minimize(f, np.concatenate([[t0], w0]), jac=g, bounds=[(0, 1)] * (n + 1),
                     constraints=[{"type": "eq", "fun": lambda z: z[1:].sum() - 1, "jac": lambda z: np.concatenate([[0.0], np.ones(n)])}],
                     method="SLSQP", options={"ftol": 1e-16, "maxiter": 500})
        if best is None or r.fun < best.fun:
            best = r
    return float(np.sqrt(max(best.fun, 0.0)))
