"""Independent numpy restatement of forward kinematics / mass matrix / potential energy, used to
check the C oracle's kinematics, CRB and RNE stages by finite differences (test helper)."""
import numpy as np

from robopianist_b200.scene import mathutil as mu


def fk(fm, q):
    nb = fm.meta["nbody"]
    xpos = np.zeros((nb, 3)); xquat = np.zeros((nb, 4)); xquat[0] = [1, 0, 0, 0]
    nv = fm.meta["nv"]
    anchor = np.zeros((nv, 3)); axis = np.zeros((nv, 3))
    for b in range(1, nb):
        p = fm.body_parentid[b]
        pos = xpos[p] + mu.rotate(xquat[p], fm.body_pos[b])
        quat = mu.quat_mul(xquat[p], fm.body_quat[b])
        for k in range(fm.body_dofnum[b]):
            j = fm.body_dofadr[b] + k
            anchor[j] = pos + mu.rotate(quat, fm.jnt_pos[j])
            axis[j] = mu.rotate(quat, fm.jnt_axis[j])
            if fm.jnt_type[j] == 2:
                pos = pos + axis[j] * q[j]
            else:
                quat = mu.quat_mul(quat, mu.axis_angle_to_quat(fm.jnt_axis[j], q[j]))
                pos = anchor[j] - mu.rotate(quat, fm.jnt_pos[j])
        xpos[b] = pos; xquat[b] = mu.quat_normalize(quat)
    return xpos, xquat, anchor, axis


def chains(fm):
    nb = fm.meta["nbody"]
    ch = [[] for _ in range(nb)]
    for b in range(1, nb):
        ch[b] = list(ch[fm.body_parentid[b]]) + [fm.body_dofadr[b] + k for k in range(fm.body_dofnum[b])]
    return ch


def mass_matrix(fm, q):
    xpos, xquat, anchor, axis = fk(fm, q)
    nv, nb = fm.meta["nv"], fm.meta["nbody"]
    ch = chains(fm)
    M = np.zeros((nv, nv))
    for b in range(1, nb):
        m = fm.body_mass[b]
        if m <= 0:
            continue
        com = xpos[b] + mu.rotate(xquat[b], fm.body_ipos[b])
        jp = np.zeros((3, nv)); jr = np.zeros((3, nv))
        for j in ch[b]:
            if fm.jnt_type[j] == 3:
                jr[:, j] = axis[j]; jp[:, j] = np.cross(axis[j], com - anchor[j])
            else:
                jp[:, j] = axis[j]
        R = mu.quat_to_mat(mu.quat_mul(xquat[b], fm.body_iquat[b]))
        Iw = R @ np.diag(fm.body_inertia[b]) @ R.T
        M += m * jp.T @ jp + jr.T @ Iw @ jr
    M[np.diag_indices(nv)] += fm.dof_armature
    return M


def potential(fm, q):
    xpos, xquat, _, _ = fk(fm, q)
    g = fm.opt_gravity
    e = 0.0
    for b in range(1, fm.meta["nbody"]):
        com = xpos[b] + mu.rotate(xquat[b], fm.body_ipos[b])
        e -= fm.body_mass[b] * np.dot(g, com)
    return e


def bias_fd(fm, q, v, dofs, eps=1e-6):
    """qfrc_bias restricted to `dofs` by finite differences: c = Mdot v - 0.5 d(v'Mv)/dq + dU/dq."""
    nv = fm.meta["nv"]
    M0 = mass_matrix(fm, q)
    dM = {}
    out = np.zeros(nv)
    Mdot = np.zeros((nv, nv))
    for k in dofs:
        qp = q.copy(); qp[k] += eps; qm = q.copy(); qm[k] -= eps
        dMk = (mass_matrix(fm, qp) - mass_matrix(fm, qm)) / (2 * eps)
        dM[k] = dMk
        Mdot += dMk * v[k]
        out[k] = -0.5 * v @ dMk @ v + (potential(fm, qp) - potential(fm, qm)) / (2 * eps)
    out += Mdot @ v
    return out
