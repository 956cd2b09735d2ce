"""Mesh -> primitive fitting for the fingertip colliders.

``primitive_fingertip_collisions=True`` in the reference flips the five ``*distal_pst`` mesh
geoms to ``type="capsule"`` (robopianist/models/hands/shadow_hand.py:144-152) and leaves the
sizing to MuJoCo's compiler, which fits the primitive to the mesh's equivalent inertia box
(MuJoCo XML reference, ``compiler/fitaabb=false``).  MuJoCo is not available here, so the fit is
restated: uniform-density inertia of the mesh's convex hull -> centre of mass + principal axes
-> equivalent inertia box half-sizes -> capsule ``radius=(bx+by)/2``,
``half_length=max(0, bz-radius/2)`` with z the axis of least inertia.  UNVERIFIED against a
real MuJoCo compile (parity unpinned).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import mathutil as mu


def load_obj_vertices(path: Path, scale) -> np.ndarray:
    verts = []
    with open(path, "r") as f:
        for line in f:
            if line.startswith("v "):
                p = line.split()
                verts.append([float(p[1]), float(p[2]), float(p[3])])
    return np.asarray(verts, dtype=np.float64) * np.asarray(scale, dtype=np.float64)


def hull_mass_properties(verts: np.ndarray):
    """Volume, centre of mass and inertia tensor (about the COM, unit density) of the convex hull."""
    from scipy.spatial import ConvexHull

    hull = ConvexHull(verts)
    center = verts[hull.vertices].mean(axis=0)
    vol = 0.0
    com = np.zeros(3)
    cov = np.zeros((3, 3))  # integral of x x^T
    canon = np.array([[2.0, 1, 1], [1, 2, 1], [1, 1, 2]]) / 120.0
    for simplex, eq in zip(hull.simplices, hull.equations):
        a, b, c = verts[simplex] - center
        # orient outward
        if np.dot(np.cross(b - a, c - a), eq[:3]) < 0:
            b, c = c, b
        A = np.stack([a, b, c], axis=1)  # columns
        det = np.linalg.det(A)
        v = det / 6.0
        vol += v
        com += v * (a + b + c) / 4.0
        cov += det * (A @ canon @ A.T)
    com /= vol
    # shift covariance to COM: C' = C - vol * com com^T
    cov_c = cov - vol * np.outer(com, com)
    inertia = np.trace(cov_c) * np.eye(3) - cov_c
    return vol, com + center, inertia


def fit_capsule(verts: np.ndarray):
    """Returns (pos, quat, radius, half_length) of the inertia-box capsule in the mesh frame."""
    vol, com, inertia = hull_mass_properties(verts)
    w, v = np.linalg.eigh(inertia)
    order = np.argsort(-w)  # decreasing principal inertia: z = axis of least inertia
    w = w[order]
    v = v[:, order]
    if np.linalg.det(v) < 0:
        v[:, 2] = -v[:, 2]
    # canonical sign: make the long axis point towards +z of the mesh frame when possible
    if v[2, 2] < 0:
        v[:, 2] = -v[:, 2]
        v[:, 1] = -v[:, 1]
    quat = mu.mat_to_quat(v)
    mass = vol
    box = 0.5 * np.sqrt(np.maximum(0.0, 6.0 * np.array([w[1] + w[2] - w[0], w[0] + w[2] - w[1], w[0] + w[1] - w[2]]) / mass))
    radius = 0.5 * (box[0] + box[1])
    half = max(0.0, box[2] - 0.5 * radius)
    return com, quat, float(radius), float(half)
