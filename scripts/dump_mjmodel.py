#!/usr/bin/env python
"""Dumps what pins this repository to real MuJoCo — run it on a machine that HAS the reference's dependencies
(`pip install mujoco dm_control mujoco_utils note_seq ...` + robopianist's install_deps.sh); none of them is installable
in the build container or on the GPU boxes, which is why DESIGN.md says "parity unpinned".

    python scripts/dump_mjmodel.py [--out tests/golden/mujoco_reference.npz] [--gravcomp 1]
    python scripts/dump_mjmodel.py --hand-collisions --hands-together --out tests/golden/mujoco_reference_hc.npz
    python scripts/dump_mjmodel.py --reduced-action-space --out tests/golden/mujoco_reference_ras.npz

Writes one .npz with
  * the compiled mjModel fields the scene flattener restates (scene/compile.py): sizes, body / joint / dof / geom /
    actuator arrays incl. the compiler-derived ones (dof_damping of the forearm slides = 2 sqrt(kp M0):
    shadow_hand.py:299-301, the inertia-box capsule fitted to the fingertip meshes: shadow_hand.py:144-152,
    dof_invweight0 / body_invweight0, dof_M0);
  * a 1000-substep mj_step trajectory (legacy control-step loop of dm_control: 10 x (mj_step2, mj_step1)) of the
    recorded action replay (examples/twinkle_twinkle_actions.npy) with the solver overridden to what
    BASELINE.json's configs name (PGS, elliptic cones, 50 iterations, tolerance 1e-8): qpos, qvel, ncon and
    solver_niter after every control step.
tests/test_vs_mujoco.py consumes the file (and, when `mujoco` is importable, regenerates the comparison live).
"""
import argparse
from pathlib import Path

import numpy as np

MODEL_FIELDS = [
    "nq", "nv", "nu", "nbody", "ngeom", "njnt", "nsite", "ntendon",
    "body_parentid", "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "body_invweight0",
    "body_gravcomp", "jnt_type", "jnt_bodyid", "jnt_pos", "jnt_axis", "jnt_range", "jnt_limited", "jnt_stiffness",
    "jnt_solref", "jnt_solimp", "jnt_margin", "qpos_spring", "dof_parentid", "dof_damping", "dof_armature", "dof_frictionloss",
    "dof_solref", "dof_solimp", "dof_invweight0", "dof_M0", "geom_type", "geom_bodyid", "geom_pos", "geom_quat", "geom_size",
    "geom_contype", "geom_conaffinity", "geom_condim", "geom_priority", "geom_friction", "geom_solref", "geom_solimp",
    "geom_solmix", "geom_margin", "geom_gap", "actuator_trntype", "actuator_trnid", "actuator_gear", "actuator_gainprm",
    "actuator_biasprm", "actuator_ctrlrange", "actuator_forcerange", "actuator_ctrllimited", "actuator_forcelimited",
    "tendon_adr", "tendon_num", "wrap_objid", "wrap_prm", "site_bodyid", "site_pos",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(Path(__file__).resolve().parents[1] / "tests" / "golden" / "mujoco_reference.npz"))
    ap.add_argument("--gravcomp", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100, help="control steps (x10 physics substeps)")
    ap.add_argument("--hand-collisions", action="store_true", help="disable_hand_collisions=False (the reference default): the sphere / "
                    "capsule pair routines restated here are MuJoCo's own, so hand : hand contacts can be pinned exactly")
    ap.add_argument("--hands-together", action="store_true", help="overwrite the forearm_tx actions so that the hands run into each other "
                    "(cross-hand contacts; with the piano far below the crossed thumbs no capsule : box contact interferes at first)")
    ap.add_argument("--reduced-action-space", action="store_true")
    args = ap.parse_args()

    import mujoco
    from robopianist import suite

    env = suite.load(
        "RoboPianist-debug-TwinkleTwinkleRousseau-v0",
        task_kwargs=dict(trim_silence=True, gravity_compensation=bool(args.gravcomp), primitive_fingertip_collisions=True,
                         disable_hand_collisions=not args.hand_collisions, reduced_action_space=args.reduced_action_space,
                         n_steps_lookahead=1, control_timestep=0.05))
    physics = env.physics
    m, d = physics.model.ptr, physics.data.ptr
    m.opt.solver = mujoco.mjtSolver.mjSOL_PGS
    m.opt.cone = mujoco.mjtCone.mjCONE_ELLIPTIC
    m.opt.iterations = 50
    m.opt.tolerance = 1e-8
    out = {f: np.array(getattr(m, f)) for f in MODEL_FIELDS if hasattr(m, f)}
    out["opt"] = np.array([m.opt.timestep, m.opt.impratio, m.opt.tolerance, m.opt.iterations, m.stat.meaninertia])
    out["opt_gravity"] = np.array(m.opt.gravity)
    for kind, n, name_fn in (("body", m.nbody, mujoco.mjtObj.mjOBJ_BODY), ("joint", m.njnt, mujoco.mjtObj.mjOBJ_JOINT),
                             ("geom", m.ngeom, mujoco.mjtObj.mjOBJ_GEOM), ("actuator", m.nu, mujoco.mjtObj.mjOBJ_ACTUATOR)):
        out[f"{kind}_names"] = np.array([mujoco.mj_id2name(m, name_fn, i) or "" for i in range(n)])
    # action replay (canonical [-1, 1] -> spec, as CanonicalSpecWrapper / the tutorial do)
    acts = np.load(Path(suite.__file__).resolve().parents[2] / "examples" / "twinkle_twinkle_actions.npy")
    spec = env.action_spec()
    names = [mujoco.mj_id2name(m, mujoco.mjtObj.mjOBJ_ACTUATOR, i) for i in range(m.nu)]
    if args.reduced_action_space:  # the fixture has the full 45 columns: drop the removed actuators' (matched by name on the full model)
        full = suite.load("RoboPianist-debug-TwinkleTwinkleRousseau-v0", task_kwargs=dict(primitive_fingertip_collisions=True))
        fm_ = full.physics.model.ptr
        full_names = [mujoco.mj_id2name(fm_, mujoco.mjtObj.mjOBJ_ACTUATOR, i) for i in range(fm_.nu)]
        acts = acts[:, [full_names.index(n) for n in names] + [acts.shape[1] - 1]]
    acts = spec.minimum + (acts + 1) / 2 * (spec.maximum - spec.minimum)
    if args.hands_together:
        for i, n in enumerate(names):
            if n.endswith("forearm_tx"):
                acts[:, i] = (-0.2 if "rh_" in n else 0.2) * np.minimum(1.0, np.arange(acts.shape[0]) / 20.0)
    ts = env.reset()
    qpos, qvel, ncon, niter, rew = [np.array(d.qpos)], [np.array(d.qvel)], [d.ncon], [0], [0.0]
    for t in range(min(args.steps, acts.shape[0])):
        ts = env.step(acts[t])
        qpos.append(np.array(d.qpos)); qvel.append(np.array(d.qvel)); ncon.append(d.ncon)
        niter.append(int(d.solver_niter[0]) if hasattr(d.solver_niter, "__len__") else int(d.solver_niter))
        rew.append(float(ts.reward or 0.0))
    out["flags"] = np.array([int(args.hand_collisions), int(args.hands_together), int(args.reduced_action_space), int(args.gravcomp)])
    out.update(traj_qpos=np.array(qpos), traj_qvel=np.array(qvel), traj_ncon=np.array(ncon), traj_niter=np.array(niter),
               traj_reward=np.array(rew), traj_actions=acts[:args.steps], obs_keys=np.array(list(ts.observation.keys())),
               mujoco_version=np.array(mujoco.__version__))
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, {k: v.shape for k, v in out.items() if k.startswith("traj")})


if __name__ == "__main__":
    main()
