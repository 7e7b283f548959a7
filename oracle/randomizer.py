"""Domain randomizer of the reference, restated (test infrastructure only).

Follows OpenRoboRL/envs/utilities/randomizer/controllable_env_randomizer_from_config.py:92-122 (randomize_env: one
pass over the parameter names in SORTED order) with the ranges of minitaur_env_randomizer_config.py:22-36
("all_params") and param_bounds = (-1, 1) (:40-45).  Every draw is np.random.uniform(b0, b1[, size]) =
b0 + (b1 - b0) * u with u from random_sample(), mapped to (sample - b0) / (b1 - b0) * (upper - lower) + lower.

Pinned by tests/golden/ref_randomizer_kat.npz: the reference's own class driven with a recording robot stub and a
seeded RandomState (scripts/make_golden.py: randomizer_kat).
"""
import numpy as np

ALL_PARAMS = {                      # minitaur_env_randomizer_config.py:22-36
    "mass": (0.8, 1.2), "inertia": (0.5, 1.5), "motor strength": (0.8, 1.2), "motor friction": (0.0, 0.05),
    "latency": (0.0, 0.04), "lateral friction": (0.5, 1.25), "battery": (14.0, 16.8), "joint friction": (0.0, 0.05),
}
PARAM_BOUNDS = (-1.0, 1.0)          # controllable_env_randomizer_from_config.py:43
NUM_MOTORS = 12


def draws_per_param(num_knee_joints=8):
    """Number of uniforms each parameter consumes (:196, :316, :357, :344, :423, :441, :376, :483)."""
    return {"battery": 1, "inertia": 2, "joint friction": num_knee_joints, "latency": 1, "lateral friction": 1, "mass": 2,
            "motor friction": 1, "motor strength": NUM_MOTORS}


def total_draws(num_knee_joints=8):
    return sum(draws_per_param(num_knee_joints).values())


def sample(u, num_knee_joints=8, params=ALL_PARAMS, bounds=PARAM_BOUNDS):
    """u: (..., total_draws) uniforms in [0, 1) in the reference's consumption order (sorted parameter names).

    Returns {name: physical value(s)} with the reference's arithmetic (float64)."""
    u = np.asarray(u, dtype=np.float64)
    out, pos = {}, 0
    cnt = draws_per_param(num_knee_joints)
    b0, b1 = bounds
    for name in sorted(params):
        lo, hi = params[name]
        n = cnt[name]
        s = b0 + (b1 - b0) * u[..., pos:pos + n]                 # np.random.uniform
        out[name] = (s - b0) / (b1 - b0) * (hi - lo) + lo
        pos += n
    assert pos == u.shape[-1]
    return out


def apply(values, ids, link_mass, knee_order):
    """What the robot setters do with the sampled values (minitaur.py:935-1070, 1136-1143).

    ids: classify_links() of the robot; link_mass: (n_links + 1,) URDF masses, index 0 = base (link id -1).
    Returns per-robot arrays: mass (N, nb), inertia_scale (N, nb), motor_strength (N, 12), latency (N,),
    foot_friction (N,), joint_friction (N, n_links) (Coulomb torque limit on the joint of that link id).
    Battery voltage and motor viscous damping are sampled but have no effect (minitaur_motor.py:92-102)."""
    mass_r, inertia_r = values["mass"], values["inertia"]
    N = mass_r.shape[0]
    nb = link_mass.shape[0]
    ratio_m = np.ones((N, nb)); ratio_i = np.ones((N, nb))
    for i in ids["chassis"]:                                       # set_base_mass / set_base_inertia
        ratio_m[:, i + 1] = mass_r[:, 0]; ratio_i[:, i + 1] = inertia_r[:, 0]
    for i in list(ids["leg"]) + list(ids["motor"]):               # set_leg_mass / set_leg_inertia
        ratio_m[:, i + 1] = mass_r[:, 1]; ratio_i[:, i + 1] = inertia_r[:, 1]
    jf = np.zeros((N, nb - 1))
    for link_id, col in zip(ids["foot"], range(values["joint friction"].shape[1])):   # set_joint_friction: zip(foot ids, samples)
        jf[:, link_id] = values["joint friction"][:, col]
    return dict(mass=ratio_m * link_mass[None, :], inertia_scale=ratio_i, motor_strength=values["motor strength"],
                latency=values["latency"][:, 0], foot_friction=values["lateral friction"][:, 0], joint_friction=jf)
