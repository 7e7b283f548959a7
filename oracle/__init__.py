"""CPU oracle for the batched motion-imitation env step (TEST INFRASTRUCTURE ONLY).

This package is a float64 NumPy restatement of the reference hot path
(`OpenRoboRL/envs/quadruped_robot/...`).  It is the
checker for the CUDA path: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  Nothing
under `openroborl_b200/` imports it, and the product path has no CPU fallback.

Pinning status
--------------
* Task logic (motion-file post-processing, frame lookup/interpolation, quaternion
  helpers, Butterworth filter, PD motor model, reward terms, termination, target
  observation, reset/RSI bookkeeping, observation pipeline): PINNED against the
  reference's own, unmodified Python modules imported in the build container:
  - scripts/make_golden.py -> tests/golden/ref_motion_kat.npz: known answers of
    motion_data.py / pose3d.py / action_filter.py / minitaur_motor.py
    (tests/test_oracle_golden.py; frame indices and blend bit-exact);
  - scripts/make_trace.py -> tests/golden/ref_task_trace.npz: step-by-step traces of the
    reference's Minitaur + ImitationTask + LocomotionGymEnv + WrapperEnv classes running on
    a fake BulletClient (scripts/fake_bullet.py) that serves the PyBullet calls from
    oracle/rbd.py (tests/test_oracle_trace.py: observations 1e-6, rewards 1e-7, done flags
    and the motion clock exact);
  - the observation-space bounds serialised inside the reference's policy zips;
  - scripts/make_golden.py -> tests/golden/ref_randomizer_kat.npz: sampled parameters and robot-setter arguments of
    the reference's own ControllableEnvRandomizerFromConfig (oracle/randomizer.py, bit-exact).
* Learner arithmetic (oracle/ppo.py): GAE with the reference's index quirk and the MpiAdam update are PINNED to the
  reference's own functions, executed unmodified under stub tensorflow / mpi4py / stable_baselines modules
  (scripts/make_ppo_golden.py -> tests/golden/ref_ppo_kat.npz: add_vtarg_and_adv bit for bit, MpiAdam.update on one and
  two ranks to float32 round-off; tests/test_oracle_ppo_golden.py).  The PPO1 loss needs a TF graph to evaluate and stays
  a restatement (torch autograd vs the analytic gradient).
  The third-party `pybullet_utils.transformations` (pybullet==2.7.3, requirements.txt:5) is
  absent from the container, so its published algorithm is restated in `oracle/xform.py`.
* Rigid-body arithmetic that the reference delegates to PyBullet 2.7.3
  (`stepSimulation`, `getLinkState`, contacts): PARITY UNPINNED.  PyBullet is not
  installable here (no wheel, no network), so forward kinematics, the mass matrix,
  inverse dynamics and the contact model are checked against an independent
  textbook formulation (`oracle/rbd.py`: maximal-coordinate Jacobians on the raw
  URDF tree) and by physical invariants, not against Bullet outputs.
"""
