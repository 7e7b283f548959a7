"""Ad-hoc parity report on a GPU box: prints the error of every entry point against the oracle.

    python scripts/gpu_check.py [N]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common  # noqa: E402


def report(name, a, b, kind="abs"):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    scale = np.abs(b).max()
    print("%-28s max abs %.3e  (scale %.3e, rel-to-scale %.3e)" % (name, err.max(), scale, err.max() / max(scale, 1e-30)))


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    for robot, clipn in (("laikago", ["laikago_pace"]), ("mini_cheetah", ["minicheetah_trot"]), ("laikago", ["laikago_spin", "laikago_hopturn"])):
        print("=====", robot, clipn)
        env = common.make_product(robot, clipn, N, seed=5)
        orc = common.make_oracle(robot, clipn, N, seed=5, phys=common.oracle_phys_from_product(env.phys))
        rng = np.random.default_rng(1)
        # motion tables + lookup
        mt = env.motion_tables()
        report("motion frames", mt["frames"], orc.lib.frames)
        report("motion vels", mt["vels"], orc.lib.vels)
        clip = rng.integers(0, len(clipn), size=500)
        t = rng.uniform(-0.3, 25.0, size=500)
        lk = env.motion_lookup(clip, t)
        f0, f1, bl = orc.lib.calc_blend_idx(clip, t)
        print("f0 exact", np.array_equal(lk["f0"], f0), "f1 exact", np.array_equal(lk["f1"], f1), "blend exact",
              np.array_equal(lk["blend"], bl), "cycle exact", np.array_equal(lk["cycle"], orc.lib.calc_cycle_count(clip, t)))
        report("calc_frame", lk["pose"], orc.lib.calc_frame(clip, t))
        report("calc_frame_vel", lk["vel"], orc.lib.calc_frame_vel(clip, t))
        # FK / M / ID / FD on random states
        pose, vel = common.random_states(rng, N, robot)
        p64, v64 = pose.astype(np.float64), vel.astype(np.float64)
        report("fk link com", env.fk(pose), orc.tree.link_com(p64[:, 0:3], p64[:, 3:7], p64[:, 7:]))
        perm = [3, 4, 5, 0, 1, 2] + list(range(6, 18))   # interface [lin, ang, q] <- oracle [ang, lin, q]
        M_o = orc.tree.mass_matrix(p64[:, 0:3], p64[:, 3:7], p64[:, 7:])[:, perm][:, :, perm]
        report("mass matrix", env.mass_matrix(pose), M_o)
        acc = rng.normal(size=(N, 18)).astype(np.float32) * 5
        nu = np.concatenate([v64[:, 3:6], v64[:, 0:3], v64[:, 6:]], axis=1)
        nud = np.concatenate([acc[:, 3:6], acc[:, 0:3], acc[:, 6:]], axis=1).astype(np.float64)
        tau_o = orc.tree.inverse_dynamics(p64[:, 0:3], p64[:, 3:7], p64[:, 7:], nu, nud)[:, perm]
        report("inverse dynamics", env.inverse_dynamics(pose, vel, acc), tau_o)
        # forward dynamics with contact: put robots near the ground
        pose2 = pose.copy()
        pose2[:, 2] = rng.uniform(0.2, 0.5, size=N)
        pose2[:, 3:7] = np.array(env.model.spec.INIT_QUAT) + rng.normal(size=(N, 4)) * 0.1
        pose2[:, 3:7] /= np.linalg.norm(pose2[:, 3:7], axis=1, keepdims=True)
        tau = rng.normal(size=(N, 12)).astype(np.float32) * 5
        orc.pos, orc.quat, orc.q = pose2[:, 0:3].astype(np.float64), pose2[:, 3:7].astype(np.float64), pose2[:, 7:].astype(np.float64)
        orc.lin, orc.ang, orc.qd = v64[:, 0:3], v64[:, 3:6], v64[:, 6:]
        acc_o = orc.forward_dynamics(tau.astype(np.float64))[:, perm]
        report("forward dynamics+contact", env.forward_dynamics(pose2, vel, tau), acc_o)
        # reset + steps
        obs = env.reset().cpu().numpy()
        obs_o = orc.reset()
        report("reset obs", obs, obs_o)
        st = env.get_state()
        for k in ("pos", "quat", "q", "lin", "ang", "qd", "ref_pose", "ref_vel", "origin_pos", "origin_rot"):
            report("  reset " + k, st[k], getattr(orc, k))
        print("  reset clip/warm/toff exact:", np.array_equal(st["clip"], orc.clip), np.array_equal(st["warmup"].astype(bool), orc.warmup),
              np.array_equal(st["time_offset"], orc.time_offset), np.array_equal(st["prev_phase"], orc.prev_phase))
        for step in range(3):
            common.sync_oracle_from_product(orc, env.get_state())
            act = rng.normal(scale=0.2, size=(N, 12)).astype(np.float32)
            o, r, d = env.step(torch.as_tensor(act, device=env.device))
            torch.cuda.synchronize()
            t0 = time.time()
            oo, ro, do = orc.step(act.astype(np.float64))
            st = env.get_state()
            print("-- step", step, "(oracle %.1fs) done equal:" % (time.time() - t0), np.array_equal(d.cpu().numpy().astype(bool), do), "n_done", int(do.sum()))
            report("  reward", r.cpu().numpy(), ro)
            report("  obs", o.cpu().numpy(), oo)
            for k in ("pos", "quat", "q", "lin", "ang", "qd", "ref_pose", "ref_vel", "origin_pos", "imu_hist", "act_hist", "ang_hist",
                      "filt_x", "filt_y", "prev_filtered"):
                report("  " + k, st[k], getattr(orc, k))
            report("  ring1", st["ring1"], orc.hist[:, 1, 0:12])
            report("  ring2", st["ring2"], orc.hist[:, 2, 0:12])
            print("  prev_phase exact", np.array_equal(st["prev_phase"], orc.prev_phase), "counters",
                  np.array_equal(st["substep_counter"], orc.substep_counter), np.array_equal(st["env_step_counter"], orc.env_step_counter))
            report("  reward_terms", env.reward_terms(), orc.reward_terms())
        print("stats", env.stats(), "launches", env.launch_count)
        env.close()
    # quick timing
    for n in (4096, 8192):
        env = common.make_product("laikago", ["laikago_pace"], n, seed=0, auto_reset=True)
        env.reset()
        act = torch.zeros((n, 12), device=env.device)
        for _ in range(20):
            env.step(act)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 200
        for _ in range(K):
            env.step(act)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print("N=%d  %.3f ms/step  %.3e env-steps/s" % (n, ms, n / ms * 1e3), env.stats())
        env.close()


if __name__ == "__main__":
    main()
