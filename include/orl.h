/*
 * orl.h -- C ABI of the B200-native batched motion-imitation environment.
 *
 * The reference (Derek-TH-Wang/OpenRoboRL) is pure Python and has NO FFI/plugin interface for its
 * env step; its boundary is the duck-typed Python env object.  This header is
 * the C boundary that sits directly under that Python surface: every entry point names the reference
 * method(s) it replaces (R/ = OpenRoboRL/).  The Python mirror of the reference classes
 * (openroborl_b200/compat.py) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every call returns 0 on success, a negative orl_status otherwise; no C++ exception crosses;
 *     orl_last_error(env) gives the message of the last failure on that handle.
 *   - `*_dev` pointers are CUDA device pointers on the env's device, `stream` is a cudaStream_t passed
 *     as void* (NULL = legacy default stream).  Calls are asynchronous on that stream unless the name
 *     ends in `_host`.  A handle is not re-entrant: one host thread per handle / GPU.
 *   - the caller owns all buffers it passes; the env owns its state and hands out non-owning views.
 *   - pose layout  (19 floats): root pos xyz, root quat xyzw, 12 joint angles in URDF joint order
 *     velocity layout (18)   : root linear vel (world), root angular vel (world), 12 joint velocities
 *     (R/envs/quadruped_robot/task/motion_data.py:41-44,278-415).
 *   - observation layout (160): IMU(4)x3, LastAction(12)x3, MotorAngle(12)x3, newest first, then 4 target
 *     frames of 19 (R/envs/quadruped_robot/robots/minitaur.py:529-541, R/.../wrapper_env.py:109-125).
 */
#ifndef ORL_H_
#define ORL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORL_ABI_VERSION 3
#define ORL_NUM_LEGS 4
#define ORL_LEG_DOF 3
#define ORL_NUM_MOTORS 12
#define ORL_POSE_SIZE 19
#define ORL_VEL_SIZE 18
#define ORL_OBS_SIZE 160
#define ORL_NUM_TAR_FRAMES 4
#define ORL_MAX_CP 24          /* collision points per leg list (leg bodies + its share of the base) */
#define ORL_MAX_LINKS 32       /* URDF links (without the base) reported by orl_fk */
#define ORL_CYCLE_TABLE 128    /* pre-summed cycle offsets per rotating clip */

typedef enum {
  ORL_OK = 0,
  ORL_ERR_INVALID = -1,      /* bad argument / config (the reference raises ValueError) */
  ORL_ERR_CUDA = -2,         /* CUDA runtime failure */
  ORL_ERR_STATE = -3,        /* call order (e.g. step before motions are loaded) */
  ORL_ERR_ASSERT = -4        /* a reference `assert` would have fired */
} orl_status;

typedef struct orl_env orl_env;

/* Environment + task configuration.
 * Replaces the hard-coded constructor arguments of R/run.py:49-76, the reward weights/scales of
 * R/envs/quadruped_robot/task/imitation_task.py:34-58 and R/config/pybullet_sim_param.yaml. */
typedef struct {
  int32_t abi_version;          /* ORL_ABI_VERSION */
  int32_t num_robot;            /* robots stepped by this handle (this GPU's shard) */
  int32_t device;               /* CUDA device ordinal */
  int32_t gid0;                 /* global id of robot 0 of this shard (keys the RNG; results do not depend on sharding) */
  uint64_t seed;
  int32_t action_repeat;        /* 33   (laikago.py:26) */
  int32_t max_episode_steps;    /* wrapper_env.py:151-159; see orl_set_max_episode_steps */
  int32_t auto_reset;           /* 0 = reference semantics (caller resets), 1 = per-robot auto reset */
  int32_t enable_cycle_sync;    /* run.py:60 */
  double sim_time_step;         /* 0.001 (pybullet_sim_param.yaml:3) */
  double env_time_step;         /* action_repeat * sim_time_step as the caller computes it (quadruped_gym_env.py:177) */
  double warmup_time;           /* 0.25 (run.py:64) */
  double ref_state_init_prob;   /* 0.9  (run.py:63) */
  int32_t tar_frame_steps[ORL_NUM_TAR_FRAMES]; /* 1,2,10,30 (run.py:61) */
  /* reward (imitation_task.py:45-55) */
  float pose_weight, velocity_weight, end_effector_weight, root_pose_weight, root_velocity_weight;
  float pose_err_scale, velocity_err_scale, end_effector_err_scale, end_effector_height_err_scale;
  float root_pose_err_scale, root_velocity_err_scale;
  /* termination (imitation_task.py:518) */
  float dist_fail_threshold, rot_fail_threshold;
  /* physics model of this repo (DESIGN.md "physics model"; replaces pybullet.stepSimulation) */
  float gravity_z;              /* -10 (quadruped_gym_env.py:200) */
  float contact_kn, contact_bn, contact_d0, contact_bt;
  float limit_k, limit_d, limit_e0;
  float max_coordinate_velocity; /* clamp of every generalised velocity coordinate (btMultiBody default: 100) */
  /* Domain randomizer.  Replaces ControllableEnvRandomizerFromConfig.randomize_env
   * (R/envs/utilities/randomizer/controllable_env_randomizer_from_config.py:92-122), run by Minitaur.reset when
   * enable_randomizer is set (minitaur.py:273-276).  Ranges in the reference's SORTED parameter order
   * (ORL_RAND_*; defaults minitaur_env_randomizer_config.py:22-36). */
  int32_t enable_randomizer;
  float rand_lo[8], rand_hi[8];
  float joint_friction_k;       /* slope of the regularised Coulomb joint friction, N m s / rad (this repo's model) */
  float mu_plane;               /* lateral friction of the ground plane (combined with the randomised foot friction) */
  /* Task options R/run.py leaves at their defaults (imitation_task.py:36-57).  Any non-default value selects the
   * option instantiation of the step / reset kernels; the defaults run the same code as ABI 2. */
  float reward_weight;          /* `weight`: reward() returns reward * weight (:104, 356); 1 */
  int32_t enable_rand_init_time;/* 0: every episode starts at motion time offset 0 (:1103-1106); 1 */
  double clip_time_min, clip_time_max; /* clip switching: a new clip after U(min, max) seconds of motion time
                                   (:1057-1101, 740-754); INFINITY (either) = off */
  double perturb_init_state_prob; /* probability of a perturbed start state (:192-195, 1199-1243); 0 */
  int32_t enable_tar_obs_noise; /* tar_obs_noise is not None (:274-275) */
  float tar_obs_noise;          /* tar_obs_noise[0]: std of the heading noise of the target observation */
} orl_config;

enum { ORL_RAND_BATTERY = 0, ORL_RAND_INERTIA = 1, ORL_RAND_JOINT_FRICTION = 2, ORL_RAND_LATENCY = 3,
       ORL_RAND_LATERAL_FRICTION = 4, ORL_RAND_MASS = 5, ORL_RAND_MOTOR_FRICTION = 6, ORL_RAND_MOTOR_STRENGTH = 7 };
#define ORL_RAND_DRAWS 28       /* 1 + 2 + 8 + 1 + 1 + 2 + 1 + 12 uniforms per reset, in that order */

/* One collision point: a sphere (radius may be 0) fixed to a body. */
typedef struct {
  float local[3];
  float radius;
  float mu;                     /* link lateral friction x plane friction */
  float margin;                 /* contact_fall margin (Bullet contact-breaking threshold) */
  int32_t body;                 /* -1 base, 0..2 leg body */
  int32_t is_foot;              /* link id in Minitaur._foot_link_ids (minitaur.py:842-848) */
} orl_contact_point;

/* Compiled kinematic/dynamic model of one leg (hip -> upper -> lower, fixed children merged).
 * Body frame k has its origin at joint k and its z axis along the joint axis. */
typedef struct {
  float R_fix[ORL_LEG_DOF][9];  /* row-major rotation parent body frame -> body frame at q = 0 */
  float p_fix[ORL_LEG_DOF][3];  /* joint origin in the parent body frame */
  float mass[ORL_LEG_DOF];
  float com[ORL_LEG_DOF][3];
  float inertia[ORL_LEG_DOF][6];/* about the COM, body frame: xx xy xz yy yz zz */
  float inertia_par[ORL_LEG_DOF][6]; /* the parallel-axis share of `inertia` (merged fixed children): it scales with the
                                   randomised MASS ratio, the rest with the INERTIA ratio (minitaur.py:935-1039) */
  int32_t rand_class[ORL_LEG_DOF];   /* 0: links of the chassis set (base ratios), 1: leg/motor set (leg ratios), -1: mixed */
  int32_t friction_draw[ORL_LEG_DOF];/* index of this joint's sample in set_joint_friction (minitaur.py:1063-1070) or -1 */
  float lower[ORL_LEG_DOF], upper[ORL_LEG_DOF], damping[ORL_LEG_DOF];
  /* motor model (laikago.py:50-66, mini_cheetah.py:51-67, minitaur_motor.py:104-173) */
  float dir[ORL_LEG_DOF], offset[ORL_LEG_DOF], kp[ORL_LEG_DOF], kd[ORL_LEG_DOF], init_angle[ORL_LEG_DOF];
  int32_t motor[ORL_LEG_DOF];   /* motor (= action/observation) index of joint k */
  int32_t n_ee;                 /* end-effector links on this leg (imitation_task.py:439-445) */
  int32_t ee_body[2];
  float ee_local[2][3];         /* COM of the end-effector link in its body frame */
  int32_t n_cp;
  orl_contact_point cp[ORL_MAX_CP];
} orl_leg_model;

/* Flattened robot model produced by the URDF compiler (openroborl_b200/model.py).
 * Replaces loadURDF + the id bookkeeping of R/envs/quadruped_robot/robots/minitaur.py:812-903. */
typedef struct {
  float base_mass;
  float base_inertia[6];        /* about the COM, base state frame (= base inertial frame) */
  float init_pos[3];            /* laikago.py:48 */
  float init_quat[4];           /* laikago.py:49 */
  float max_angle_change;       /* laikago.py:71 */
  orl_leg_model leg[ORL_NUM_LEGS];
  int32_t n_links;              /* URDF links without the base, PyBullet link-index order */
  int32_t link_leg[ORL_MAX_LINKS];   /* leg of link i */
  int32_t link_body[ORL_MAX_LINKS];  /* body of link i on that leg */
  float link_com[ORL_MAX_LINKS][3];  /* COM of link i in that body's frame (getLinkState()[0]) */
} orl_robot_model;

/* Offsets (in floats) of the named fields inside one robot's state record, for parity tests,
 * checkpointing and zero-copy torch views (SURVEY.md section 5 "checkpoint / resume"). */
typedef struct {
  int32_t stride;
  int32_t pos, quat, q, lin, ang, qd;
  int32_t filt_x, filt_y, prev_filtered, ring1, ring2;
  int32_t imu_hist, act_hist, ang_hist;
  int32_t ref_pose, ref_vel, origin_pos, origin_rot;
  int32_t time_offset /* f64 */, prev_phase /* f64 */;
  int32_t clip, warmup, substep_counter, step_counter, env_step_counter, episode /* int32 each */;
  int32_t ep_return;
  int32_t obs_quat;
  int32_t motor_strength;
  int32_t rand_params;          /* 10: base/leg mass ratio, base/leg inertia ratio, latency, foot friction x plane (-1: model), knee friction[4] */
  int32_t clip_change_time;     /* f64 (2 floats): motion time of the next clip change (imitation_task.py:123, 1057-1067; +inf: none) */
  int32_t clip_changes;         /* i32: clip changes of the running episode (RNG counter) */
} orl_state_layout;

/* ---- lifetime -------------------------------------------------------------------------------- */
/* Replaces build_env (R/run.py:49-76): Minitaur(...) x N, ImitationTask(...) x N, LocomotionGymEnv, WrapperEnv. */
int orl_create(const orl_config* cfg, const orl_robot_model* model, orl_env** out);
int orl_destroy(orl_env* env);
const char* orl_last_error(const orl_env* env);
int orl_abi_version(void);

/* Replaces MotionData.load/_postprocess_frames/_calc_frame_vels/_calc_cycle_delta_*
 * (R/envs/quadruped_robot/task/motion_data.py:72-112, 527-589, 635-680) and
 * ImitationTask._load_ref_motions (imitation_task.py:574-598).
 * frames: concatenated RAW frames [sum(n_frames)][19] exactly as in the JSON "Frames" key;
 * flags bit0 = LoopMode Wrap, bit1 = EnableCycleOffsetPosition, bit2 = EnableCycleOffsetRotation. */
int orl_load_motions(orl_env* env, const double* frames, const int32_t* n_frames, const double* frame_dur,
                     const int32_t* flags, int32_t n_clips);
/* Post-processed tables back on the host (for tests): frames_out [sum F][19], vels_out [sum F][18],
 * per clip: duration, cycle delta pos[3], cycle delta heading. Any pointer may be NULL. */
int orl_get_motion_tables(orl_env* env, double* frames_out, double* vels_out, double* duration,
                          double* delta_pos, double* delta_heading);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* Replaces WrapperEnv.reset -> LocomotionGymEnv.reset -> Minitaur.reset + ImitationTask.reset
 * (R/envs/quadruped_robot/wrapper_env.py:87-107, quadruped_gym_env.py:63-104, minitaur.py:232-278,
 * imitation_task.py:166-199).  mask_dev NULL = all robots.  obs_dev [N][160] may be NULL. */
int orl_reset(orl_env* env, const uint8_t* mask_dev, float* obs_dev, void* stream);
/* Replaces WrapperEnv.step -> LocomotionGymEnv._step (33 substeps) -> reward/update/termination ->
 * build_target_obs (wrapper_env.py:58-85, quadruped_gym_env.py:213-239, imitation_task.py:254-572).
 * actions_dev [N][12] policy residuals (NOT modified; the reference adds INIT in place, minitaur.py:281),
 * obs_dev [N][160], reward_dev [N], done_dev [N]. */
int orl_step(orl_env* env, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
             void* stream);
/* Same through HOST buffers (what the list-based reference API needs): H2D, step, D2H, synchronise.
 * Page-locked buffers (cudaMallocHost / cudaHostRegister) are used for DMA directly; pageable ones go through
 * the library's pinned staging area in overlapped chunks.  obs_dev of orl_reset/orl_step must be 16-byte aligned. */
int orl_reset_host(orl_env* env, const uint8_t* mask, float* obs);
int orl_step_host(orl_env* env, const float* actions, float* obs, float* reward, uint8_t* done);

/* wrapper_env.py:151-159 (curriculum) and imitation_task.py:337-339 */
int orl_set_max_episode_steps(orl_env* env, int32_t steps);
int orl_set_ref_state_init_prob(orl_env* env, double prob);

/* ---- state access ------------------------------------------------------------------------------ */
int orl_state_layout_get(const orl_env* env, orl_state_layout* out);
int orl_state_ptr(orl_env* env, float** state_dev);               /* [N][stride] floats, device */
int orl_get_state_host(orl_env* env, float* state);               /* copies [N][stride] to the host */
int orl_set_state_host(orl_env* env, const float* state);
/* [sum reward, sum finished-episode length, finished episodes, robot steps] since the last call
 * with reset_after != 0.  Feeds the rollout-statistic all-reduce. */
int orl_stats(orl_env* env, double out[4], int32_t reset_after);

/* ---- parity / diagnostics entry points (same device functions as the fused step) ------------- */
/* motion_data.py:682-718, 234-253, 417-476 + imitation_task.py:920-983 without origin offset:
 * for each i: clip_dev[i], time_dev[i] -> f0,f1,cycle (int32), blend (f64), pose [19], vel [18]. */
int orl_motion_lookup(orl_env* env, const int32_t* clip_dev, const double* time_dev, int32_t n,
                      int32_t* f0_dev, int32_t* f1_dev, int32_t* cycle_dev, double* blend_dev,
                      float* pose_dev, float* vel_dev, void* stream);
/* pybullet getLinkState(body, i)[0] for every link i: pose_dev [n][19] -> com_dev [n][n_links][3] */
int orl_fk(orl_env* env, const float* pose_dev, int32_t n, float* com_dev, void* stream);
/* Joint-space inertia matrix (CRBA), generalised-velocity order [root lin(3), root ang(3), joints(12)]:
 * pose_dev [n][19] -> M_dev [n][18][18].  Oracle: pybullet.calculateMassMatrix (north_star). */
/* Observation history behind the randomised control latency (minitaur.py:313-357), for tests and checkpoints:
 * ring_q [n][ORL_LAT_RING][12] true motor angles, ring_b [n][ORL_LAT_RING][8] (orientation rel. to init (4), body-frame
 * angular velocity (3), pad); entry E of a robot sits in slot E % ORL_LAT_RING, E = substep counter + 1.  Either may be NULL. */
#define ORL_LAT_RING 42
int orl_get_latency_rings_host(orl_env* env, float* ring_q, float* ring_b);
int orl_set_latency_rings_host(orl_env* env, const float* ring_q, const float* ring_b);
/* MotorModel.convert_to_torque (R/envs/quadruped_robot/robots/minitaur_motor.py:104-173) for n robots, all three
 * control modes.  The env path itself is POSITION only, as in the reference (minitaur.py:113 hard-codes it and
 * Minitaur.step never passes another mode); this entry point serves callers of the motor model.
 *   mode ORL_MOTOR_POSITION: commands [n][12] desired motor angles
 *        ORL_MOTOR_TORQUE  : commands [n][12] torques (scaled by the strength ratios, no clipping: :137-140)
 *        ORL_MOTOR_HYBRID  : commands [n][60], per motor (position, kp, velocity, kd, torque) (:153-162, 25-31)
 * motor_angle / motor_velocity [n][12] in motor space; kp / kd [12] (ignored in HYBRID / TORQUE); strength [n][12] or NULL
 * (= 1); torque_limits [12] or NULL (no clipping, :167-171).  All pointers are device pointers.  out [n][12]. */
enum { ORL_MOTOR_POSITION = 1, ORL_MOTOR_TORQUE = 2, ORL_MOTOR_HYBRID = 3 };
int orl_motor_convert_to_torque(int32_t mode, const float* commands_dev, const float* motor_angle_dev,
                                const float* motor_velocity_dev, const float* kp_dev, const float* kd_dev,
                                const float* strength_dev, const float* torque_limits_dev, int32_t n,
                                float* out_dev, void* stream);

/* Randomizer values [n][12] (layout of orl_state_layout.rand_params) used by the three dynamics entry points below;
 * NULL (default) = the model's own masses, inertias and friction. */
int orl_set_dynamics_rand(orl_env* env, const float* rand_dev);
int orl_mass_matrix(orl_env* env, const float* pose_dev, int32_t n, float* M_dev, void* stream);
/* Inverse dynamics (RNEA), no contact: pose [n][19], vel [n][18], acc [n][18] (classical time derivative
 * of vel) -> generalised forces tau [n][18].  Oracle: pybullet.calculateInverseDynamics (north_star). */
int orl_inverse_dynamics(orl_env* env, const float* pose_dev, const float* vel_dev, const float* acc_dev,
                         int32_t n, float* tau_dev, void* stream);
/* One physics substep's acceleration incl. the contact model: pose, vel, joint torque [n][12] (URDF joint
 * space) -> acc [n][18].  Replaces the arithmetic of pybullet.stepSimulation (quadruped_gym_env.py:223). */
int orl_forward_dynamics(orl_env* env, const float* pose_dev, const float* vel_dev, const float* tau_dev,
                         int32_t n, float* acc_dev, void* stream);
/* The five reward terms of the CURRENT state (imitation_task.py:358-516): terms_dev [N][5]. */
int orl_reward_terms(orl_env* env, float* terms_dev, void* stream);
/* Number of kernels this handle has launched since creation (bench.py "gpu_launches"). */
int64_t orl_launch_count(const orl_env* env);
/* ---- next row: rollout post-processing on the device -------------------------- */
/* Per-robot GAE(lambda) over the runner's time-major interleaved buffers (index t*N + j), replacing
 * add_vtarg_and_adv (R/agents/ppo_imitation.py:68-93) and the per-robot advantage standardisation
 * (ppo_imitation.py:329-338).  reference_index_quirk != 0 reproduces the reference's non-terminal lookup
 * episode_starts[(t*N + i) + (1 + i)] (:88) instead of episode_starts[(t+1)*N + i]; out-of-range reads are
 * "not a start" like the reference's [False]*num_robot padding.  standardize != 0 applies (adv - mean_t)/std_t
 * per robot (population std) AFTER tdlamret = adv + vpred has been formed.  No env handle: pure function. */
int orl_gae(const float* rewards_dev, const float* vpred_dev, const float* nextvpred_dev,
            const uint8_t* episode_starts_dev, int32_t T, int32_t N, float gamma, float lam,
            int32_t reference_index_quirk, int32_t standardize, float* adv_dev, float* tdlamret_dev, void* stream);
/* Measurement helper (no reference counterpart): sustained FP32 FFMA throughput of `device` in TFLOP/s,
 * the denominator of the FP32-pipe roofline (the driver measures no FP32 figure). */
int orl_fp32_peak(int32_t device, int32_t repeats, double* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* ORL_H_ */
