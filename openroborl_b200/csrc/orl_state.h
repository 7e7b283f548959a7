// Per-robot state record: one contiguous row of ORL_STRIDE floats per robot in HBM ([N][ORL_STRIDE]).
// A warp owns 8 consecutive robots = one contiguous 8*ORL_STRIDE*4-byte chunk, moved to and from shared memory
// with one TMA bulk copy each way (DESIGN.md "data layout").
#pragma once

// dynamics state (R/envs/quadruped_robot/task/imitation_task.py:795-829 pose/vel layout)
#define S_POS 0          // 3  base COM position (robot-local world: xy origin = reset slot)
#define S_QUAT 3         // 4  base orientation xyzw
#define S_Q 7            // 12 joint angles, URDF order
#define S_LIN 19         // 3  base linear velocity (world)
#define S_ANG 22         // 3  base angular velocity (world)
#define S_QD 25          // 12 joint velocities
// 37..39 pad
// action path (R/envs/utilities/action_filter.py:95-127, minitaur.py:287-293, 336-357)
#define S_FX 40          // 2x12 filter input history, newest first, motor order
#define S_FY 64          // 2x12 filter output history
#define S_PREV 88        // 12 filtered action of the previous control step
#define S_RING1 100      // 12 true motor angles one substep ago
#define S_RING2 112      // 12 true motor angles two substeps ago (= the control observation)
// sensor histories, newest first (R/envs/utilities/sensors/sensor_wrappers.py:122-142)
#define S_IMU 124        // 3x4
#define S_ACT 136        // 3x12
#define S_ANGH 172       // 3x12
// task state (imitation_task.py:126-139, 725-729)
#define S_REFPOSE 208    // 19
#define S_REFVEL 227     // 18
#define S_OPOS 245       // 3 origin offset position
#define S_OROT 248       // 4 origin offset rotation
#define S_TOFF 252       // f64 motion time offset
#define S_PPHASE 254     // f64 previous motion phase
#define S_CLIP 256       // i32
#define S_WARM 257       // i32
#define S_SUBSTEP 258    // i32 Minitaur._state_action_counter
#define S_STEP 259       // i32 Minitaur.step_counter
#define S_ENVSTEP 260    // i32 env_step_counter (per robot; all equal without auto-reset)
#define S_EPISODE 261    // i32 episode index (RNG counter)
#define S_EPRET 262      // f32 return of the running episode
#define S_NCHANGE 263    // i32 clip changes of the running episode (option kernels; RNG counter)
#define S_OBSQ 264       // 4 base orientation (relative to the init orientation) of the control observation,
                         //   i.e. the state two substeps before the end of the control step (minitaur.py:336-357)
#define S_STRENGTH 268   // 12 motor strength ratios (minitaur_motor.py:68-75)
// domain randomizer (controllable_env_randomizer_from_config.py; values the robot setters receive)
#define S_RAND 280       // 12: [0] base mass ratio [1] leg mass ratio [2] base inertia ratio [3] leg inertia ratio
                         //     [4] control latency (s; sampled, not applied yet) [5] foot lateral friction x plane friction
                         //     (< 0: use the model's) [6..9] knee Coulomb friction torque per leg
#define S_CHANGET 290    // f64 motion time of the next clip change (imitation_task.py:1057-1067; +inf: none)
#define ORL_STRIDE 292   // multiple of 4 (float4) and of 2 (f64 fields stay 8-byte aligned)
#define ORL_ROBOTS_PER_WARP 8
