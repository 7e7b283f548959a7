// The fused control-step kernel, the reset kernel and the parity entry-point kernels.
//
// Thread mapping: one lane per leg, 4 lanes (a "quad") per robot, 8 robots per warp, 4 independent warps per
// thread block (they share one copy of the packed model).  The 8 robots of a warp are one contiguous chunk of the
// [N][ORL_STRIDE] state array: it is fetched into shared memory with one TMA bulk copy, the 33 physics substeps run
// out of registers (plus a lane-private scratch column in shared memory), and the record plus the 160-float
// observation rows leave through bulk copies again.
#pragma once
#include "orl_math.cuh"
#include "orl_motion.cuh"
#include "orl_dynamics.cuh"
#include "orl_state.h"
#define S_OBSW 37           // scratch in the record's padding: body-frame angular velocity of the control observation

struct DevCfg {
  int n;                      // robots on this device
  int gid0;
  unsigned long long seed;
  int action_repeat;
  int max_episode_steps;
  int auto_reset;
  int cycle_sync;
  double sim_dt, env_dt, warmup_time, rsi_prob;
  int tar_steps[ORL_NUM_TAR_FRAMES];
  float w_pose, w_vel, w_ee, w_root, w_rootvel;
  float s_pose, s_vel, s_ee, s_eeh, s_root, s_rootvel;
  float dist_thr2, rot_thr;
  float fb0, fb1, fb2, fa1, fa2;   // Butterworth (action_filter.py:129-217)
  PhysParams phys;
  int randomize;                   // domain randomizer on reset (minitaur.py:273-276)
  float mu_plane;
  float rand_lo[8], rand_hi[8];    // ranges in the reference's sorted parameter order (ORL_RAND_*)
  // Task options run.py leaves off (imitation_task.py:36-57): only read by the OPT instantiations of the kernels
  float reward_weight;
  int rand_init_time;
  double clip_tmin, clip_tmax, perturb_prob;
  int tar_noise_on; float tar_noise;
};

struct DevEnv {
  float* state;                       // [n][ORL_STRIDE]
  const DevModel* model;              // device copy (packed, float4 aligned)
  MotionTables mt;
  DevCfg cfg;
  double* stats;                      // [4]
  // Randomised control latency (minitaur.py:336-357 with latency in [0, 0.04] s): the last ORL_LAT_RING true
  // observations of every robot, entry E in slot E % ORL_LAT_RING.  Only allocated when the randomizer is on.
  float* ring_q;                      // [n + 1][ORL_LAT_RING][12] true motor angles (row n: scratch for padding quads)
  float* ring_b;                      // [n + 1][ORL_LAT_RING][8]  base orientation rel. to the init orientation (4), body-frame angular velocity (3)
  // Episodes prepared ahead of time (auto-reset): the record a robot will start its NEXT episode from only depends on
  // (seed, global robot id, episode index), so a few extra thread blocks of the step kernel compute it while the other
  // blocks are busy with the substep loop; the in-kernel reset then copies 292 floats instead of running ~2000
  // instructions of cold code at the very end of the slowest warps (profiles/r02_step_kernel.md "tail").
  float* next_ep;                     // [n][ORL_STRIDE] record of episode `prepared[i]` (null: feature off)
  int* prepared;                      // [n] episode index next_ep[i] holds (0: none)
  int prep_blocks;                    // leading thread blocks of a step launch that prepare episodes
  // Who needs a new record: every reset appends its robot to the list the NEXT launch consumes (two lists, `prep_parity`
  // = the one consumed by this launch); prep_scan = 1 makes the prepare blocks walk the whole batch instead (after a
  // restore or a configuration change, when the lists say nothing).
  int* todo;                          // [2][2 n] robot indices
  int* todo_count;                    // [2]
  int prep_parity, prep_scan;
  // 1: the observation leaves in pieces (the history part right after the action filter, the rest at the end); set when the
  // observation buffer is mapped HOST memory, where it takes 1.1 of the 2.6 MB per step off the PCIe tail of the launch.
  // 0: one bulk copy of whole rows per warp at the end (device buffers: six small copies per robot only cost TMA issue slots).
  int obs_split;
  // Reference packets computed one launch ahead (profiles/r02_step_kernel.md section 9): everything the once-per-step code
  // reads from the motion clips -- the reference sample at the end of a step and the four target frames of its observation
  // -- only depends on (clip, warm-up flag, substep counter, time offset).  Unless a robot is reset, the record of the NEXT
  // launch has the same clip / flag / offset and a counter that is one control step further, so the leading blocks of this
  // launch compute the NEXT launch's row of every robot while the workers are in the substep loop (a robot that is reset
  // gets the row of its new episode's first step from the prepared episode instead).  The workers copy 480 bytes per
  // robot instead of running five FP64 index computations, five quaternion slerps and 72 joint blends of cold code.
  // [3][8 ceil(n / 8)][ORL_PK_STRIDE]: part `ahead_parity` (0 / 1) is read by this launch, the other one written by its
  // leading blocks; part 2 holds the first-step rows of freshly started episodes, written by the reset (sole writer per
  // robot, so a late leading block can never mix its bytes into such a row) and read while the substep counter is 0
  float* ahead;                       // (null: off)
  int ahead_parity;
  int ahead_rows;                     // rows per half
};
// ORL_LAT_RING (orl.h) = 42: latency <= 0.04 s = 40 substeps, entries E-41 .. E are live

// Thread block = ORL_WARPS_PER_BLOCK independent warps sharing one copy of the packed model in shared memory.
// (Running them in lockstep with a barrier per substep, so that they would share fetched instruction lines, was
// measured and brought nothing: profiles/r01_v4_step_kernel.md; -DORL_LOCKSTEP re-enables it.)
#ifndef ORL_WARPS_PER_BLOCK
#define ORL_WARPS_PER_BLOCK 4
#endif
#define ORL_BLOCK_THREADS (32 * ORL_WARPS_PER_BLOCK)
#define ORL_ROBOTS_PER_BLOCK (ORL_ROBOTS_PER_WARP * ORL_WARPS_PER_BLOCK)
#define ORL_SMEM_MODEL_BYTES ((sizeof(DevModel) + 15) / 16 * 16)
#define ORL_SMEM_OBS_BYTES (ORL_ROBOTS_PER_WARP * ORL_OBS_SIZE * 4)
#define ORL_SMEM_X_BYTES (ORL_X_FLOAT4_PER_LANE * ORL_X_STRIDE * 16)
#define ORL_SMEM_WARP_BYTES (ORL_ROBOTS_PER_WARP * ORL_STRIDE * 4 + (ORL_SMEM_X_BYTES > ORL_SMEM_OBS_BYTES ? ORL_SMEM_X_BYTES : ORL_SMEM_OBS_BYTES))
#define ORL_SMEM_BAR_OFFSET (ORL_SMEM_MODEL_BYTES + ORL_WARPS_PER_BLOCK * ORL_SMEM_WARP_BYTES)
#define ORL_SMEM_BYTES (ORL_SMEM_BAR_OFFSET + 8 * (ORL_WARPS_PER_BLOCK + 1) + 8)      // + one mbarrier per warp + one for the model

ORL_DI int& as_int(float& f) { return *reinterpret_cast<int*>(&f); }
ORL_DI double ld_f64(const float* p) { return *reinterpret_cast<const double*>(p); }
ORL_DI void st_f64(float* p, double v) { *reinterpret_cast<double*>(p) = v; }

// ---- shared-memory staging ---------------------------------------------------------------------
// One thread block = one warp = 8 robots.  The packed model (same bytes for every block, L2 resident) and the
// warp's 8 contiguous state records are fetched with two TMA bulk copies (cp.async.bulk, SASS UBLKCP) that
// complete on one mbarrier: one round trip to L2/HBM instead of ~30 dependent LDG->STS iterations per lane.
ORL_DI uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
ORL_DI void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
ORL_DI void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src), "r"(bytes) : "memory");
}
ORL_DI void mbar_wait(uint32_t bar) {
  uint32_t ready = 0;
  while (!ready)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ready) : "r"(bar) : "memory");
}
// First robot row of this warp and how many of its 8 rows exist (<= 0 for a warp past the end of the batch: it leaves
// after the block barrier of stage_in; idle quads of a partly filled warp work on a copy of its last record).
ORL_DI int warp_row0(const DevEnv& E) { return (((int)blockIdx.x - E.prep_blocks) * ORL_WARPS_PER_BLOCK + (threadIdx.x >> 5)) * ORL_ROBOTS_PER_WARP; }
ORL_DI bool stage_in(const DevEnv& E, unsigned char* smem, const DevModel*& Mo, float*& st, float*& ob) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Mo = reinterpret_cast<const DevModel*>(smem);
  st = reinterpret_cast<float*>(smem + ORL_SMEM_MODEL_BYTES + warp * ORL_SMEM_WARP_BYTES);
  ob = st + ORL_ROBOTS_PER_WARP * ORL_STRIDE;
  const int row0 = warp_row0(E);
  const int nrow = min(ORL_ROBOTS_PER_WARP, E.cfg.n - row0);
  const int src0 = nrow > 0 ? row0 : E.cfg.n - 1, nsrc = nrow > 0 ? nrow : 1;
  const uint32_t bar_model = smem_addr(smem + ORL_SMEM_BAR_OFFSET), bar = bar_model + 8u * (warp + 1);
  if (threadIdx.x == 0) {
    for (int i = 0; i <= ORL_WARPS_PER_BLOCK; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_model + 8u * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
#ifndef ORL_LOCKSTEP
  if (nrow <= 0) return false;           // a warp past the end of the batch: nothing to do (no block barrier follows)
#endif
  if (lane == 0) {
    const uint32_t sbytes = (uint32_t)nsrc * ORL_STRIDE * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(sbytes) : "memory");
    bulk_g2s(smem_addr(st), E.state + (size_t)src0 * ORL_STRIDE, sbytes, bar);
    if (warp == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_model), "r"((uint32_t)ORL_SMEM_MODEL_BYTES) : "memory");
      bulk_g2s(smem_addr(smem), E.model, (uint32_t)ORL_SMEM_MODEL_BYTES, bar_model);
    }
  }
  __syncwarp();
  mbar_wait(bar);
  mbar_wait(bar_model);
  if (nsrc < ORL_ROBOTS_PER_WARP) {        // ragged end: idle quads work on a copy of the last valid record
    // (a rolled float4 copy per idle quad: the flat loop with its `% ORL_STRIDE` unrolled into 300 instructions in front
    // of the substep loop, and code in front of that loop is not free: profiles/r02_step_kernel.md "code placement")
    if ((lane >> 2) >= nsrc) {
      const float4* src = reinterpret_cast<const float4*>(st + (nsrc - 1) * ORL_STRIDE);
      float4* dst = reinterpret_cast<float4*>(st + (lane >> 2) * ORL_STRIDE);
#pragma unroll 1
      for (int i = lane & 3; i < ORL_STRIDE / 4; i += 4) dst[i] = src[i];
    }
    __syncwarp();
  }
  return true;
}
ORL_DI void stage_out(const DevEnv& E, const float* st, const float* ob, float* obs_out) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the bulk copy
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    const int row0 = warp_row0(E);
    const int nrow = min(ORL_ROBOTS_PER_WARP, E.cfg.n - row0);
    if (nrow > 0) {
      bulk_s2g(E.state + (size_t)row0 * ORL_STRIDE, smem_addr(st), (uint32_t)nrow * ORL_STRIDE * 4u);
      if (obs_out) bulk_s2g(obs_out + (size_t)row0 * ORL_OBS_SIZE, smem_addr(ob), (uint32_t)nrow * ORL_OBS_SIZE * 4u);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
}

// ---- helpers on the staged record ----------------------------------------------------------------
ORL_DI float3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
ORL_DI Q4 ldq(const float* p) { return q4(p[0], p[1], p[2], p[3]); }
ORL_DI void st3(float* p, float3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
ORL_DI void stq(float* p, Q4 q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }

ORL_DI RefFrame load_ref_frame(float* S, const DevCfg& C) {
  RefFrame F;
  F.clip = as_int(S[S_CLIP]);
  F.warmup = as_int(S[S_WARM]) != 0;
  F.warmup_time = C.warmup_time;
  F.origin_rot = ldq(S + S_OROT);
  F.origin_pos = ld3(S + S_OPOS);
  return F;
}
// ImitationTask._get_motion_time (imitation_task.py:831-848); episode_start_time_offset is -0.0
ORL_DI double motion_time(float* S, const DevCfg& C) {
  double t = __dmul_rn((double)as_int(S[S_SUBSTEP]), C.sim_dt);
  t = __dadd_rn(t, ld_f64(S + S_TOFF));
  t = __dadd_rn(t, -0.0);
  if (as_int(S[S_WARM])) t = __dsub_rn(t, C.warmup_time);
  return t;
}
ORL_DI Q4 ldq4(float4 v) { return q4(v.x, v.y, v.z, v.w); }

// ---- random draws of the task options: further blocks of the (robot, episode) Philox stream (oracle/philox.py) ------
#define ORL_OPT_BLOCK 9u
#define ORL_NORMAL_BLOCK0 10u
#define ORL_AXIS_BLOCK 26u
#define ORL_CHANGE_BLOCK0 32u
#define ORL_NOISE_BLOCK0 0x40000000u
ORL_DI void opt_block(const DevCfg& C, int gid, int episode, uint32_t ctr, uint32_t w[4]) {
  philox4x32((uint32_t)gid, (uint32_t)episode, ctr, 0u, (uint32_t)(C.seed & 0xffffffffull), (uint32_t)(C.seed >> 32), w);
}
ORL_DI double u01_32(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }
// Box-Muller in FP64 (rare paths only): sqrt(-2 ln u1) cos(2 pi u2)
ORL_DI double normal_f64(uint32_t w0, uint32_t w1) {
  return sqrt(-2.0 * log(u01_32(w0))) * cos(2.0 * 3.141592653589793 * u01_32(w1));
}
// standard normal number i of the initial-state perturbation of (robot, episode)
ORL_DI double perturb_normal(const DevCfg& C, int gid, int episode, int i) {
  uint32_t w[4];
  opt_block(C, gid, episode, ORL_NORMAL_BLOCK0 + (uint32_t)(i >> 1), w);
  return (i & 1) ? normal_f64(w[2], w[3]) : normal_f64(w[0], w[1]);
}

// ---- reference packets ------------------------------------------------------------------------------
// One row per robot; root pose / velocity BEFORE the per-episode origin (the cycle-sync re-anchoring of the update moves
// the origin with the simulated robot, so the origin is applied by the consumer), joints, the phase / motion-over
// results of the FP64 clock, the four target frames (root before the origin, joints), and the key the row was computed
// from.  The same two producers run in the leading blocks (into global memory), in the blocks that prepare episodes
// (target frames of the first observation of the next episode) and, when a row is missing, inline in the worker (into
// the shared-memory copy of the row): one arithmetic, whoever runs it.
#define PK_ROOT 0        // 3 + 4
#define PK_VEL 7         // 3 + 3
#define PK_J 13          // 12
#define PK_JV 25         // 12
#define PK_OVER 37       // i32 motion_is_over
#define PK_PHASE 38      // f64
#define PK_TAR 40        // 4 x 19
#define PK_KEY 116       // i32 substep counter the row is for, i32 2 clip + warm, f64 time offset
#define ORL_PK_STRIDE 120
#define ORL_PK_TAR_FLOATS (19 * ORL_NUM_TAR_FRAMES)
#define ORL_NEXT_STRIDE (ORL_STRIDE + ORL_PK_TAR_FLOATS + ORL_PK_STRIDE)   // a prepared episode: its record, the target frames of its first observation, the packet row of its first step
static_assert(ORL_PK_STRIDE % 4 == 0 && PK_TAR % 4 == 0 && PK_PHASE % 2 == 0 && ORL_NEXT_STRIDE % 4 == 0, "packet rows are copied as float4");
// ImitationTask._get_motion_time for a given substep counter (imitation_task.py:831-848)
ORL_DI double motion_time_at(int substep, double toff, bool warm, const DevCfg& C) {
  double t = __dmul_rn((double)substep, C.sim_dt);
  t = __dadd_rn(t, toff);
  t = __dadd_rn(t, -0.0);
  if (warm) t = __dsub_rn(t, C.warmup_time);
  return t;
}
// reference sample of task.update at `time`: every lane of the quad evaluates the sample, lane 0 stores the root part
ORL_DI void produce_update(const MotionTables& T, const RefFrame& F, double time, float* pk, int leg) {
  const RefSample smp = ref_sample(T, F, time);
  float3 p, l, a; Q4 r;
  ref_root_raw(T, F, smp, p, r);
  ref_root_vel_raw(T, F, smp, l, a);
  const double ph = motion_phase(T, F.clip, time);
  if (leg == 0) {
    st3(pk + PK_ROOT, p); stq(pk + PK_ROOT + 3, r); st3(pk + PK_VEL, l); st3(pk + PK_VEL + 3, a);
    *reinterpret_cast<int*>(pk + PK_OVER) = motion_is_over(T, F.clip, time) ? 1 : 0;
    *reinterpret_cast<double*>(pk + PK_PHASE) = ph;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    pk[PK_J + 3 * leg + k] = ref_joint(T, F, smp, 3 * leg + k);
    pk[PK_JV + 3 * leg + k] = ref_joint_vel(T, F, smp, 3 * leg + k);
  }
}
// target frames of the observation built at clock value t0 (imitation_task.py:254-301): lane j evaluates frame j
ORL_DI void produce_targets(const MotionTables& T, const RefFrame& F, double t0, const DevCfg& C, float* tar, int leg) {
  const double tt = __dadd_rn(t0, __dmul_rn((double)C.tar_steps[leg], C.env_dt));
  const RefSample s = ref_sample(T, F, tt);
  float3 p; Q4 r;
  ref_root_raw(T, F, s, p, r);
  float* o = tar + 19 * leg;
  st3(o, p); stq(o + 3, r);
#pragma unroll 4
  for (int j = 0; j < 12; ++j) o[7 + j] = ref_joint(T, F, s, j);
}
// One complete row for the step that ENDS with substep counter `substep1` (what reference_ahead, prepare_episodes and --
// through the two producers -- the inline fallback of the worker compute).
ORL_DI void produce_row(const MotionTables& T, const DevCfg& C, int clip, bool warm, double toff, int substep1, float* pk, int leg) {
  RefFrame F;
  F.clip = clip; F.warmup = warm; F.warmup_time = C.warmup_time;
  F.origin_rot = q4(0.f, 0.f, 0.f, 1.f); F.origin_pos = v3(0.f, 0.f, 0.f);
  const double time = motion_time_at(substep1, toff, warm, C);
  produce_update(T, F, time, pk, leg);
  produce_targets(T, F, time, C, pk + PK_TAR, leg);
  if (leg == 0) {
    *reinterpret_cast<int*>(pk + PK_KEY) = substep1; *reinterpret_cast<int*>(pk + PK_KEY + 1) = 2 * clip + (warm ? 1 : 0);
    *reinterpret_cast<double*>(pk + PK_KEY + 2) = toff;
  }
}
// shared-memory home of the warp's eight packet rows: behind the observation staging, inside the dynamics scratch (dead
// once the substep loop is over)
ORL_DI float* pk_rows(float* ob) { return ob + ORL_ROBOTS_PER_WARP * ORL_OBS_SIZE; }
ORL_DI void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

// Observation rows: sensor histories (already in the record) + the target frames.
// ImitationTask.build_target_obs (imitation_task.py:254-301): lane j of the quad builds frame j.
// `tar`: this robot's target-frame block of its packet row; have_tar: it already holds the frames of the record's clock.
template <bool OPT>
ORL_DI void build_obs(float* S, float* O, const DevEnv& E, int leg, int gid, float* tar, bool have_tar) {
  const DevCfg& C = E.cfg;
  static_assert(S_IMU % 4 == 0 && ORL_STRIDE % 4 == 0 && ORL_OBS_SIZE % 4 == 0, "history block must be float4 aligned");
#pragma unroll 1
  for (int i = leg; i < 21; i += 4)           // 84 floats of sensor history: the quad copies 21 float4
    reinterpret_cast<float4*>(O)[i] = reinterpret_cast<const float4*>(S + S_IMU)[i];
  const RefFrame F = load_ref_frame(S, C);
  if (!have_tar) produce_targets(E.mt, F, motion_time(S, C), C, tar, leg);
  __syncwarp();
  const float3 ref_base = ld3(S + S_REFPOSE);
  // robot.get_base_orientation(): delayed orientation -> euler -> quaternion (minitaur.py:513-527, 630-638)
  // Away from the gimbal branch of Bullet's ZYX euler extraction the heading of that round trip IS the yaw of the
  // original quaternion (rotated x axis = Rz(yaw) Ry(pitch) e_x, cos(pitch) > 0), so no trigonometry is needed.
  const Q4 oq = ldq(S + S_OBSQ);
  float sh, ch, shh, chh;                     // sin / cos of -heading and of -heading / 2 (imitation_task.py:270-274)
  if (fabsf(-2.f * (oq.x * oq.z - oq.w * oq.y)) < 0.99999f) {
    qheading_neg_sincos(oq, sh, ch);
  } else {
    float roll, pitch, yaw;
    bullet_euler(oq, roll, pitch, yaw);
    qheading_neg_sincos(bullet_quat_from_euler(roll, pitch, yaw), sh, ch);
  }
  if (OPT && C.tar_noise_on) {                // heading += randn(0, tar_obs_noise[0]) (imitation_task.py:274-275)
    uint32_t w[4];
    opt_block(C, gid, as_int(S[S_EPISODE]) - 1, ORL_NOISE_BLOCK0 + (uint32_t)as_int(S[S_ENVSTEP]), w);
    const double nz = (double)C.tar_noise * normal_f64(w[0], w[1]) + 0.0;
    float sn, cn;
    sincos_f64arg(nz, sn, cn);                // (sh, ch) = sin / cos of -heading: rotate by -noise
    const float s2 = sh * cn - ch * sn, c2 = ch * cn + sh * sn;
    sh = s2; ch = c2;
  }
  half_angle(sh, ch, shh, chh);
  const Q4 inv_rot = q4(0.f, 0.f, shh, chh);
  const float* t = tar + 19 * leg;
  float3 p = ld3(t); Q4 r = ldq(t + 3);
  ref_apply_origin(F, p, r);
  p = qrot_z(sh, ch, p - ref_base);
  r = qstd(qmul(inv_rot, r));
  float* o = O + 84 + 19 * leg;
  st3(o, p); stq(o + 3, r);
#pragma unroll
  for (int j = 0; j < 12; ++j) o[7 + j] = t[7 + j];
}

// Episode reset of one robot, executed by its quad on the staged record.
// Minitaur.reset (minitaur.py:232-278) + ImitationTask.reset/_reset_ref_motion/_sync_sim_model
// (imitation_task.py:166-199, 694-732, 778-793).  `qmask` = the 4 lanes of this quad.
template <bool OPT>
ORL_DI void reset_robot(float* S, const DevEnv& E, const DevModel& Mo, int gid, int leg, unsigned qmask) {
  const DevCfg& C = E.cfg;
  const DevLeg& L = Mo.leg[leg];
  const int episode = as_int(S[S_EPISODE]);
  uint32_t ra[4], rb[4];
  philox4x32((uint32_t)gid, (uint32_t)episode, 0u, 0u, (uint32_t)(C.seed & 0xffffffffull), (uint32_t)(C.seed >> 32), ra);
  philox4x32((uint32_t)gid, (uint32_t)episode, 1u, 0u, (uint32_t)(C.seed & 0xffffffffull), (uint32_t)(C.seed >> 32), rb);
  const double u_rsi = u01_53(ra[0], ra[1]);
  const int clip = (int)(((unsigned long long)ra[2] * (unsigned long long)E.mt.n_clips) >> 32);
  const double u_time = u01_53(rb[0], rb[1]);
  const bool ref_init = u_rsi < C.rsi_prob;                          // imitation_task.py:183-184
  const bool warm = !ref_init && C.warmup_time > 0.0;                 // :186-188
  double toff = warm ? __dadd_rn(0.0, __dmul_rn(__dsub_rn(C.warmup_time, 0.0), u_time))
                     : __dadd_rn(0.0, __dmul_rn(__dsub_rn(E.mt.dur[clip], 0.0), u_time));   // :1103-1123
  if (OPT && !C.rand_init_time) toff = 0.0;                          // :1104-1105
  __syncwarp(qmask);
  // per-motor fields owned by this lane
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int m = __float_as_int(L.j[k].m1.z);
    const float qm_def = L.j[k].m1.y * L.j[k].m0.y;                 // true motor angle of the default pose (INIT * DIR)
    S[S_FX + m] = 0.f; S[S_FX + 12 + m] = 0.f; S[S_FY + m] = 0.f; S[S_FY + 12 + m] = 0.f;   // action_filter.py:103-109
    S[S_PREV + m] = 0.f;
    S[S_RING1 + m] = qm_def; S[S_RING2 + m] = qm_def;
#pragma unroll
    for (int h = 0; h < 3; ++h) { S[S_ACT + 12 * h + m] = 0.f; S[S_ANGH + 12 * h + m] = wrap_pi(qm_def); }
  }
  if (leg == 0) {
    for (int i = 0; i < 12; ++i) S[S_IMU + i] = 0.f;
    as_int(S[S_CLIP]) = clip;
    as_int(S[S_WARM]) = warm ? 1 : 0;
    as_int(S[S_SUBSTEP]) = 0; as_int(S[S_STEP]) = 0; as_int(S[S_ENVSTEP]) = 0;
    as_int(S[S_EPISODE]) = episode + 1;
    S[S_EPRET] = 0.f;
    st_f64(S + S_TOFF, toff);
    stq(S + S_OBSQ, q4(0.f, 0.f, 0.f, 1.f));
    st3(S + S_OBSW, v3(0.f, 0.f, 0.f));          // scratch; written so that a record never carries bytes of an older episode
  }
  // domain randomizer (Minitaur.reset -> randomize_env, minitaur.py:273-276): ORL_RAND_DRAWS uniforms of this
  // (robot, episode), consumed in the reference's sorted parameter order; u = (word + 0.5) / 2^32
  if (C.randomize) {
    // No arrays: every draw is mapped and stored where it belongs as soon as its Philox block exists (a value array
    // indexed by the block counter lived in local memory: 288 bytes of stack in the step kernel).
    int mot[3], fdraw[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { mot[k] = __float_as_int(L.j[k].m1.z); fdraw[k] = __float_as_int(L.j[k].i3.w); }
    if (fdraw[0] < 0 && fdraw[1] < 0 && fdraw[2] < 0) S[S_RAND + 6 + leg] = 0.f;
#pragma unroll 1
    for (int b = 0; b < ORL_RAND_DRAWS / 4; ++b) {
      uint32_t w[4];
      philox4x32((uint32_t)gid, (uint32_t)episode, 2u + (uint32_t)b, 0u, (uint32_t)(C.seed & 0xffffffffull), (uint32_t)(C.seed >> 32), w);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = 4 * b + j;                  // draw index; groups start at 0, 1, 3, 11, 12, 13, 15, 16 (sorted parameter names)
        const int p = (i >= 1) + (i >= 3) + (i >= 11) + (i >= 12) + (i >= 13) + (i >= 15) + (i >= 16);
        const double u = ((double)w[j] + 0.5) * (1.0 / 4294967296.0);
        const double smp = -1.0 + 2.0 * u;                                              // np.random.uniform(-1, 1)
        const float v = (float)((smp - (-1.0)) / 2.0 * ((double)C.rand_hi[p] - (double)C.rand_lo[p]) + (double)C.rand_lo[p]);
        if (i >= 16) {                                                                  // motor strength of motor i - 16
          const int m = i - 16;
          if (m == mot[0] || m == mot[1] || m == mot[2]) S[S_STRENGTH + m] = v;
        } else if (i >= 3 && i < 11) {                                                  // joint friction draw i - 3
          if (fdraw[0] == i - 3 || fdraw[1] == i - 3 || fdraw[2] == i - 3) S[S_RAND + 6 + leg] = v;
        } else if (leg == 0) {
          // mass: base, legs (13, 14) -> 0, 1; inertia: base, legs (1, 2) -> 2, 3; latency (11) -> 4;
          // lateral friction of the foot links (12) -> 5, times the plane's; battery (0) and motor damping (15): ignored
          if (i == 13 || i == 14) S[S_RAND + (i - 13)] = v;
          else if (i == 1 || i == 2) S[S_RAND + 1 + i] = v;
          else if (i == 11) S[S_RAND + 4] = v;
          else if (i == 12) S[S_RAND + 5] = v * C.mu_plane;
        }
      }
    }
  }
  // _reset_ref_motion
  RefFrame F;
  F.clip = clip; F.warmup = warm; F.warmup_time = C.warmup_time;
  F.origin_rot = q4(0.f, 0.f, 0.f, 1.f); F.origin_pos = v3(0.f, 0.f, 0.f);
  double time = __dadd_rn(__dadd_rn(__dmul_rn(0.0, C.sim_dt), toff), -0.0);
  if (warm) time = __dsub_rn(time, C.warmup_time);
  const RefSample s = ref_sample(E.mt, F, time);
  float3 p0; Q4 r0;
  ref_root_raw(E.mt, F, s, p0, r0);               // the origin is still the identity here
  const float3 sim_pos = f4xyz(Mo.ipos);
  const Q4 sim_rot = ldq4(Mo.iquat);
  float3 op = sim_pos - p0; op.z = 0.f;
  const float dh = qheading(sim_rot) - qheading(r0);
  F.origin_pos = op;
  F.origin_rot = qaxis_z(dh);
  float3 rp = p0, rl, ra3; Q4 rq = r0;
  ref_apply_origin(F, rp, rq);
  ref_root_vel(E.mt, F, s, rl, ra3);
  if (leg == 0) {
    st3(S + S_OPOS, F.origin_pos); stq(S + S_OROT, F.origin_rot);
    st3(S + S_REFPOSE, rp); stq(S + S_REFPOSE + 3, rq);
    st3(S + S_REFVEL, rl); st3(S + S_REFVEL + 3, ra3);
    st_f64(S + S_PPHASE, motion_phase(E.mt, clip, time));
  }
  // options: clip timer (_reset_clip_change_time, :1057-1067, last call of _reset_ref_motion) and the perturbed start
  // state (:192-195, 1199-1243) -- the reference pose / velocity stay unperturbed, only what _sync_sim_model sets moves
  bool perturb = false;
  if (OPT) {
    uint32_t w[4];
    opt_block(C, gid, episode, ORL_OPT_BLOCK, w);
    const double u_perturb = u01_53(w[0], w[1]), u_dur = u01_53(w[2], w[3]);
    const bool switching = isfinite(C.clip_tmin) && isfinite(C.clip_tmax);
    const double change_t = switching ? __dadd_rn(time, __dadd_rn(C.clip_tmin, __dmul_rn(__dsub_rn(C.clip_tmax, C.clip_tmin), u_dur)))
                                      : (double)INFINITY;
    if (leg == 0) { st_f64(S + S_CHANGET, change_t); as_int(S[S_NCHANGE]) = 0; }
    perturb = C.perturb_prob > 0.0 && __dadd_rn(0.0, __dmul_rn(1.0, u_perturb)) < C.perturb_prob;
  }
  if (leg == 0) {
    float3 sp = rp, sl = rl, sa = ra3; Q4 sq = rq;
    if (OPT && perturb) {
      const double pi = 3.141592653589793;
      sp.x = (float)((double)sp.x + (0.025 * perturb_normal(C, gid, episode, 0) + 0.0));
      sp.y = (float)((double)sp.y + (0.025 * perturb_normal(C, gid, episode, 1) + 0.0));
      uint32_t wa[4];
      opt_block(C, gid, episode, ORL_AXIS_BLOCK, wa);
      double ax = -1.0 + 2.0 * u01_32(wa[0]), ay = -1.0 + 2.0 * u01_32(wa[1]), az = -1.0 + 2.0 * u01_32(wa[2]);
      const double inv = 1.0 / sqrt(ax * ax + ay * ay + az * az);
      const double th = 0.025 * pi * perturb_normal(C, gid, episode, 2) + 0.0;
      const double sh = sin(0.5 * th), ch = cos(0.5 * th);
      sq = qmul(q4((float)(ax * inv * sh), (float)(ay * inv * sh), (float)(az * inv * sh), (float)ch), rq);
      sl.x = (float)((double)sl.x + (0.1 * perturb_normal(C, gid, episode, 15) + 0.0));
      sl.y = (float)((double)sl.y + (0.1 * perturb_normal(C, gid, episode, 16) + 0.0));
      sa.x = (float)((double)sa.x + (0.05 * pi * perturb_normal(C, gid, episode, 17) + 0.0));
      sa.y = (float)((double)sa.y + (0.05 * pi * perturb_normal(C, gid, episode, 18) + 0.0));
      sa.z = (float)((double)sa.z + (0.05 * pi * perturb_normal(C, gid, episode, 19) + 0.0));
    }
    // _sync_sim_model
    st3(S + S_POS, sp); stq(S + S_QUAT, sq); st3(S + S_LIN, sl); st3(S + S_ANG, sa);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int j = 3 * leg + k;
    const float qj = ref_joint(E.mt, F, s, j), vj = ref_joint_vel(E.mt, F, s, j);
    S[S_REFPOSE + 7 + j] = qj; S[S_REFVEL + 6 + j] = vj;
    float sqj = qj, svj = vj;
    if (OPT && perturb) {
      const double pi = 3.141592653589793;
      sqj = (float)((double)qj + (0.05 * pi * perturb_normal(C, gid, episode, 3 + j) + 0.0));
      svj = (float)((double)vj + (0.05 * pi * perturb_normal(C, gid, episode, 20 + j) + 0.0));
    }
    S[S_Q + j] = sqj; S[S_QD + j] = svj;
  }
  __syncwarp(qmask);
}
// Second half of the reset in train mode: the observation rings of the randomised control latency restart.
ORL_DI void reset_rings(float* S, const DevEnv& E, const DevModel& Mo, int gid, int leg, unsigned qmask, bool valid) {
  const DevCfg& C = E.cfg;
  const DevLeg& L = Mo.leg[leg];
  if (E.ring_q) {
    // Randomised control latency: the observation history restarts with entry 0 = default pose (receive_obs in
    // Minitaur.reset, minitaur.py:265) and entry 1 = synced pose (receive_obs after _sync_sim_model).
    // padding quads (robot index past the batch) work on a copy of the last record: their ring traffic goes to the scratch row n
    const int ring_row = valid ? gid - C.gid0 : C.n;
    float* rq_ = E.ring_q + (size_t)ring_row * (ORL_LAT_RING * 12);
    const double lat = (double)S[S_RAND + 4] / C.sim_dt;
    const int n = (int)lat;
    const float al = (float)(lat - (double)n);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int m = __float_as_int(L.j[k].m1.z);
      const float qm_def = L.j[k].m1.y * L.j[k].m0.y;
      const float qm_sync = (S[S_Q + 3 * leg + k] - L.j[k].m0.z) * L.j[k].m0.y;
      rq_[m] = qm_def; rq_[12 + m] = qm_sync;
      // control observation the first set_act sees: history = [synced, default]
      S[S_RING2 + m] = n == 0 ? (1.0f - al) * qm_sync + al * qm_def : qm_def;
    }
    if (leg == 0) {
      float4* rb = reinterpret_cast<float4*>(E.ring_b + (size_t)ring_row * (ORL_LAT_RING * 8));
      const Q4 oq = qmul(ldq(S + S_QUAT), qconj(ldq4(Mo.iquat)));
      const float3 ow = qrot(qconj(oq), ld3(S + S_ANG));
      rb[0] = make_float4(0.f, 0.f, 0.f, 1.f); rb[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[2] = make_float4(oq.x, oq.y, oq.z, oq.w); rb[3] = make_float4(ow.x, ow.y, ow.z, 0.f);
    }
    __syncwarp(qmask);
  }
}


// A robot that just started episode e has no prepared record for e + 1 yet: queue it for the next step launch.
ORL_DI void queue_for_preparation(const DevEnv& E, int robot) {
  if (!E.todo) return;
  const int list = E.prep_parity ^ 1;
  const int slot = atomicAdd(E.todo_count + list, 1);
  if (slot < 2 * E.cfg.n) E.todo[(size_t)list * 2 * E.cfg.n + slot] = robot;
}
// The whole in-kernel reset of one robot: from the prepared record when it is there, else computed on the spot.
// Returns true when `tar` (the target-frame block of the robot's packet row) now holds the frames of the new episode's
// first observation (they travel with the prepared record).
template <bool OPT>
ORL_DI bool finish_episode(float* S, const DevEnv& E, const DevModel& Mo, int robot_c, int leg, unsigned qmask, bool valid, float* tar) {
  bool ready = false;
  if (E.next_ep) {
    int have = 0;
    if (leg == 0) asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(have) : "l"(E.prepared + robot_c) : "memory");
    have = __shfl_sync(qmask, have, (threadIdx.x & 31) & ~3);
    ready = have == as_int(S[S_EPISODE]) + 1;
  }
  if (ready) {                                          // (the acquire load of the flag orders the data loads after it)
    const float4* src = reinterpret_cast<const float4*>(E.next_ep + (size_t)robot_c * ORL_NEXT_STRIDE);
    float4* dst = reinterpret_cast<float4*>(S);
    __syncwarp(qmask);
#pragma unroll 1
    for (int i = leg; i < ORL_PK_TAR_FLOATS / 4; i += 4) cp_async16(reinterpret_cast<float4*>(tar) + i, src + ORL_STRIDE / 4 + i);
    asm volatile("cp.async.commit_group;" ::: "memory");
    // all loads in flight before the first store (a rolled load-store loop paid one memory latency per float4: 19 in a row)
    constexpr int kPerLane = (ORL_STRIDE / 4 + 3) / 4;
    float4 v[kPerLane];
#pragma unroll
    for (int k = 0; k < kPerLane; ++k) { const int i = leg + 4 * k; if (i < ORL_STRIDE / 4) v[k] = __ldcg(src + i); }
    // the new episode's first step finds its packet row like every other step: same round trip as the record
    const bool row = E.ahead && valid;
    const float4* rsrc = src + (ORL_STRIDE + ORL_PK_TAR_FLOATS) / 4;
    float4* rdst = reinterpret_cast<float4*>(E.ahead + ((size_t)2 * E.ahead_rows + robot_c) * ORL_PK_STRIDE);
    constexpr int kRow = (ORL_PK_STRIDE / 4 + 3) / 4;
    float4 w[kRow];
#pragma unroll
    for (int k = 0; k < kRow; ++k) { const int i = leg + 4 * k; if (row && i < ORL_PK_STRIDE / 4) w[k] = __ldcg(rsrc + i); }
#pragma unroll
    for (int k = 0; k < kPerLane; ++k) { const int i = leg + 4 * k; if (i < ORL_STRIDE / 4) dst[i] = v[k]; }
#pragma unroll
    for (int k = 0; k < kRow; ++k) { const int i = leg + 4 * k; if (row && i < ORL_PK_STRIDE / 4) __stcg(rdst + i, w[k]); }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp(qmask);
  } else {
    reset_robot<OPT>(S, E, Mo, E.cfg.gid0 + robot_c, leg, qmask);
  }
  reset_rings(S, E, Mo, E.cfg.gid0 + robot_c, leg, qmask, valid);
  if (valid && leg == 0) queue_for_preparation(E, robot_c);
#ifdef ORL_DEBUG_PREP
  if (valid && leg == 0) atomicAdd(E.stats + (ready ? 4 : 5), 1.0);
#endif
  return ready;
}
// Thread blocks [0, prep_blocks) of a step launch: one quad per queued robot (or, in scan mode, per robot of the batch
// whose record is missing) writes the record of that robot's next episode.  Scratch records and a copy of the model
// live in shared memory like in the worker blocks.
template <bool OPT>
ORL_DI void prepare_episodes(const DevEnv& E, unsigned char* smem) {
  const int lane = threadIdx.x & 31, leg = lane & 3, rw = lane >> 2, warp = threadIdx.x >> 5;
  {
    const float4* src = reinterpret_cast<const float4*>(E.model);
    float4* dst = reinterpret_cast<float4*>(smem);
    for (int i = threadIdx.x; i < (int)(ORL_SMEM_MODEL_BYTES / 16); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
  }
  const DevModel& Mo = *reinterpret_cast<const DevModel*>(smem);
  float* S = reinterpret_cast<float*>(smem + ORL_SMEM_MODEL_BYTES) + (size_t)(warp * ORL_ROBOTS_PER_WARP + rw) * ORL_STRIDE;
  const unsigned qmask = 0xFu << (lane & ~3);
  const int n = E.cfg.n;
  if (!E.next_ep) return;
  const int quads = E.prep_blocks * ORL_ROBOTS_PER_BLOCK;
  const int count = E.prep_scan ? n : min(*reinterpret_cast<volatile const int*>(E.todo_count + E.prep_parity), 2 * n);
  const int* list = E.todo + (size_t)E.prep_parity * 2 * n;
  for (int base = ((int)blockIdx.x * ORL_WARPS_PER_BLOCK + warp) * ORL_ROBOTS_PER_WARP; base < count; base += quads) {
    const int item = base + rw;
    bool need = false;
    int episode = 0, robot = 0;
    if (item < count) {
      robot = E.prep_scan ? item : __ldcg(list + item);
      episode = *reinterpret_cast<volatile const int*>(E.state + (size_t)robot * ORL_STRIDE + S_EPISODE);
      need = *reinterpret_cast<volatile const int*>(E.prepared + robot) != episode + 1;
    }
    if (!__any_sync(0xffffffffu, need)) continue;
    if (need) {
      const float4* src = reinterpret_cast<const float4*>(E.state + (size_t)robot * ORL_STRIDE);
      float4* rec = reinterpret_cast<float4*>(S);
      {
        constexpr int kPerLane = (ORL_STRIDE / 4 + 3) / 4;
        float4 v[kPerLane];
#pragma unroll
        for (int k = 0; k < kPerLane; ++k) { const int i = leg + 4 * k; if (i < ORL_STRIDE / 4) v[k] = __ldcg(src + i); }
#pragma unroll
        for (int k = 0; k < kPerLane; ++k) { const int i = leg + 4 * k; if (i < ORL_STRIDE / 4) rec[i] = v[k]; }      // base: fields a reset leaves alone
      }
      __syncwarp(qmask);
      if (leg == 0) as_int(S[S_EPISODE]) = episode;     // the record may be mid write-back by its own warp: this word decides
      __syncwarp(qmask);
      reset_robot<OPT>(S, E, Mo, E.cfg.gid0 + robot, leg, qmask);
      __syncwarp(qmask);
      float4* dst = reinterpret_cast<float4*>(E.next_ep + (size_t)robot * ORL_NEXT_STRIDE);
#pragma unroll 1
      for (int i = leg; i < ORL_STRIDE / 4; i += 4) __stcg(dst + i, rec[i]);
      // target frames of the new episode's first observation (clock of the fresh record)
      produce_targets(E.mt, load_ref_frame(S, E.cfg), motion_time(S, E.cfg), E.cfg, E.next_ep + (size_t)robot * ORL_NEXT_STRIDE + ORL_STRIDE, leg);
      // and the packet row of its first step (substep counter 0 -> action_repeat)
      if (!OPT) produce_row(E.mt, E.cfg, as_int(S[S_CLIP]), as_int(S[S_WARM]) != 0, ld_f64(S + S_TOFF), E.cfg.action_repeat,
                            E.next_ep + (size_t)robot * ORL_NEXT_STRIDE + ORL_STRIDE + ORL_PK_TAR_FLOATS, leg);
      __threadfence();                                  // the other three lanes' stores before the release below
      __syncwarp(qmask);
      if (leg == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(E.prepared + robot), "r"(episode + 1) : "memory");
    }
    __syncwarp();
  }
}

// Thread blocks [0, prep_blocks) of a step launch, before they prepare episodes: one quad per robot writes the packet row the
// robot's worker will read in the NEXT launch (unless the robot is reset in this one: its worker then reads the row
// finish_episode left in part 2).  A row is a pure function of four fields of the record (stored with it as its key); a
// worker only takes a row whose key matches.
ORL_DI void reference_ahead(const DevEnv& E) {
  if (!E.ahead) return;
  const DevCfg& C = E.cfg;
  const int lane = threadIdx.x & 31, leg = lane & 3, rw = lane >> 2, warp = threadIdx.x >> 5;
  const int chunks = (C.n + ORL_ROBOTS_PER_WARP - 1) / ORL_ROBOTS_PER_WARP;
  float* half = E.ahead + (size_t)(E.ahead_parity ^ 1) * E.ahead_rows * ORL_PK_STRIDE;
  for (int chunk = (int)blockIdx.x * ORL_WARPS_PER_BLOCK + warp; chunk < chunks; chunk += E.prep_blocks * ORL_WARPS_PER_BLOCK) {
    const int robot = chunk * ORL_ROBOTS_PER_WARP + rw;
    const float* R = E.state + (size_t)min(robot, C.n - 1) * ORL_STRIDE;
    const int clip = __float_as_int(__ldcg(R + S_CLIP)), warm = __float_as_int(__ldcg(R + S_WARM));
    const int substep2 = __float_as_int(__ldcg(R + S_SUBSTEP)) + 2 * C.action_repeat;
    const double toff = __ldcg(reinterpret_cast<const double*>(R + S_TOFF));
    // (the buffer has whole chunks: rows past the batch are scratch; a clip index read mid write-back is clamped, its row's key will not match)
    produce_row(E.mt, C, min(max(clip, 0), E.mt.n_clips - 1), warm != 0, toff, substep2, half + (size_t)robot * ORL_PK_STRIDE, leg);
    __syncwarp();
  }
}

// Per-lane state advanced by the physics substeps, and the per-step values of the action path.  Everything that
// is constant over the step and cheap to re-read (joint/motor constants in the shared-memory model, motor
// strengths in the staged record) is NOT kept here: the substep loop is register bound (255) and every
// loop-invariant register costs a spill in the hot path (profiles/r01_v3_step_kernel.md).
struct LaneDyn {
  float3 pos, lin, ang;
  Q4 quat;
  float q[3], qd[3];
  float r1[3], r2[3];         // true motor angles one / two substeps ago (control latency ring)
};
struct LaneCtl {
  float target[3], prevf[3];
  float strength[3];
  float4 c1[3], m0[3], m1[3];   // joint limits / motor constants of this lane's joints (registers are not scarce since the ABA rewrite)
  float fric[3];             // Coulomb friction torque of this lane's joints (randomizer; 0: none)
  RandLane rl;
  int lat_n; float lat_a;    // control latency = (lat_n + lat_a) substeps (LAT kernels)
  int ar;                    // substeps per control step
  int el0;                   // index of the newest observation entry at the first substep of this step (LAT kernels)
  float* ring_q; float* ring_b;   // this robot's rings (LAT kernels)
  float* qs;                      // this lane's joint angles in the staged record (S_Q + 3 leg); velocities at + (S_QD - S_Q)
  float mac, inv_ar;
  int capture;
  bool first_step;
};

// One 1 ms physics substep of this lane's leg (+ the redundantly integrated base).
ORL_DI void cp_async4(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_addr(smem_dst)), "l"(gmem_src) : "memory");
}
// Ring slots of the two entries the control observation of entry index `el` blends (minitaur.py:336-357):
// history[n] and history[n + 1] counted back from the newest entry el; both the oldest entry (0) while the
// history is still shorter than the latency.
ORL_DI void lat_slots(int el, int n, int& a, int& b) {
  const bool oldest = n >= el;              // n + 1 >= len(history), len = el + 1
  a = oldest ? 0 : (el - n) % ORL_LAT_RING;
  b = oldest ? 0 : (el - n - 1) % ORL_LAT_RING;
}
ORL_DI void obs_base(const LaneDyn& D, const DevModel& Mo, Q4& oq, float3& ow) {
  oq = qmul(D.quat, qconj(ldq4(Mo.iquat)));                          // minitaur.py:325-331
  ow = qrot(qconj(oq), D.ang);                                       // :652-672
}

template <bool LAT>
ORL_DI void substep(const DevLeg& L, const DevModel& Mo, const PhysParams& P, const LaneCtl& Ctl, LaneDyn& D, float* S, float4* X, int s) {
  const M3 Rb = qmat(D.quat);
  const int el = Ctl.el0 + s;                                         // LAT: the current true state is observation entry el
  if (LAT) {
    asm volatile("cp.async.wait_all;" ::: "memory");                  // prefetched history entries of this substep
    // base part of the observation: only the entries an IMU reading will blend are stored -- history[n], history[n+1]
    // counted back from the last entry of this control step (el0 + AR) or of the next one (el0 + 2 AR)
    const int back = Ctl.el0 + Ctl.ar - el - Ctl.lat_n;
    if ((threadIdx.x & 3) == 0 && (back == 0 || back == 1 || back == -Ctl.ar || back == 1 - Ctl.ar)) {
      Q4 oq; float3 ow;
      obs_base(D, Mo, oq, ow);
      float4* rb = reinterpret_cast<float4*>(Ctl.ring_b + (el % ORL_LAT_RING) * 8);
      rb[0] = make_float4(oq.x, oq.y, oq.z, oq.w); rb[1] = make_float4(ow.x, ow.y, ow.z, 0.f);
    }
  }
  const float lerp = (float)(s + 1) * Ctl.inv_ar;
  float tau[3], qm_now[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 c1 = Ctl.c1[k], m0 = Ctl.m0[k], m1 = Ctl.m1[k];   // (comy comz lower upper) (damping dir offset kp) (kd init motor brad)
    const float strength = Ctl.strength[k];
    const float qm = (D.q[k] - m0.z) * m0.y;                               // :549-553
    float ctrl = D.r2[k];
    if (LAT) {
      const int m = __float_as_int(m1.z);
      Ctl.ring_q[(el % ORL_LAT_RING) * 12 + m] = qm;                       // receive_obs: entry el (minitaur.py:313-334)
      const float hA = Ctl.lat_n == 0 ? qm : S[S_RING1 + m], hB = S[S_RING2 + m];
      ctrl = (1.0f - Ctl.lat_a) * hA + Ctl.lat_a * hB;                      // _get_delay_obs
    }
    const float qobs = wrap_pi_fast(ctrl);                                 // get_motor_angles(): delayed + wrapped
    const float prev = Ctl.first_step ? qobs : Ctl.prevf[k];               // minitaur.py:450-453
    float cmd = fmaf(lerp, Ctl.target[k] - prev, prev);                    // :455-456
    cmd = fminf(fmaxf(cmd, qobs - Ctl.mac), qobs + Ctl.mac);               // :706-723
    const float qdm = D.qd[k] * m0.y;                                      // :575-578
    const float tm = strength * (-1.0f * (m0.w * (qm - cmd)) - m1.x * qdm);   // minitaur_motor.py:163-166
    const float e = fmaxf(0.f, c1.z - D.q[k]) - fmaxf(0.f, D.q[k] - c1.w);
    tau[k] = fmaf(tm, m0.y, fmaf(P.klim, e, -(fmaf(P.dlim, fminf(1.0f, fabsf(e) * P.inv_e0), m0.x)) * D.qd[k]));
    tau[k] -= fminf(fmaxf(P.kfric * D.qd[k], -Ctl.fric[k]), Ctl.fric[k]);     // set_joint_friction (minitaur.py:1063-1070), regularised
    qm_now[k] = qm;
  }
  if (LAT) {      // request the two history entries of the NEXT control observation (entry el + 1) into the record's ring slots
    int a, b;
    lat_slots(el + 1, Ctl.lat_n, a, b);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int m = __float_as_int(L.j[k].m1.z);
      if (Ctl.lat_n != 0) cp_async4(S + S_RING1 + m, Ctl.ring_q + a * 12 + m);
      cp_async4(S + S_RING2 + m, Ctl.ring_q + b * 12 + m);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  float3 alpha, alin; float qdd[3];
  leg_dynamics<true, false, true>(L, Mo, P, Rb, D.pos.z, D.ang, D.lin, D.q, D.qd, tau, X, Ctl.rl, alpha, alin, qdd, nullptr, Ctl.qs, Ctl.qs + (S_QD - S_Q));
  // semi-implicit Euler: velocities first, then positions with the new velocities
  const float3 lin_dot = cross_add(D.ang, D.lin, alin);
  D.ang = clamp3(fma3(P.dt, alpha, D.ang), P.vmax);
  D.lin = clamp3(fma3(P.dt, lin_dot, D.lin), P.vmax);
  D.pos = fma3(P.dt, D.lin, D.pos);
  {
    // quaternion exponential of w*dt: sin(h)/|w| and cos(h), h = |w| dt / 2 (Taylor series while h < 0.05)
    const float w2 = dot(D.ang, D.ang);
    const float h2 = 0.25f * P.dt * P.dt * w2;
    float kk = 0.5f * P.dt * fmaf(h2, fmaf(h2, 1.0f / 120.0f, -1.0f / 6.0f), 1.0f);
    float ch = fmaf(h2, fmaf(h2, 1.0f / 24.0f, -0.5f), 1.0f);
    if (h2 >= 2.5e-3f) {
      const float iw = rsqrtf(w2);
      float sh;
      __sincosf(0.5f * w2 * iw * P.dt, &sh, &ch);
      kk = sh * iw;
    }
    D.quat = qnormalize(qmul(q4(D.ang.x * kk, D.ang.y * kk, D.ang.z * kk, ch), D.quat));
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    D.qd[k] = fminf(fmaxf(fmaf(P.dt, qdd[k], D.qd[k]), -P.vmax), P.vmax);
    D.q[k] = fmaf(P.dt, D.qd[k], D.q[k]);
    Ctl.qs[k] = D.q[k]; Ctl.qs[(S_QD - S_Q) + k] = D.qd[k];        // mirror for the rolled per-joint loop of the next substep
    D.r2[k] = D.r1[k]; D.r1[k] = qm_now[k];
  }
  if (!LAT && s == Ctl.capture) {   // the control observation at the end of the step is 2 substeps old (minitaur.py:336-357)
    const Q4 oq = qmul(D.quat, qconj(ldq4(Mo.iquat)));                 // :325-331
    const float3 ow = qrot(qconj(oq), D.ang);                          // :652-672
    stq(S + S_OBSQ, oq); st3(S + S_OBSW, ow);                          // identical in the four lanes of the quad
  }
}

// =================================================================================================
// The fused control step (WrapperEnv.step): SURVEY.md A.3 steps 1-7.
// =================================================================================================
#ifndef ORL_STEP_MIN_BLOCKS
#define ORL_STEP_MIN_BLOCKS (8 / ORL_WARPS_PER_BLOCK)
#endif
template <bool LAT, bool OPT>
ORL_DI void step_worker(const DevEnv& E, unsigned char* smem, const float* __restrict__ actions, float* __restrict__ obs_out,
                        float* __restrict__ reward_out, uint8_t* __restrict__ done_out) {
#ifdef ORL_PROFILE_RARE
  const long long tk0 = clock64();
#endif
  // the list this launch FILLS (resets happen a whole step from now; the prepare blocks read the other one)
  if (E.todo && (int)blockIdx.x == E.prep_blocks && threadIdx.x == 0) E.todo_count[E.prep_parity ^ 1] = 0;
  const DevModel* Mo; float* st; float* ob;
  if (!stage_in(E, smem, Mo, st, ob)) return;
#ifdef ORL_PROFILE_RARE
  const long long tk1 = clock64();
#endif
  const DevCfg& C = E.cfg;
  const int lane = threadIdx.x & 31, leg = lane & 3, rw = lane >> 2;
  const int robot = warp_row0(E) + rw;
  const bool valid = robot < C.n;
  const int robot_c = min(robot, C.n - 1);
  float* S = st + rw * ORL_STRIDE;
  float* O = ob + rw * ORL_OBS_SIZE;
  float4* X = reinterpret_cast<float4*>(ob) + lane;      // dynamics scratch; aliases the observation staging (dead until build_obs)
  const DevLeg& L = Mo->leg[leg];
  const PhysParams P = C.phys;

  // ---- registers -----------------------------------------------------------------------------
  float3 pos = ld3(S + S_POS), lin = ld3(S + S_LIN), ang = ld3(S + S_ANG);
  Q4 quat = ldq(S + S_QUAT);
  float q[3], qd[3], target[3], act[3], prevf[3], r1[3], r2[3];
  int mi[3];
  const int substep0 = as_int(S[S_SUBSTEP]);
  const bool first_step = as_int(S[S_STEP]) == 0;
  // 1. set_act: a += INIT; Butterworth (minitaur.py:280-285, 1169-1178; action_filter.py:111-127)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 m1 = L.j[k].m1;
    const int m = __float_as_int(m1.z);
    mi[k] = m;
    q[k] = S[S_Q + 3 * leg + k]; qd[k] = S[S_QD + 3 * leg + k];
    r1[k] = S[S_RING1 + m]; r2[k] = S[S_RING2 + m];
    prevf[k] = S[S_PREV + m];
    act[k] = __ldg(actions + (size_t)robot_c * 12 + m) + m1.y;
    float x0 = S[S_FX + m], x1 = S[S_FX + 12 + m], y0 = S[S_FY + m], y1 = S[S_FY + 12 + m];
    if (substep0 == 0) { const float d = wrap_pi_fast(r2[k]); x0 = x1 = y0 = y1 = d; }
    const float y = act[k] * C.fb0 + (x0 * C.fb1 + x1 * C.fb2) - (y0 * C.fa1 + y1 * C.fa2);
    S[S_FX + m] = act[k]; S[S_FX + 12 + m] = x0; S[S_FY + m] = y; S[S_FY + 12 + m] = y0;
    target[k] = y;
  }
  // 1b. The part of the observation that does not depend on this step's dynamics leaves NOW: the sensor histories are
  // shifted (sensor_wrappers.py:131-135, newest first) and the two older IMU / motor-angle readings plus the three last
  // actions (68 of the 160 floats) go out with three small bulk copies per robot.  When the observation buffer is mapped
  // host memory (orl_step_host) that is 1.1 of the 2.6 MB per step crossing PCIe behind the substep loop instead of after
  // the kernel's last warp.  A robot that is reset at the end of the step rewrites its whole row.
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    S[S_ACT + 24 + mi[k]] = S[S_ACT + 12 + mi[k]]; S[S_ACT + 12 + mi[k]] = S[S_ACT + mi[k]]; S[S_ACT + mi[k]] = act[k];
    S[S_ANGH + 24 + mi[k]] = S[S_ANGH + 12 + mi[k]]; S[S_ANGH + 12 + mi[k]] = S[S_ANGH + mi[k]];
  }
  if (leg == 0) {
#pragma unroll
    for (int i = 11; i >= 4; --i) S[S_IMU + i] = S[S_IMU + i - 4];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (E.obs_split && leg == 0 && valid) {
    float* row = obs_out + (size_t)robot * ORL_OBS_SIZE;
    bulk_s2g(row + 4, smem_addr(S + S_IMU + 4), 8 * 4);
    bulk_s2g(row + 12, smem_addr(S + S_ACT), 36 * 4);
    bulk_s2g(row + 48 + 12, smem_addr(S + S_ANGH + 12), 24 * 4);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  // 2. physics substeps (quadruped_gym_env.py:220-225)
  // The loop body is kept below the SM's 32 KB instruction cache (one warp per scheduler cannot hide fetch
  // misses): single copy of the contact code, rolled exact tests, SFU trigonometry (profiles/r01_v3_step_kernel.md).
  LaneDyn D;
  D.pos = pos; D.quat = quat; D.lin = lin; D.ang = ang;
#pragma unroll
  for (int k = 0; k < 3; ++k) { D.q[k] = q[k]; D.qd[k] = qd[k]; D.r1[k] = r1[k]; D.r2[k] = r2[k]; }
  LaneCtl Ctl;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    Ctl.target[k] = target[k]; Ctl.prevf[k] = prevf[k]; Ctl.strength[k] = S[S_STRENGTH + __float_as_int(L.j[k].m1.z)];
    Ctl.fric[k] = __float_as_int(L.j[k].i3.w) >= 0 ? S[S_RAND + 6 + leg] : 0.f;
    Ctl.c1[k] = L.j[k].c1; Ctl.m0[k] = L.j[k].m0; Ctl.m1[k] = L.j[k].m1;
  }
  Ctl.rl.base_mass_ratio = S[S_RAND + 0]; Ctl.rl.base_inertia_ratio = S[S_RAND + 2]; Ctl.rl.mu_foot = S[S_RAND + 5];
  rand_lane_toe(L, *Mo, Ctl.rl);
  {
    const float ratios[4] = {S[S_RAND + 0], S[S_RAND + 1], S[S_RAND + 2], S[S_RAND + 3]};
    fill_scaled_inertia(L, ratios, X + ORL_XI_BASE * ORL_X_STRIDE);
  }
  Ctl.first_step = first_step; Ctl.mac = Mo->b1.w;
  Ctl.qs = S + S_Q + 3 * leg;
  const int AR = C.action_repeat;
  Ctl.inv_ar = 1.0f / (float)AR; Ctl.capture = AR - 3;
  Ctl.lat_n = 0; Ctl.lat_a = 0.f; Ctl.el0 = substep0 + 1; Ctl.ring_q = nullptr; Ctl.ring_b = nullptr; Ctl.ar = AR;
  if (LAT) {
    // entries: 0 = default pose (Minitaur.reset), 1 = synced pose (ImitationTask.reset), 2 + j = after substep j
    const double lat = (double)S[S_RAND + 4] / C.sim_dt;
    Ctl.lat_n = (int)lat;                                              // int(latency / time_step)
    Ctl.lat_a = (float)(lat - (double)Ctl.lat_n);                      // remaining_latency / time_step
    const int ring_row = valid ? robot : C.n;                          // padding quads: scratch row (rings have n + 1 rows)
    Ctl.ring_q = E.ring_q + (size_t)ring_row * (ORL_LAT_RING * 12);
    Ctl.ring_b = E.ring_b + (size_t)ring_row * (ORL_LAT_RING * 8);
    int a, b;
    lat_slots(Ctl.el0, Ctl.lat_n, a, b);
    __syncwarp();                                                      // r1 / r2 were read from these record slots above
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (Ctl.lat_n != 0) cp_async4(S + S_RING1 + mi[k], Ctl.ring_q + a * 12 + mi[k]);
      cp_async4(S + S_RING2 + mi[k], Ctl.ring_q + b * 12 + mi[k]);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  __syncwarp();
#ifdef ORL_PROFILE_RARE
  const long long tl0 = clock64();
  if ((threadIdx.x & 31) == 0) { atomicAdd(&orl_prof[12], (unsigned long long)(tk1 - tk0)); atomicAdd(&orl_prof[13], (unsigned long long)(tl0 - tk1)); }
#endif
  for (int s = 0; s < AR; ++s) {
#if ORL_WARPS_PER_BLOCK > 1 && defined(ORL_LOCKSTEP)
    __syncthreads();                       // experiment: keep the block's warps in the same code region (no gain measured)
#endif
    substep<LAT>(L, *Mo, P, Ctl, D, S, X, s);
  }
#ifdef ORL_PROFILE_RARE
  const long long tend_loop = clock64();
  if (lane == 0) {
    atomicAdd(&orl_prof[3], (unsigned long long)(clock64() - tl0)); atomicAdd(&orl_prof[4], 1ull);
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned long long rc = orl_prof_warp[0][gw], rt = orl_prof_warp[1][gw];
    orl_prof_warp[0][gw] = 0; orl_prof_warp[1][gw] = 0;
    int b = (int)(rc / 8192); if (b > 7) b = 7;
    (void)b;
    int t = (int)(rt / 16); if (t > 3) t = 3;
    atomicAdd(&orl_prof[16 + t], 1ull);                   // histogram of rare trips per warp-step, buckets of 16
    atomicMax(&orl_prof[5], rc); atomicMax(&orl_prof[6], rt);
  }
#endif
  if (E.obs_split && leg == 0 && valid) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the early observation pieces have left shared memory
  __syncwarp();
  pos = D.pos; quat = D.quat; lin = D.lin; ang = D.ang;
#pragma unroll
  for (int k = 0; k < 3; ++k) { q[k] = D.q[k]; qd[k] = D.qd[k]; r1[k] = D.r1[k]; r2[k] = D.r2[k]; }
  if (LAT) {
    // receive_obs after the last substep: entry el of the final state, then the control observation the sensors
    // read (and the next step starts from) out of the ring with this robot's latency
    const int el = Ctl.el0 + AR;
    asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float qm = (q[k] - L.j[k].m0.z) * L.j[k].m0.y;
      Ctl.ring_q[(el % ORL_LAT_RING) * 12 + mi[k]] = qm;
      const float hA = Ctl.lat_n == 0 ? qm : S[S_RING1 + mi[k]], hB = S[S_RING2 + mi[k]];
      r2[k] = (1.0f - Ctl.lat_a) * hA + Ctl.lat_a * hB;
    }
    if (leg == 0) {
      Q4 oq; float3 ow;
      obs_base(D, *Mo, oq, ow);
      float4* rb = reinterpret_cast<float4*>(Ctl.ring_b + (el % ORL_LAT_RING) * 8);
      rb[0] = make_float4(oq.x, oq.y, oq.z, oq.w); rb[1] = make_float4(ow.x, ow.y, ow.z, 0.f);
      int a, b;
      lat_slots(el, Ctl.lat_n, a, b);
      const float4 a0 = Ctl.lat_n == 0 ? rb[0] : reinterpret_cast<const float4*>(Ctl.ring_b + a * 8)[0];
      const float4 a1 = Ctl.lat_n == 0 ? rb[1] : reinterpret_cast<const float4*>(Ctl.ring_b + a * 8)[1];
      const float4 b0 = reinterpret_cast<const float4*>(Ctl.ring_b + b * 8)[0], b1 = reinterpret_cast<const float4*>(Ctl.ring_b + b * 8)[1];
      const float wa = 1.0f - Ctl.lat_a, wb = Ctl.lat_a;              // linear blend of the raw observation vectors
      stq(S + S_OBSQ, q4(wa * a0.x + wb * b0.x, wa * a0.y + wb * b0.y, wa * a0.z + wb * b0.z, wa * a0.w + wb * b0.w));
      st3(S + S_OBSW, v3(wa * a1.x + wb * b1.x, wa * a1.y + wb * b1.y, wa * a1.z + wb * b1.z));
    }
    __syncwarp();
  }
  const Q4 obs_q = ldq(S + S_OBSQ);
  const float3 obs_w = ld3(S + S_OBSW);
  // ---- reference packets of this warp's robots: requested now, read in sections 5 and 7 --------------------
  float* const pk = pk_rows(ob) + rw * ORL_PK_STRIDE;
  bool ahead = !OPT && E.ahead != nullptr;
  if (ahead) {        // written by the previous launch: no flag to wait for, the copies overlap the write-back and the reward
    const int nrow = min(ORL_ROBOTS_PER_WARP, C.n - warp_row0(E));
    const float4* src = reinterpret_cast<const float4*>(E.ahead) + (size_t)warp_row0(E) * (ORL_PK_STRIDE / 4);
    float4* dst = reinterpret_cast<float4*>(pk_rows(ob));
#pragma unroll 1
    for (int i = lane; i < ORL_ROBOTS_PER_WARP * ORL_PK_STRIDE / 4; i += 32) {
      const int row = i / (ORL_PK_STRIDE / 4), col = i - row * (ORL_PK_STRIDE / 4);
      const int part = as_int(st[row * ORL_STRIDE + S_SUBSTEP]) == 0 ? 2 : E.ahead_parity;      // first step of an episode: the row its reset left
      cp_async16(dst + i, src + ((size_t)part * E.ahead_rows + min(row, nrow - 1)) * (ORL_PK_STRIDE / 4) + col);   // idle quads: the last robot's row, like their record
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // ---- write the dynamic state back to the record ---------------------------------------------------

#pragma unroll
  for (int k = 0; k < 3; ++k) {
    S[S_Q + 3 * leg + k] = q[k]; S[S_QD + 3 * leg + k] = qd[k];
    S[S_RING1 + mi[k]] = r1[k]; S[S_RING2 + mi[k]] = r2[k];
    S[S_PREV + mi[k]] = target[k];                                     // minitaur.py:291-292
    // 3. sensors.on_step: the newest motor-angle reading (the older entries were shifted in 1b)
    S[S_ANGH + mi[k]] = wrap_pi_fast(r2[k]);
  }
  const int substep1 = substep0 + AR;
  const int envstep1 = as_int(S[S_ENVSTEP]) + 1;
  if (leg == 0) {
    st3(S + S_POS, pos); stq(S + S_QUAT, quat); st3(S + S_LIN, lin); st3(S + S_ANG, ang);
    stq(S + S_OBSQ, obs_q);
    as_int(S[S_SUBSTEP]) = substep1;
    as_int(S[S_STEP]) = as_int(S[S_STEP]) + 1;
    as_int(S[S_ENVSTEP]) = envstep1;
    float roll, pitch, yaw;
    bullet_euler(obs_q, roll, pitch, yaw);                             // IMUSensor channels R, P, dR, dP
    S[S_IMU + 0] = roll; S[S_IMU + 1] = pitch; S[S_IMU + 2] = obs_w.x; S[S_IMU + 3] = obs_w.y;
  }
  // 3b. host buffers: the newest IMU / motor-angle readings (16 floats) cross PCIe behind the rest of the once-per-step code
  // instead of after it (a robot that is reset below waits for these copies and rewrites its whole row)
  if (E.obs_split) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (leg == 0 && valid) {
      float* row = obs_out + (size_t)robot * ORL_OBS_SIZE;
      bulk_s2g(row, smem_addr(S + S_IMU), 4 * 4);
      bulk_s2g(row + 48, smem_addr(S + S_ANGH), 12 * 4);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
#ifdef ORL_PROFILE_RARE
  long long tsec = clock64(); if (lane == 0) atomicAdd(&orl_prof[8], (unsigned long long)(tsec - tend_loop));
#endif
  // 4. reward against the reference state of the END of the previous step (imitation_task.py:341-516)
  const M3 Rb = qmat(quat);
  LegFK Ks, Kr;
  leg_fk<true>(L, Rb, q, Ks);
  const float3 rp = ld3(S + S_REFPOSE);
  const Q4 rq = ldq(S + S_REFPOSE + 3);
  float rj[3];
  float pose_err = 0.f, vel_err = 0.f, ee_err = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    rj[k] = S[S_REFPOSE + 7 + 3 * leg + k];
    const float dq = rj[k] - q[k], dv = S[S_REFVEL + 6 + 3 * leg + k] - qd[k];
    pose_err += dq * dq; vel_err += dv * dv;
  }
  leg_fk<true>(L, qmat(rq), rj, Kr);
  {
    const Q4 inv_def = qconj(ldq4(Mo->iquat));                      // default root rotation, :1071-1075, 1245-1252
    float ss, cs, sr, cr;
    qheading_neg_sincos(qmul(quat, inv_def), ss, cs);
    qheading_neg_sincos(qmul(rq, inv_def), sr, cr);
#pragma unroll 1              // one copy of the body: once-per-step code pays per fetched instruction
    for (int e = 0; e < L.meta.x; ++e) {
      const float4 ee = L.ee[e];
      const int kb = __float_as_int(ee.w);
      const float3 loc = f4xyz(ee);
      const float3 es = pickp(Ks, kb) + mul(pickR(Ks, kb), loc);       // rel. sim root
      const float3 er = pickp(Kr, kb) + mul(pickR(Kr, kb), loc);       // rel. ref root
      const float3 d = qrot_z(sr, cr, er) - qrot_z(ss, cs, es);
      const float dh = (rp.z + er.z) - (pos.z + es.z);
      ee_err += d.x * d.x + d.y * d.y + C.s_eeh * dh * dh;
    }
  }
  pose_err = quad_sum(pose_err); vel_err = quad_sum(vel_err); ee_err = quad_sum(ee_err);
  float reward;
  {
    const float3 dp = rp - pos;
    const float a = qangle_normalized(qmul(rq, qconj(quat)));
    const float root_err = dot(dp, dp) + 0.5f * a * a;
    const float3 dv = ld3(S + S_REFVEL) - lin, dw = ld3(S + S_REFVEL + 3) - ang;
    const float rootvel_err = dot(dv, dv) + 0.1f * dot(dw, dw);
    reward = C.w_pose * __expf(-C.s_pose * pose_err) + C.w_vel * __expf(-C.s_vel * vel_err) + C.w_ee * __expf(-C.s_ee * ee_err) +
             C.w_root * __expf(-C.s_root * root_err) + C.w_rootvel * __expf(-C.s_rootvel * rootvel_err);   // ex2.approx: rel. error < 2e-6
  }
  // Guard (no reference counterpart): a non-finite state ends the episode with reward 0 instead of leaking NaN into
  // observations and statistics.  Velocity clamping keeps the integrator finite; this only catches corrupted input.
  const float chk = pos.x + pos.y + pos.z + quat.x + quat.y + quat.z + quat.w + q[0] + q[1] + q[2] + qd[0] + qd[1] + qd[2] +
                    lin.x + lin.y + lin.z + ang.x + ang.y + ang.z;
  const int broken = quad_or(!(fabsf(chk) < 3.0e38f) ? 1 : 0);
  if (OPT) reward *= C.reward_weight;                                  // imitation_task.py:356
  if (broken) reward = 0.f;
  __syncwarp();
#ifdef ORL_PROFILE_RARE
  { const long long tn = clock64(); if (lane == 0) atomicAdd(&orl_prof[9], (unsigned long long)(tn - tsec)); tsec = tn; }
#endif
  // 5. task.update (imitation_task.py:201-207, 734-761, 1020-1055)
  RefFrame F = load_ref_frame(S, C);
  if (ahead) {
    // the rows have landed; a row is only good for the record it was computed from (quad-uniform)
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    ahead = as_int(pk[PK_KEY]) == as_int(S[S_SUBSTEP]) && as_int(pk[PK_KEY + 1]) == 2 * F.clip + (F.warmup ? 1 : 0) &&
            as_int(pk[PK_KEY + 2]) == as_int(S[S_TOFF]) && as_int(pk[PK_KEY + 3]) == as_int(S[S_TOFF + 1]);
  }
#ifdef ORL_DEBUG_PREP
  if (leg == 0 && valid) atomicAdd(E.stats + (ahead ? 6 : 7), 1.0);
#endif
  double time = 0.0;
  bool change = false;
  double new_phase;
  if (!ahead) time = motion_time(S, C);
  if (OPT) {
    // clip switching (_update_ref_motion, imitation_task.py:740-746): a new clip, a new timer measured from the clock
    // value BEFORE the new time offset is drawn, the new offset; the phase compared below is the NEW clip's at the OLD clock
    change = time >= ld_f64(S + S_CHANGET) && E.mt.n_clips > 1;      // _check_change_clip :1096-1101
    const unsigned qm = 0xFu << (lane & ~3);
    if (change) {
      const int nchange = as_int(S[S_NCHANGE]);
      const int gid = C.gid0 + robot_c, ekey = as_int(S[S_EPISODE]) - 1;
      uint32_t wa[4], wb[4];
      opt_block(C, gid, ekey, ORL_CHANGE_BLOCK0 + 2u * (uint32_t)nchange, wa);
      opt_block(C, gid, ekey, ORL_CHANGE_BLOCK0 + 1u + 2u * (uint32_t)nchange, wb);
      const int nclip = (int)(((unsigned long long)wa[0] * (unsigned long long)E.mt.n_clips) >> 32);
      const double dur_v = __dadd_rn(C.clip_tmin, __dmul_rn(__dsub_rn(C.clip_tmax, C.clip_tmin), u01_53(wa[2], wa[3])));
      const double toff = __dadd_rn(0.0, __dmul_rn(__dsub_rn(E.mt.dur[nclip], 0.0), u01_53(wb[0], wb[1])));   // _sample_time_offset
      __syncwarp(qm);                                   // every lane of the quad has read the old fields
      if (leg == 0) {
        as_int(S[S_CLIP]) = nclip; as_int(S[S_NCHANGE]) = nchange + 1;
        st_f64(S + S_CHANGET, __dadd_rn(time, dur_v)); st_f64(S + S_TOFF, toff);
      }
      __syncwarp(qm);
      F.clip = nclip;
    }
    new_phase = motion_phase(E.mt, F.clip, time);
    if (change) time = motion_time(S, C);               // _sync_ref_origin / _update_ref_state re-read the clock (:1033, 765)
  }
  if (!ahead) produce_update(E.mt, F, time, pk, leg);       // no row from the leading blocks: the same arithmetic, here
  __syncwarp();
  if (!OPT) new_phase = ld_f64(pk + PK_PHASE);
  const bool motion_over = as_int(pk[PK_OVER]) != 0;
  float3 nrp = ld3(pk + PK_ROOT), nrl = ld3(pk + PK_VEL), nra = ld3(pk + PK_VEL + 3);
  Q4 nrq = ldq(pk + PK_ROOT + 3);                 // sampled once: the cycle-sync re-anchoring only moves the origin
  if (OPT && change) {                            // sync_root_rotation=change_clip (:1037-1045), heading rel. to the default rotation
    const Q4 inv_def = qconj(ldq4(Mo->iquat));
    F.origin_rot = qaxis_z(qheading(qmul(quat, inv_def)) - qheading(qmul(nrq, inv_def)));
  }
  if ((C.cycle_sync && new_phase < ld_f64(S + S_PPHASE)) || (OPT && change)) {
    float3 op = pos - qrot(F.origin_rot, nrp); op.z = 0.f;
    F.origin_pos = op;
  }
  ref_apply_origin(F, nrp, nrq);
  nrl = qrot(F.origin_rot, nrl); nra = qrot(F.origin_rot, nra);
  __syncwarp();
  if (leg == 0) {
    st3(S + S_OPOS, F.origin_pos);
    if (OPT && change) stq(S + S_OROT, F.origin_rot);
    st3(S + S_REFPOSE, nrp); stq(S + S_REFPOSE + 3, nrq); st3(S + S_REFVEL, nrl); st3(S + S_REFVEL + 3, nra);
    st_f64(S + S_PPHASE, new_phase);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    S[S_REFPOSE + 7 + 3 * leg + k] = pk[PK_J + 3 * leg + k];
    S[S_REFVEL + 6 + 3 * leg + k] = pk[PK_JV + 3 * leg + k];
  }
#ifdef ORL_PROFILE_RARE
  { const long long tn = clock64(); if (lane == 0) atomicAdd(&orl_prof[10], (unsigned long long)(tn - tsec)); tsec = tn; }
#endif
  // 6. termination against the freshly updated reference (imitation_task.py:518-572, wrapper_env.py:78-80)
  int fall = 0;
  {
    const int4 cnt = L.cnt;
    const float mmax = __int_as_float(L.meta.z);          // largest margin of a non-foot point
    int first = 0;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {          // rolled on purpose: once-per-step code is instruction-fetch bound
      const float3 rz = b == 0 ? Ks.R[0].r2 : (b == 1 ? Ks.R[1].r2 : (b == 2 ? Ks.R[2].r2 : Rb.r2));
      const float oz = pos.z + (b == 0 ? Ks.p[0].z : (b == 1 ? Ks.p[1].z : (b == 2 ? Ks.p[2].z : 0.f)));
      const float brad = b == 3 ? __int_as_float(L.meta.y) : L.j[b == 3 ? 0 : b].m1.w;
      const int num = b == 0 ? cnt.x : (b == 1 ? cnt.y : (b == 2 ? cnt.z : cnt.w));
      if (oz < brad + mmax) {
#pragma unroll 1
        for (int i = first; i < first + num; ++i) {
          const float4 a = L.cp[i].a, bb = L.cp[i].b;
          const float z = oz + dot(rz, f4xyz(a)) - a.w - bb.y;
          fall |= (__float_as_int(bb.z) == 0) && (z < 0.f);
        }
      }
      first += num;
    }
  }
  fall = quad_or(fall);      // Minitaur.step_counter > 0 always holds after a step (imitation_task.py:539)
  int done;
  {
    const float3 dp = nrp - pos;
    const float a = qangle_normalized(qmul(nrq, qconj(quat)));
    done = motion_over || fall || (dot(dp, dp) > C.dist_thr2) || (fabsf(a) > C.rot_thr);
    done |= (envstep1 >= C.max_episode_steps) || broken;
  }
  __syncwarp();
#ifdef ORL_PROFILE_RARE
  { const long long tn = clock64(); if (lane == 0) atomicAdd(&orl_prof[11], (unsigned long long)(tn - tsec)); tsec = tn; }
#endif
  // rollout statistics (SURVEY.md section 5 "metrics"): [sum reward, sum episode length, episodes, robot steps]
  float ep_ret = S[S_EPRET] + reward;
  const bool finish = done && C.auto_reset;
  {
    const bool mine = valid && leg == 0;
    float v0 = mine ? reward : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v0 += __shfl_xor_sync(0xffffffffu, v0, o);
    // the three counts are integers: one warp-reduce instruction each instead of five shuffle rounds
    const int v1 = __reduce_add_sync(0xffffffffu, (mine && finish) ? envstep1 : 0);
    const int v2 = __reduce_add_sync(0xffffffffu, (mine && finish) ? 1 : 0), v3_ = __reduce_add_sync(0xffffffffu, mine ? 1 : 0);
    if (lane == 0) {
      atomicAdd(E.stats + 0, (double)v0); if (v1) atomicAdd(E.stats + 1, (double)v1);
      if (v2) atomicAdd(E.stats + 2, (double)v2); atomicAdd(E.stats + 3, (double)v3_);
    }
  }
  __syncwarp();
  if (leg == 0) S[S_EPRET] = ep_ret;
  if (valid && leg == 0) { reward_out[robot] = reward; done_out[robot] = (uint8_t)done; }
  __syncwarp();
#ifdef ORL_PROFILE_RARE
  const long long tr0 = clock64();
  const bool any_finish = __any_sync(0xffffffffu, finish);
#endif
  bool have_tar = ahead && !(OPT && change);       // (a clip change moves the clock the target frames are read at)
  if (finish) {
    if (E.obs_split) {           // the early pieces of this row have landed before the record changes and the whole row follows
      if (leg == 0 && valid) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __syncwarp(0xFu << (lane & ~3));
    }
    have_tar = finish_episode<OPT>(S, E, *Mo, robot_c, leg, 0xFu << (lane & ~3), valid, pk + PK_TAR);
  }
#ifdef ORL_PROFILE_RARE
  __syncwarp();
  if (lane == 0 && any_finish) { atomicAdd(&orl_prof[7], (unsigned long long)(clock64() - tr0)); atomicAdd(&orl_prof[0], 0ull); atomicAdd(&orl_prof[23], 1ull); }
  const long long to0 = clock64();
#endif
  __syncwarp();
  // 7. observation (histories + target frames)
  build_obs<OPT>(S, O, E, leg, C.gid0 + robot_c, pk + PK_TAR, have_tar);
#ifdef ORL_PROFILE_RARE
  __syncwarp();
  if (lane == 0) atomicAdd(&orl_prof[22], (unsigned long long)(clock64() - to0));
  const long long ts0 = clock64();
#endif
  // late part of the observation: newest IMU / motor-angle readings and the target frames; the whole row after a reset
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (E.obs_split && leg == 0 && valid) {
    float* row = obs_out + (size_t)robot * ORL_OBS_SIZE;
    if (finish) {
      bulk_s2g(row, smem_addr(O), ORL_OBS_SIZE * 4);
    } else {
      bulk_s2g(row + 84, smem_addr(O + 84), 76 * 4);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  stage_out(E, st, ob, E.obs_split ? nullptr : obs_out);
  if (E.obs_split && leg == 0 && valid) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#ifdef ORL_PROFILE_RARE
  if (lane == 0) { atomicAdd(&orl_prof[14], (unsigned long long)(clock64() - ts0)); atomicAdd(&orl_prof[15], (unsigned long long)(ts0 - tend_loop)); }
#endif
}

// The kernel: worker blocks first in the code (the position of the substep loop inside the kernel matters: every 20 KB
// of code in front of it cost 6 us per step at N = 4096, profiles/r02_step_kernel.md "code placement"), the blocks that
// prepare episodes after it.
template <bool LAT, bool OPT>
__global__ void __launch_bounds__(ORL_BLOCK_THREADS, ORL_STEP_MIN_BLOCKS) orl_step_kernel(DevEnv E, const float* __restrict__ actions, float* __restrict__ obs_out,
                                                      float* __restrict__ reward_out, uint8_t* __restrict__ done_out) {
  extern __shared__ __align__(16) unsigned char smem[];
#ifdef ORL_PAD_KB       // experiment: ORL_PAD_KB x ~2.5 KB of never-executed instructions in front of everything (code placement probe)
  if (E.cfg.n < 0) {
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < ORL_PAD_KB * 64; ++i) { unsigned t; asm volatile("mov.u32 %0, %%clock;" : "=r"(t)); acc ^= t; }
    if (acc == 12345u) reward_out[0] = 1.f;
  }
#endif
  if ((int)blockIdx.x >= E.prep_blocks) { step_worker<LAT, OPT>(E, smem, actions, obs_out, reward_out, done_out); return; }
  if (!OPT) reference_ahead(E);        // first: ~50 k cycles at 4096 robots, and the episodes are only needed when a step ends
  prepare_episodes<OPT>(E, smem);
}

// =================================================================================================
template <bool OPT>
__global__ void __launch_bounds__(ORL_BLOCK_THREADS) orl_reset_kernel(DevEnv E, const uint8_t* __restrict__ mask, float* __restrict__ obs_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const DevModel* Mo; float* st; float* ob;
  if (!stage_in(E, smem, Mo, st, ob)) return;
  const int lane = threadIdx.x & 31, leg = lane & 3, rw = lane >> 2;
  const int robot = warp_row0(E) + rw;
  const int robot_c = min(robot, E.cfg.n - 1);
  float* S = st + rw * ORL_STRIDE;
  const bool doit = mask == nullptr || mask[robot_c] != 0;
  if (doit) {
    reset_robot<OPT>(S, E, *Mo, E.cfg.gid0 + robot_c, leg, 0xFu << (lane & ~3));
    reset_rings(S, E, *Mo, E.cfg.gid0 + robot_c, leg, 0xFu << (lane & ~3), robot < E.cfg.n);
    if (robot < E.cfg.n && leg == 0) queue_for_preparation(E, robot);
  }
  __syncwarp();
  build_obs<OPT>(S, ob + rw * ORL_OBS_SIZE, E, leg, E.cfg.gid0 + robot_c, pk_rows(ob) + rw * ORL_PK_STRIDE + PK_TAR, false);
  stage_out(E, st, ob, obs_out);
}

// ---- parity entry points --------------------------------------------------------------------------
__global__ void orl_motion_lookup_kernel(MotionTables T, const int32_t* clip, const double* time, int n, int32_t* f0,
                                         int32_t* f1, int32_t* cycle, double* blend, float* pose, float* vel) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = clip[i];
  const BlendIdx b = motion_blend_idx(T, c, time[i]);
  if (f0) f0[i] = b.f0;
  if (f1) f1[i] = b.f1;
  if (cycle) cycle[i] = b.cycle;
  if (blend) blend[i] = b.blend;
  if (pose) {
    float3 p; Q4 r;
    motion_root(T, c, b, motion_cycle_offset(T, c, b.cycle), p, r);
    float* o = pose + (size_t)i * 19;
    st3(o, p); stq(o + 3, r);
    for (int j = 0; j < 12; ++j) o[7 + j] = motion_joint(T, c, b, j);
  }
  if (vel) {
    float3 l, a;
    motion_root_vel(T, c, b, motion_cycle_offset(T, c, b.cycle), l, a);
    float* o = vel + (size_t)i * 18;
    st3(o, l); st3(o + 3, a);
    for (int j = 0; j < 12; ++j) o[6 + j] = motion_joint_vel(T, c, b, j);
  }
}

// one quad per query, model read from global memory (diagnostic paths, not performance critical)
__global__ void orl_fk_kernel(const DevModel* Mo, const float* pose, int n, float* com) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = min(t >> 2, n - 1), leg = t & 3;
  const bool valid = (t >> 2) < n;
  const float* p = pose + (size_t)i * 19;
  const DevLeg& L = Mo->leg[leg];
  float q[3] = {p[7 + 3 * leg], p[8 + 3 * leg], p[9 + 3 * leg]};
  LegFK K;
  leg_fk<false>(L, qmat(ldq(p + 3)), q, K);
  if (!valid) return;
  for (int l = 0; l < Mo->n_links; ++l) {
    if (Mo->link_leg[l] != leg) continue;
    const int kb = Mo->link_body[l];
    const float3 c = ld3(p) + pickp(K, kb) + mul(pickR(K, kb), f4xyz(Mo->link_com[l]));
    st3(com + ((size_t)i * Mo->n_links + l) * 3, c);
  }
}

#define ORL_DYN_BLOCK 64
// mode 0: mass matrix [n][18][18]; 1: inverse dynamics tau [n][18]; 2: forward dynamics acc [n][18]
// generalised velocity order of the interface: [lin(3), ang(3), joints(12)]
__global__ void orl_dyn_kernel(const DevModel* Mo, PhysParams P, int mode, const float* pose, const float* vel,
                               const float* in, const float* rand /* [n][12] randomizer values (S_RAND layout) or null */, int n, float* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = min(t >> 2, n - 1), leg = t & 3;
  const bool valid = (t >> 2) < n;
  const float* p = pose + (size_t)i * 19;
  const DevLeg& L = Mo->leg[leg];
  float q[3] = {p[7 + 3 * leg], p[8 + 3 * leg], p[9 + 3 * leg]};
  float3 lin = v3(0.f, 0.f, 0.f), ang = v3(0.f, 0.f, 0.f);
  float qd[3] = {0.f, 0.f, 0.f}, tau[3] = {0.f, 0.f, 0.f};
  if (vel) {
    const float* v = vel + (size_t)i * 18;
    lin = ld3(v); ang = ld3(v + 3);
    for (int k = 0; k < 3; ++k) qd[k] = v[6 + 3 * leg + k];
  }
  if (mode == 2) for (int k = 0; k < 3; ++k) tau[k] = in[(size_t)i * 12 + 3 * leg + k];
  const M3 Rb = qmat(ldq(p + 3));
  float3 alpha, alin; float qdd[3];
  DynExport ex;
  __shared__ float4 Xs[ORL_DYN_BLOCK / 32 * ORL_X_FLOAT4_PER_LANE * ORL_X_STRIDE];     // per-lane dynamics scratch
  float4* X = Xs + (threadIdx.x >> 5) * (ORL_X_FLOAT4_PER_LANE * ORL_X_STRIDE) + (threadIdx.x & 31);
  RandLane RL; RL.base_mass_ratio = rand ? rand[(size_t)i * 12 + 0] : 1.f; RL.base_inertia_ratio = rand ? rand[(size_t)i * 12 + 2] : 1.f;
  RL.mu_foot = rand ? rand[(size_t)i * 12 + 5] : -1.f;
  rand_lane_toe(L, *Mo, RL);
  {
    const float ratios[4] = {RL.base_mass_ratio, rand ? rand[(size_t)i * 12 + 1] : 1.f, RL.base_inertia_ratio, rand ? rand[(size_t)i * 12 + 3] : 1.f};
    fill_scaled_inertia(L, ratios, X + ORL_XI_BASE * ORL_X_STRIDE);
  }
  if (mode == 2) leg_dynamics<true, false, false>(L, *Mo, P, Rb, p[2], ang, lin, q, qd, tau, X, RL, alpha, alin, qdd, nullptr);
  else leg_dynamics<false, true, false>(L, *Mo, P, Rb, p[2], ang, lin, q, qd, tau, X, RL, alpha, alin, qdd, &ex);
  if (mode == 2) {
    if (!valid) return;
    float* o = out + (size_t)i * 18;
    if (leg == 0) { st3(o, alin + cross(ang, lin)); st3(o + 3, alpha); }
    for (int k = 0; k < 3; ++k) o[6 + 3 * leg + k] = qdd[k];
    return;
  }
  // interface index of internal base coordinate c (internal order: ang 0..2, lin 3..5)
  const int bi[6] = {3, 4, 5, 0, 1, 2};
  float Fc[3][6];
  for (int k = 0; k < 3; ++k) {
    Fc[k][0] = ex.Fcol[k].a.x; Fc[k][1] = ex.Fcol[k].a.y; Fc[k][2] = ex.Fcol[k].a.z;
    Fc[k][3] = ex.Fcol[k].l.x; Fc[k][4] = ex.Fcol[k].l.y; Fc[k][5] = ex.Fcol[k].l.z;
  }
  const int tri3[3][3] = {{0, 1, 3}, {1, 2, 4}, {3, 4, 5}};
  if (mode == 0) {
    if (!valid) return;
    float* M = out + (size_t)i * 324;
    if (leg == 0)
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b <= a; ++b) { M[bi[a] * 18 + bi[b]] = ex.Hbb[TRI(a, b)]; M[bi[b] * 18 + bi[a]] = ex.Hbb[TRI(a, b)]; }
    for (int k = 0; k < 3; ++k) {
      const int jk = 6 + 3 * leg + k;
      for (int c = 0; c < 6; ++c) { M[bi[c] * 18 + jk] = Fc[k][c]; M[jk * 18 + bi[c]] = Fc[k][c]; }
      for (int j = 0; j < 3; ++j) M[jk * 18 + 6 + 3 * leg + j] = ex.Hll[tri3[k][j]];
    }
    return;
  }
  // mode 1: tau = H (nudot - e) + C, e = [0; w x v; 0]
  const float* ac = in + (size_t)i * 18;
  const float3 e = cross(ang, lin);
  float ab[6] = {ac[3], ac[4], ac[5], ac[0] - e.x, ac[1] - e.y, ac[2] - e.z};
  float qa[3] = {ac[6 + 3 * leg], ac[7 + 3 * leg], ac[8 + 3 * leg]};
  float tb[6];
  for (int c = 0; c < 6; ++c) {
    float sacc = 0.f;
    for (int k = 0; k < 3; ++k) sacc += Fc[k][c] * qa[k];
    tb[c] = quad_sum(sacc);
  }
  if (!valid) return;
  float* o = out + (size_t)i * 18;
  if (leg == 0)
    for (int a = 0; a < 6; ++a) {
      float sacc = tb[a] + ex.pb[a];
      for (int b = 0; b < 6; ++b) sacc += ex.Hbb[a >= b ? TRI(a, b) : TRI(b, a)] * ab[b];
      o[bi[a]] = sacc;
    }
  for (int k = 0; k < 3; ++k) {
    float sacc = ex.C[k];
    for (int c = 0; c < 6; ++c) sacc += Fc[k][c] * ab[c];
    for (int j = 0; j < 3; ++j) sacc += ex.Hll[tri3[k][j]] * qa[j];
    o[6 + 3 * leg + k] = sacc;
  }
}

// reward terms of the current state, one quad per robot, straight from global memory
__global__ void orl_reward_terms_kernel(DevEnv E, float* terms) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = min(t >> 2, E.cfg.n - 1), leg = t & 3;
  const bool valid = (t >> 2) < E.cfg.n;
  const DevCfg& C = E.cfg;
  float* S = E.state + (size_t)i * ORL_STRIDE;
  const DevModel* Mo = E.model;
  const DevLeg& L = Mo->leg[leg];
  const float3 pos = ld3(S + S_POS), lin = ld3(S + S_LIN), ang = ld3(S + S_ANG);
  const Q4 quat = ldq(S + S_QUAT);
  const float3 rp = ld3(S + S_REFPOSE);
  const Q4 rq = ldq(S + S_REFPOSE + 3);
  float q[3], rj[3], pose_err = 0.f, vel_err = 0.f, ee_err = 0.f;
  for (int k = 0; k < 3; ++k) {
    q[k] = S[S_Q + 3 * leg + k]; rj[k] = S[S_REFPOSE + 7 + 3 * leg + k];
    const float dq = rj[k] - q[k], dv = S[S_REFVEL + 6 + 3 * leg + k] - S[S_QD + 3 * leg + k];
    pose_err += dq * dq; vel_err += dv * dv;
  }
  LegFK Ks, Kr;
  leg_fk<false>(L, qmat(quat), q, Ks);
  leg_fk<false>(L, qmat(rq), rj, Kr);
  const Q4 inv_def = qconj(ldq4(Mo->iquat));
  float ss, cs, sr, cr;
  qheading_neg_sincos(qmul(quat, inv_def), ss, cs);
  qheading_neg_sincos(qmul(rq, inv_def), sr, cr);
  for (int e = 0; e < L.meta.x; ++e) {
    const float4 ee = L.ee[e];
    const int kb = __float_as_int(ee.w);
    const float3 loc = f4xyz(ee);
    const float3 es = pickp(Ks, kb) + mul(pickR(Ks, kb), loc);
    const float3 er = pickp(Kr, kb) + mul(pickR(Kr, kb), loc);
    const float3 d = qrot_z(sr, cr, er) - qrot_z(ss, cs, es);
    const float dh = (rp.z + er.z) - (pos.z + es.z);
    ee_err += d.x * d.x + d.y * d.y + C.s_eeh * dh * dh;
  }
  pose_err = quad_sum(pose_err); vel_err = quad_sum(vel_err); ee_err = quad_sum(ee_err);
  const float3 dp = rp - pos;
  const float a = qangle_normalized(qmul(rq, qconj(quat)));
  const float root_err = dot(dp, dp) + 0.5f * a * a;
  const float3 dv = ld3(S + S_REFVEL) - lin, dw = ld3(S + S_REFVEL + 3) - ang;
  const float rootvel_err = dot(dv, dv) + 0.1f * dot(dw, dw);
  if (valid && leg == 0) {
    float* o = terms + (size_t)i * 5;
    o[0] = __expf(-C.s_pose * pose_err); o[1] = __expf(-C.s_vel * vel_err); o[2] = __expf(-C.s_ee * ee_err);
    o[3] = __expf(-C.s_root * root_err); o[4] = __expf(-C.s_rootvel * rootvel_err);
  }
}
