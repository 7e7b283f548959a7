// C ABI of the B200-native batched motion-imitation environment (include/orl.h).
// Host side: handle management, motion-table preparation (double precision, restating
// R/envs/quadruped_robot/task/motion_data.py:527-680), kernel launches.  No torch types, no exceptions
// across the boundary.
#include <cuda_runtime.h>
#include <math.h>
#include <cmath>
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/orl.h"
#include "orl_kernels.cuh"

#define ORL_D2H_CHUNKS 4
struct orl_env {
  orl_config cfg;
  orl_robot_model model;
  DevEnv dev;
  int device;
  bool motions_loaded;
  int64_t launches;
  std::string err;
  // device allocations
  float* d_state;
  float* d_next_ep; int* d_prepared; int* d_todo; int* d_todo_count;   // episodes prepared ahead of time (auto-reset only)
  int prep_parity; bool prep_scan;
  float* d_ahead; int ahead_parity; size_t ahead_bytes;   // reference packets computed one launch ahead (DevEnv::ahead)
  int num_sms;
  bool opt;                        // a task option of ABI 3 is on: the OPT instantiations of the step / reset kernels run
  DevModel* d_model;
  double* d_stats;
  std::vector<void*> d_motion;     // owned motion-table buffers
  // host copies of the motion tables (tests / diagnostics)
  std::vector<double> h_frames, h_vels, h_dur, h_dpos, h_dheading;
  std::vector<int> h_nframes;
  // pinned staging for the *_host entry points
  float* p_act; float* p_obs; float* p_rew; uint8_t* p_done; uint8_t* p_mask;
  float* d_act; float* d_obs; float* d_rew; uint8_t* d_done; uint8_t* d_mask;
  cudaStream_t stream;
  cudaEvent_t ev_chunk[ORL_D2H_CHUNKS];
  const float* parity_rand;        // orl_set_dynamics_rand
  // orl_step_host: the caller's four page-locked buffers of the last call and their device aliases (eight driver queries
  // per step otherwise: ~5 us of a 130 us call)
  const void* zc_host[4]; void* zc_dev[4];
};

static int fail(orl_env* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (e) e->err = buf;
  return code;
}
#define CUDA_TRY(e, call)                                                                              \
  do {                                                                                                 \
    cudaError_t _s = (call);                                                                           \
    if (_s != cudaSuccess) return fail(e, ORL_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_s)); \
  } while (0)

// ---- double-precision quaternion helpers (host) ------------------------------------------------------
struct Qd { double x, y, z, w; };
static Qd qd_mul(Qd a, Qd b) {
  return Qd{a.x * b.w + a.y * b.z - a.z * b.y + a.w * b.x, -a.x * b.z + a.y * b.w + a.z * b.x + a.w * b.y,
            a.x * b.y - a.y * b.x + a.z * b.w + a.w * b.z, -a.x * b.x - a.y * b.y - a.z * b.z + a.w * b.w};
}
static Qd qd_conj(Qd q) { return Qd{-q.x, -q.y, -q.z, q.w}; }
static void qd_rot(Qd q, const double v[3], double out[3]) {
  const double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  Qd p{v[0], v[1], v[2], 0.0};
  Qd qi{-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
  Qd r = qd_mul(qd_mul(q, p), qi);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
static double qd_heading(Qd q) {   // pose3d.calc_heading
  const double ex[3] = {1, 0, 0};
  double r[3];
  qd_rot(q, ex, r);
  return atan2(r[1], r[0]);
}
static Qd qd_about_z(double angle) { return Qd{0, 0, sin(angle / 2), cos(angle / 2)}; }

extern "C" int orl_abi_version(void) { return ORL_ABI_VERSION; }

extern "C" const char* orl_last_error(const orl_env* env) { return env ? env->err.c_str() : "null handle"; }

extern "C" int64_t orl_launch_count(const orl_env* env) { return env ? env->launches : 0; }

static void fill_dev_cfg(orl_env* e) {
  const orl_config& c = e->cfg;
  DevCfg& d = e->dev.cfg;
  d.n = c.num_robot; d.gid0 = c.gid0; d.seed = c.seed; d.action_repeat = c.action_repeat;
  d.max_episode_steps = c.max_episode_steps; d.auto_reset = c.auto_reset; d.cycle_sync = c.enable_cycle_sync;
  d.sim_dt = c.sim_time_step; d.env_dt = c.env_time_step; d.warmup_time = c.warmup_time; d.rsi_prob = c.ref_state_init_prob;
  for (int i = 0; i < ORL_NUM_TAR_FRAMES; ++i) d.tar_steps[i] = c.tar_frame_steps[i];
  d.w_pose = c.pose_weight; d.w_vel = c.velocity_weight; d.w_ee = c.end_effector_weight; d.w_root = c.root_pose_weight;
  d.w_rootvel = c.root_velocity_weight;
  d.s_pose = c.pose_err_scale; d.s_vel = c.velocity_err_scale; d.s_ee = c.end_effector_err_scale;
  d.s_eeh = c.end_effector_height_err_scale; d.s_root = c.root_pose_err_scale; d.s_rootvel = c.root_velocity_err_scale;
  d.dist_thr2 = c.dist_fail_threshold * c.dist_fail_threshold; d.rot_thr = c.rot_fail_threshold;
  // 2nd-order Butterworth low-pass, cutoff 4 Hz at fs = 1/(sim_dt * action_repeat), bilinear transform
  // (scipy.signal.butter(2, [4/(0.5 fs)]) as used by action_filter.py:42-44, 196-217)
  const double fs = 1.0 / (c.sim_time_step * c.action_repeat);
  const double wn = 4.0 / (0.5 * fs);
  const double K = tan(M_PI * wn / 2.0);
  const double norm = 1.0 / (1.0 + M_SQRT2 * K + K * K);
  d.fb0 = (float)(K * K * norm); d.fb1 = (float)(2.0 * K * K * norm); d.fb2 = d.fb0;
  d.fa1 = (float)(2.0 * (K * K - 1.0) * norm); d.fa2 = (float)((1.0 - M_SQRT2 * K + K * K) * norm);
  d.phys.dt = (float)c.sim_time_step; d.phys.g = -c.gravity_z;
  d.phys.kn = c.contact_kn; d.phys.bn = c.contact_bn; d.phys.inv_d0 = 1.0f / c.contact_d0; d.phys.bt = c.contact_bt;
  d.phys.klim = c.limit_k; d.phys.dlim = c.limit_d; d.phys.inv_e0 = 1.0f / c.limit_e0;
  d.phys.vmax = c.max_coordinate_velocity;
  d.phys.kfric = c.joint_friction_k;
  d.randomize = c.enable_randomizer;
  d.mu_plane = c.mu_plane;
  for (int i = 0; i < 8; ++i) { d.rand_lo[i] = c.rand_lo[i]; d.rand_hi[i] = c.rand_hi[i]; }
  d.reward_weight = c.reward_weight; d.rand_init_time = c.enable_rand_init_time;
  d.clip_tmin = c.clip_time_min; d.clip_tmax = c.clip_time_max; d.perturb_prob = c.perturb_init_state_prob;
  d.tar_noise_on = c.enable_tar_obs_noise; d.tar_noise = c.tar_obs_noise;
  e->opt = c.reward_weight != 1.0f || !c.enable_rand_init_time || (std::isfinite(c.clip_time_min) && std::isfinite(c.clip_time_max)) ||
           c.perturb_init_state_prob > 0.0 || c.enable_tar_obs_noise;
}

static float f_from_i(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
static int32_t i_from_f(float f) { int32_t i; memcpy(&i, &f, 4); return i; }

// orl_robot_model (public, plain arrays) -> DevModel (float4 packed; collision points grouped per body)
static void pack_model(const orl_robot_model& m, DevModel& d) {
  memset(&d, 0, sizeof(d));
  d.b0 = make_float4(m.base_mass, m.base_inertia[0], m.base_inertia[1], m.base_inertia[2]);
  d.b1 = make_float4(m.base_inertia[3], m.base_inertia[4], m.base_inertia[5], m.max_angle_change);
  d.ipos = make_float4(m.init_pos[0], m.init_pos[1], m.init_pos[2], 0.f);
  d.iquat = make_float4(m.init_quat[0], m.init_quat[1], m.init_quat[2], m.init_quat[3]);
  for (int l = 0; l < ORL_NUM_LEGS; ++l) {
    const orl_leg_model& L = m.leg[l];
    DevLeg& D = d.leg[l];
    float brad[4] = {0.f, 0.f, 0.f, 0.f}, mmax = 0.f;     // bodies 0,1,2 and base
    int cnt[4] = {0, 0, 0, 0}, pos = 0;
    for (int g = 0; g < 4; ++g) {
      const int body = g == 3 ? -1 : g;
      for (int i = 0; i < L.n_cp; ++i) {
        const orl_contact_point& c = L.cp[i];
        if (c.body != body) continue;
        D.cp[pos].a = make_float4(c.local[0], c.local[1], c.local[2], c.radius);
        D.cp[pos].b = make_float4(c.mu, c.margin, f_from_i(c.is_foot ? 1 : 0), 0.f);
        if (!c.is_foot && c.margin > mmax) mmax = c.margin;
        ++pos; ++cnt[g];
      }
      // every group holds at least ORL_CP_UNROLL slots (body_contact_mask tests that many straight-line);
      // padding entries can never penetrate (radius -1e30) and count as non-foot points with zero margin
      while (cnt[g] < ORL_CP_UNROLL && pos < ORL_DEV_MAX_CP) {
        D.cp[pos].a = make_float4(0.f, 0.f, 0.f, -1e30f);
        D.cp[pos].b = make_float4(0.f, 0.f, f_from_i(0), 0.f);
        ++pos; ++cnt[g];
      }
    }
    // primary contact point: the foot-link sphere on body 2 with the largest radius (the toe); it becomes the LAST
    // slot of its group, so that the straight-line penetration tests of that group can leave it out
    int prim = -1; float prad = 0.f;
    for (int i = cnt[0] + cnt[1]; i < cnt[0] + cnt[1] + cnt[2]; ++i)
      if (i_from_f(D.cp[i].b.z) != 0 && D.cp[i].a.w > prad) { prad = D.cp[i].a.w; prim = i; }
    if (prim >= 0) {
      const int last = cnt[0] + cnt[1] + cnt[2] - 1;
      const DevCp t = D.cp[prim]; D.cp[prim] = D.cp[last]; D.cp[last] = t;
      prim = last;
    }
    // bounds of every body's collision points WITHOUT the primary point (it is evaluated unconditionally):
    // bounding radius about the body origin, and the body-frame bounding box used for the per-substep test
    // "can any point of this body be below the ground?" (lowest possible z = c.z' - sum h_i |R_zi| - rmax)
    {
      const int first[4] = {0, cnt[0], cnt[0] + cnt[1], cnt[0] + cnt[1] + cnt[2]};
      for (int g = 0; g < 4; ++g) {
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f}, rmax = 0.f;
        int n = 0;
        for (int i = first[g]; i < first[g] + cnt[g]; ++i) {
          if (i == prim || D.cp[i].a.w < -1e29f) continue;
          const float4 a = D.cp[i].a;
          const float xyz[3] = {a.x, a.y, a.z};
          for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], xyz[k]); hi[k] = fmaxf(hi[k], xyz[k]); }
          rmax = fmaxf(rmax, a.w);
          const float r = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z) + a.w;
          if (r > brad[g]) brad[g] = r;
          ++n;
        }
        if (n == 0) { D.bb[g][0] = make_float4(0.f, 0.f, 0.f, -1e30f); D.bb[g][1] = make_float4(0.f, 0.f, 0.f, 0.f); }
        else {
          D.bb[g][0] = make_float4(0.5f * (lo[0] + hi[0]), 0.5f * (lo[1] + hi[1]), 0.5f * (lo[2] + hi[2]), rmax);
          D.bb[g][1] = make_float4(0.5f * (hi[0] - lo[0]), 0.5f * (hi[1] - lo[1]), 0.5f * (hi[2] - lo[2]), 0.f);
        }
      }
    }
    for (int i = pos; i < ORL_DEV_MAX_CP; ++i) D.cp[i].a = make_float4(0.f, 0.f, 0.f, -1e30f);   // unused slots never penetrate
    for (int k = 0; k < 3; ++k) {
      DevJoint& J = D.j[k];
      J.r0 = make_float4(L.R_fix[k][0], L.R_fix[k][1], L.R_fix[k][2], L.p_fix[k][0]);
      J.r1 = make_float4(L.R_fix[k][3], L.R_fix[k][4], L.R_fix[k][5], L.p_fix[k][1]);
      J.r2 = make_float4(L.R_fix[k][6], L.R_fix[k][7], L.R_fix[k][8], L.p_fix[k][2]);
      J.i0 = make_float4(L.inertia[k][0], L.inertia[k][1], L.inertia[k][2], L.inertia[k][3]);
      J.i1 = make_float4(L.inertia[k][4], L.inertia[k][5], L.mass[k], L.com[k][0]);
      J.c1 = make_float4(L.com[k][1], L.com[k][2], L.lower[k], L.upper[k]);
      J.m0 = make_float4(L.damping[k], L.dir[k], L.offset[k], L.kp[k]);
      J.m1 = make_float4(L.kd[k], L.init_angle[k], f_from_i(L.motor[k]), brad[k]);
      J.i2 = make_float4(L.inertia_par[k][0], L.inertia_par[k][1], L.inertia_par[k][2], L.inertia_par[k][3]);
      J.i3 = make_float4(L.inertia_par[k][4], L.inertia_par[k][5], f_from_i(L.rand_class[k]), f_from_i(L.friction_draw[k]));
    }
    for (int e = 0; e < 2; ++e)
      D.ee[e] = make_float4(L.ee_local[e][0], L.ee_local[e][1], L.ee_local[e][2], f_from_i(L.ee_body[e]));
    D.meta = make_int4(L.n_ee, i_from_f(brad[3]), i_from_f(mmax), prim);
    D.cnt = make_int4(cnt[0], cnt[1], cnt[2], cnt[3]);
  }
  d.n_links = m.n_links;
  for (int i = 0; i < ORL_MAX_LINKS; ++i) {
    d.link_leg[i] = m.link_leg[i]; d.link_body[i] = m.link_body[i];
    d.link_com[i] = make_float4(m.link_com[i][0], m.link_com[i][1], m.link_com[i][2], 0.f);
  }
}

extern "C" int orl_create(const orl_config* cfg, const orl_robot_model* model, orl_env** out) {
  if (!cfg || !model || !out) return ORL_ERR_INVALID;
  *out = nullptr;
  orl_env* e = new (std::nothrow) orl_env();
  if (!e) return ORL_ERR_INVALID;
  e->cfg = *cfg; e->model = *model; e->motions_loaded = false; e->launches = 0;
  e->d_state = nullptr; e->d_model = nullptr; e->d_stats = nullptr; e->d_next_ep = nullptr; e->d_prepared = nullptr; e->d_todo = nullptr; e->d_todo_count = nullptr; e->prep_parity = 0; e->prep_scan = true; e->d_ahead = nullptr; e->ahead_parity = 0; e->ahead_bytes = 0; e->num_sms = 0;
  e->p_act = e->p_obs = e->p_rew = nullptr; e->p_done = e->p_mask = nullptr;
  e->d_act = e->d_obs = e->d_rew = nullptr; e->d_done = e->d_mask = nullptr;
  e->stream = nullptr;
  for (int i = 0; i < 4; ++i) { e->zc_host[i] = nullptr; e->zc_dev[i] = nullptr; }
  *out = e;   // returned even on failure so that orl_last_error works; the caller destroys it
  if (cfg->abi_version != ORL_ABI_VERSION) return fail(e, ORL_ERR_INVALID, "ABI version %d != %d", cfg->abi_version, ORL_ABI_VERSION);
  if (cfg->num_robot <= 0) return fail(e, ORL_ERR_INVALID, "num_robot must be positive");
  if (cfg->action_repeat < 3) return fail(e, ORL_ERR_INVALID, "action_repeat must be >= 3 (control latency of 2 substeps)");
  if (!(cfg->sim_time_step > 0)) return fail(e, ORL_ERR_INVALID, "sim_time_step must be positive");
  if (model->n_links < 0 || model->n_links > ORL_MAX_LINKS) return fail(e, ORL_ERR_INVALID, "bad n_links");
  for (int l = 0; l < ORL_NUM_LEGS; ++l) {
    if (model->leg[l].n_cp < 0 || model->leg[l].n_cp > ORL_MAX_CP) return fail(e, ORL_ERR_INVALID, "bad n_cp on leg %d", l);
    if (model->leg[l].n_ee < 0 || model->leg[l].n_ee > 2) return fail(e, ORL_ERR_INVALID, "bad n_ee on leg %d", l);
    for (int k = 0; k < 3; ++k)
      if (model->leg[l].motor[k] < 0 || model->leg[l].motor[k] >= 12) return fail(e, ORL_ERR_INVALID, "bad motor index");
    int per_body[4] = {0, 0, 0, 0};
    for (int i = 0; i < model->leg[l].n_cp; ++i) {
      if (model->leg[l].cp[i].body < -1 || model->leg[l].cp[i].body > 2) return fail(e, ORL_ERR_INVALID, "bad collision-point body");
      ++per_body[model->leg[l].cp[i].body + 1];
    }
    int slots = 0;
    for (int g = 0; g < 4; ++g) slots += per_body[g] < ORL_CP_UNROLL ? ORL_CP_UNROLL : per_body[g];
    if (slots > ORL_DEV_MAX_CP) return fail(e, ORL_ERR_INVALID, "leg %d: collision points need %d device slots (max %d)", l, slots, ORL_DEV_MAX_CP);
    for (int i = 0; i < model->leg[l].n_ee; ++i)
      if (model->leg[l].ee_body[i] < 0 || model->leg[l].ee_body[i] > 2) return fail(e, ORL_ERR_INVALID, "bad end-effector body");
  }
  if (!(cfg->contact_d0 > 0) || !(cfg->limit_e0 > 0)) return fail(e, ORL_ERR_INVALID, "contact_d0 and limit_e0 must be positive");
  if (!(cfg->max_coordinate_velocity > 0)) return fail(e, ORL_ERR_INVALID, "max_coordinate_velocity must be positive");
  if (std::isfinite(cfg->clip_time_min) && std::isfinite(cfg->clip_time_max) && !(cfg->clip_time_max >= cfg->clip_time_min && cfg->clip_time_min >= 0.0))
    return fail(e, ORL_ERR_INVALID, "clip_time_min / clip_time_max must satisfy 0 <= min <= max");
  if (cfg->clip_time_min != cfg->clip_time_min || cfg->clip_time_max != cfg->clip_time_max || !(cfg->perturb_init_state_prob >= 0.0))
    return fail(e, ORL_ERR_INVALID, "bad task option (NaN clip time or negative perturbation probability)");
  if (cfg->enable_randomizer) {
    for (int i = 0; i < 8; ++i)
      if (!(cfg->rand_hi[i] >= cfg->rand_lo[i])) return fail(e, ORL_ERR_INVALID, "randomizer range %d is empty", i);
    if (!(cfg->rand_lo[ORL_RAND_MASS] > 0) || !(cfg->rand_lo[ORL_RAND_INERTIA] > 0)) return fail(e, ORL_ERR_INVALID, "mass / inertia ratios must be positive");
    for (int l = 0; l < ORL_NUM_LEGS; ++l)
      for (int k = 0; k < 3; ++k) {
        if (model->leg[l].rand_class[k] < 0) return fail(e, ORL_ERR_INVALID, "leg %d body %d merges links of the chassis and the leg set: randomizer unsupported", l, k);
        if (model->leg[l].friction_draw[k] >= 8) return fail(e, ORL_ERR_INVALID, "bad friction_draw");
      }
  }
  e->device = cfg->device;
  CUDA_TRY(e, cudaSetDevice(e->device));
  const size_t nstate = (size_t)cfg->num_robot * ORL_STRIDE;
  CUDA_TRY(e, cudaMalloc(&e->d_state, nstate * sizeof(float)));
  CUDA_TRY(e, cudaMemset(e->d_state, 0, nstate * sizeof(float)));
  CUDA_TRY(e, cudaMalloc(&e->d_model, ORL_SMEM_MODEL_BYTES));
  CUDA_TRY(e, cudaMemset(e->d_model, 0, ORL_SMEM_MODEL_BYTES));
  {
    DevModel packed;
    pack_model(*model, packed);
    CUDA_TRY(e, cudaMemcpy(e->d_model, &packed, sizeof(DevModel), cudaMemcpyHostToDevice));
  }
  CUDA_TRY(e, cudaMalloc(&e->d_stats, 8 * sizeof(double)));
  CUDA_TRY(e, cudaMemset(e->d_stats, 0, 8 * sizeof(double)));
  // motor strengths default to 1, orientation of the control observation to identity
  {
    std::vector<float> init(nstate, 0.f);
    for (int r = 0; r < cfg->num_robot; ++r) {
      float* S = init.data() + (size_t)r * ORL_STRIDE;
      for (int m = 0; m < 12; ++m) S[S_STRENGTH + m] = 1.f;
      S[S_OBSQ + 3] = 1.f; S[S_OROT + 3] = 1.f; S[S_QUAT + 3] = 1.f;
      S[S_RAND + 0] = S[S_RAND + 1] = S[S_RAND + 2] = S[S_RAND + 3] = 1.f; S[S_RAND + 5] = -1.f;
      const double never = INFINITY;
      memcpy(S + S_CHANGET, &never, sizeof(double));
    }
    CUDA_TRY(e, cudaMemcpy(e->d_state, init.data(), nstate * sizeof(float), cudaMemcpyHostToDevice));
  }
  const int N = cfg->num_robot;
  CUDA_TRY(e, cudaMallocHost(&e->p_act, (size_t)N * 12 * sizeof(float)));
  CUDA_TRY(e, cudaMallocHost(&e->p_obs, (size_t)N * ORL_OBS_SIZE * sizeof(float)));
  CUDA_TRY(e, cudaMallocHost(&e->p_rew, (size_t)N * sizeof(float)));
  CUDA_TRY(e, cudaMallocHost(&e->p_done, (size_t)N));
  CUDA_TRY(e, cudaMallocHost(&e->p_mask, (size_t)N));
  CUDA_TRY(e, cudaMalloc(&e->d_act, (size_t)N * 12 * sizeof(float)));
  CUDA_TRY(e, cudaMalloc(&e->d_obs, (size_t)N * ORL_OBS_SIZE * sizeof(float)));
  CUDA_TRY(e, cudaMalloc(&e->d_rew, (size_t)N * sizeof(float)));
  CUDA_TRY(e, cudaMalloc(&e->d_done, (size_t)N));
  CUDA_TRY(e, cudaMalloc(&e->d_mask, (size_t)N));
  CUDA_TRY(e, cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  for (int i = 0; i < ORL_D2H_CHUNKS; ++i) CUDA_TRY(e, cudaEventCreateWithFlags(&e->ev_chunk[i], cudaEventDisableTiming));
  CUDA_TRY(e, cudaFuncSetAttribute(orl_step_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ORL_SMEM_BYTES));
  CUDA_TRY(e, cudaFuncSetAttribute(orl_step_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ORL_SMEM_BYTES));
  CUDA_TRY(e, cudaFuncSetAttribute(orl_step_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ORL_SMEM_BYTES));
  CUDA_TRY(e, cudaFuncSetAttribute(orl_step_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ORL_SMEM_BYTES));
  CUDA_TRY(e, cudaFuncSetAttribute(orl_reset_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ORL_SMEM_BYTES));
  CUDA_TRY(e, cudaFuncSetAttribute(orl_reset_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ORL_SMEM_BYTES));
  e->dev.state = e->d_state; e->dev.model = e->d_model; e->dev.stats = e->d_stats;
  e->dev.ring_q = nullptr; e->dev.ring_b = nullptr;
  e->dev.next_ep = nullptr; e->dev.prepared = nullptr; e->dev.prep_blocks = 0;
  e->dev.todo = nullptr; e->dev.todo_count = nullptr; e->dev.prep_parity = 0; e->dev.prep_scan = 0; e->dev.obs_split = 0;
  if (cfg->auto_reset) {
    CUDA_TRY(e, cudaMalloc(&e->d_next_ep, (size_t)N * ORL_NEXT_STRIDE * sizeof(float)));
    CUDA_TRY(e, cudaMalloc(&e->d_prepared, (size_t)N * sizeof(int)));
    CUDA_TRY(e, cudaMemset(e->d_prepared, 0, (size_t)N * sizeof(int)));
    CUDA_TRY(e, cudaMalloc(&e->d_todo, (size_t)4 * N * sizeof(int)));
    CUDA_TRY(e, cudaMalloc(&e->d_todo_count, 2 * sizeof(int)));
    CUDA_TRY(e, cudaMemset(e->d_todo_count, 0, 2 * sizeof(int)));
    e->dev.next_ep = e->d_next_ep; e->dev.prepared = e->d_prepared; e->dev.todo = e->d_todo; e->dev.todo_count = e->d_todo_count;
  }
  e->dev.ahead = nullptr; e->dev.ahead_parity = 0; e->dev.ahead_rows = 0;
  {
    const char* off = getenv("ORL_AHEAD");                         // knob: ORL_AHEAD=0 computes every packet inline (tests compare the two)
    CUDA_TRY(e, cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, e->device));
    if (!(off && atoi(off) == 0)) {
      const size_t rows = (((size_t)N + ORL_ROBOTS_PER_WARP - 1) / ORL_ROBOTS_PER_WARP) * ORL_ROBOTS_PER_WARP;
      e->ahead_bytes = 3 * rows * ORL_PK_STRIDE * sizeof(float);
      CUDA_TRY(e, cudaMalloc(&e->d_ahead, e->ahead_bytes));
      CUDA_TRY(e, cudaMemset(e->d_ahead, 0, e->ahead_bytes));        // key 0 matches no record that has been stepped
      e->dev.ahead = e->d_ahead; e->dev.ahead_rows = (int)rows;
    }
  }
  if (cfg->enable_randomizer) {      // observation history for the randomised control latency
    // one extra row: scratch for the padding quads of a batch that is not a multiple of 8 robots
    CUDA_TRY(e, cudaMalloc(&e->dev.ring_q, (size_t)(N + 1) * ORL_LAT_RING * 12 * sizeof(float)));
    CUDA_TRY(e, cudaMalloc(&e->dev.ring_b, (size_t)(N + 1) * ORL_LAT_RING * 8 * sizeof(float)));
    CUDA_TRY(e, cudaMemset(e->dev.ring_q, 0, (size_t)(N + 1) * ORL_LAT_RING * 12 * sizeof(float)));
    CUDA_TRY(e, cudaMemset(e->dev.ring_b, 0, (size_t)(N + 1) * ORL_LAT_RING * 8 * sizeof(float)));
  }
  memset(&e->dev.mt, 0, sizeof(e->dev.mt));
  fill_dev_cfg(e);
  return ORL_OK;
}

static void free_motions(orl_env* e) {
  for (void* p : e->d_motion) cudaFree(p);
  e->d_motion.clear();
}

extern "C" int orl_destroy(orl_env* e) {
  if (!e) return ORL_OK;
  cudaSetDevice(e->device);
  free_motions(e);
  cudaFree(e->d_state); cudaFree(e->d_model); cudaFree(e->d_stats); cudaFree(e->dev.ring_q); cudaFree(e->dev.ring_b);
  cudaFree(e->d_next_ep); cudaFree(e->d_prepared); cudaFree(e->d_todo); cudaFree(e->d_todo_count);
  cudaFree(e->d_ahead);
  cudaFreeHost(e->p_act); cudaFreeHost(e->p_obs); cudaFreeHost(e->p_rew); cudaFreeHost(e->p_done); cudaFreeHost(e->p_mask);
  cudaFree(e->d_act); cudaFree(e->d_obs); cudaFree(e->d_rew); cudaFree(e->d_done); cudaFree(e->d_mask);
  for (int i = 0; i < ORL_D2H_CHUNKS; ++i) if (e->ev_chunk[i]) cudaEventDestroy(e->ev_chunk[i]);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return ORL_OK;
}

template <typename T>
static int upload(orl_env* e, const std::vector<T>& h, const T** out) {
  void* d = nullptr;
  CUDA_TRY(e, cudaMalloc(&d, h.size() * sizeof(T) + 16));
  e->d_motion.push_back(d);
  CUDA_TRY(e, cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (const T*)d;
  return ORL_OK;
}

extern "C" int orl_load_motions(orl_env* e, const double* frames, const int32_t* n_frames, const double* frame_dur,
                                const int32_t* flags, int32_t n_clips) {
  if (!e) return ORL_ERR_INVALID;
  if (n_clips <= 0) return fail(e, ORL_ERR_INVALID, "No reference motions specified.");   // imitation_task.py:583-584
  if (n_clips > ORL_MAX_CLIPS) return fail(e, ORL_ERR_INVALID, "at most %d motion clips per environment (%d given)", ORL_MAX_CLIPS, n_clips);
  if (!frames || !n_frames || !frame_dur || !flags) return fail(e, ORL_ERR_INVALID, "null argument");
  CUDA_TRY(e, cudaSetDevice(e->device));
  free_motions(e);
  int total = 0;
  for (int c = 0; c < n_clips; ++c) {
    if (n_frames[c] <= 0) return fail(e, ORL_ERR_ASSERT, "Must have at least 1 frame.");        // motion_data.py:105
    if (!(frame_dur[c] > 0)) return fail(e, ORL_ERR_ASSERT, "Frame duration must be positive.");  // :108
    total += n_frames[c];
  }
  const int FS = 19, VS = 18;
  e->h_frames.assign(frames, frames + (size_t)total * FS);
  e->h_vels.assign((size_t)total * VS, 0.0);
  e->h_dur.resize(n_clips); e->h_dpos.resize((size_t)n_clips * 3); e->h_dheading.resize(n_clips);
  e->h_nframes.assign(n_frames, n_frames + n_clips);
  std::vector<int32_t> h_start(n_clips), h_nf(n_clips), h_flags(n_clips);
  std::vector<double> h_fd(n_clips);
  std::vector<float> h_cyc((size_t)n_clips * ORL_CYCLE_TABLE * 3, 0.f), h_warm((size_t)n_clips * ORL_FRAME_STRIDE, 0.f);
  std::vector<float> h_dposf((size_t)n_clips * 3);
  const orl_robot_model& M = e->model;
  int off = 0;
  for (int c = 0; c < n_clips; ++c) {
    const int F = n_frames[c];
    double* fr = e->h_frames.data() + (size_t)off * FS;
    double* vl = e->h_vels.data() + (size_t)off * VS;
    // _postprocess_frames, motion_data.py:527-556
    const double x0 = fr[0], y0 = fr[1];
    for (int f = 0; f < F; ++f) {
      double* r = fr + (size_t)f * FS;
      r[0] -= x0; r[1] -= y0;
      const double n = sqrt(r[3] * r[3] + r[4] * r[4] + r[5] * r[5] + r[6] * r[6]);
      if (fabs(n) <= 1e-8) return fail(e, ORL_ERR_INVALID, "Quaternion may not be zero in QuaternionNormalize");   // pose3d.py:103-107
      for (int k = 3; k < 7; ++k) r[k] /= n;
      if (r[6] < 0) for (int k = 3; k < 7; ++k) r[k] = -r[k];
    }
    // _calc_frame_vels, motion_data.py:635-680
    const double dt = frame_dur[c];
    for (int f = 0; f + 1 < F; ++f) {
      const double* a = fr + (size_t)f * FS;
      const double* b = fr + (size_t)(f + 1) * FS;
      double* v = vl + (size_t)f * VS;
      for (int k = 0; k < 3; ++k) v[k] = (b[k] - a[k]) / dt;
      Qd dq = qd_mul(Qd{b[3], b[4], b[5], b[6]}, qd_conj(Qd{a[3], a[4], a[5], a[6]}));
      // pose3d.QuaternionToAxisAngle, pose3d.py:139-187
      double ax[3] = {dq.x, dq.y, dq.z};
      const double an = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
      if (an < 1e-8) { ax[0] = 0; ax[1] = 0; ax[2] = 1; } else { ax[0] /= an; ax[1] /= an; ax[2] /= an; }
      const double angle = atan2(an, dq.w) * 2;
      for (int k = 0; k < 3; ++k) v[3 + k] = (angle / dt) * ax[k];
      for (int k = 0; k < 12; ++k) v[6 + k] = (b[7 + k] - a[7 + k]) / dt;
    }
    if (F > 1) memcpy(vl + (size_t)(F - 1) * VS, vl + (size_t)(F - 2) * VS, VS * sizeof(double));
    // cycle deltas, motion_data.py:558-589
    const double* first = fr;
    const double* last = fr + (size_t)(F - 1) * FS;
    double dp[3] = {last[0] - first[0], last[1] - first[1], 0.0};
    Qd drot = qd_mul(Qd{last[3], last[4], last[5], last[6]}, qd_conj(Qd{first[3], first[4], first[5], first[6]}));
    const double dh = qd_heading(drot);
    e->h_dur[c] = dt * (F - 1);       // motion_data.py:205-207
    for (int k = 0; k < 3; ++k) { e->h_dpos[c * 3 + k] = dp[k]; h_dposf[c * 3 + k] = (float)dp[k]; }
    e->h_dheading[c] = dh;
    h_start[c] = off; h_nf[c] = F; h_flags[c] = flags[c]; h_fd[c] = dt;
    // pre-summed rotating cycle offsets, motion_data.py:607-613
    {
      double acc[3] = {0, 0, 0};
      for (int i = 0; i < ORL_CYCLE_TABLE; ++i) {
        for (int k = 0; k < 3; ++k) h_cyc[((size_t)c * ORL_CYCLE_TABLE + i) * 3 + k] = (float)acc[k];
        double cur[3];
        qd_rot(qd_about_z(i * dh), dp, cur);
        for (int k = 0; k < 3; ++k) acc[k] += cur[k];
      }
    }
    // warm-up pose, imitation_task.py:985-1009 (pose0 = calc_frame(0) = frame 0 with no cycle offset)
    {
      Qd r0{first[3], first[4], first[5], first[6]};
      Qd rdef{M.init_quat[0], M.init_quat[1], M.init_quat[2], M.init_quat[3]};
      const double delta = qd_heading(r0) - qd_heading(rdef);
      Qd drq = qd_about_z(delta);
      double p[3] = {M.init_pos[0], M.init_pos[1], M.init_pos[2]}, pr[3];
      qd_rot(drq, p, pr);
      Qd rr = qd_mul(drq, rdef);
      float* w = h_warm.data() + (size_t)c * ORL_FRAME_STRIDE;
      w[0] = (float)pr[0]; w[1] = (float)pr[1]; w[2] = (float)pr[2];
      w[3] = (float)rr.x; w[4] = (float)rr.y; w[5] = (float)rr.z; w[6] = (float)rr.w;
      for (int l = 0; l < 4; ++l)
        for (int k = 0; k < 3; ++k) {
          const orl_leg_model& L = M.leg[l];
          w[7 + 3 * l + k] = (L.init_angle[k] + L.offset[k]) * L.dir[k];    // minitaur.py:1190-1194
        }
    }
    off += F;
  }
  std::vector<float> f32((size_t)total * ORL_FRAME_STRIDE, 0.f), v32((size_t)total * ORL_FRAME_STRIDE, 0.f);
  for (int f = 0; f < total; ++f) {
    for (int k = 0; k < FS; ++k) f32[(size_t)f * ORL_FRAME_STRIDE + k] = (float)e->h_frames[(size_t)f * FS + k];
    for (int k = 0; k < VS; ++k) v32[(size_t)f * ORL_FRAME_STRIDE + k] = (float)e->h_vels[(size_t)f * VS + k];
  }
  MotionTables& T = e->dev.mt;
  int rc;
  if ((rc = upload(e, f32, &T.frames))) return rc;
  if ((rc = upload(e, v32, &T.vels))) return rc;
  for (int c = 0; c < n_clips; ++c) {
    T.start[c] = h_start[c]; T.nframes[c] = h_nf[c]; T.flags[c] = h_flags[c];
    T.fd[c] = h_fd[c]; T.dur[c] = e->h_dur[c]; T.dheading[c] = e->h_dheading[c];
    for (int k = 0; k < 3; ++k) T.dpos[c * 3 + k] = h_dposf[c * 3 + k];
  }
  if ((rc = upload(e, h_cyc, &T.cyc_table))) return rc;
  if ((rc = upload(e, h_warm, &T.warm_pose))) return rc;
  T.n_clips = n_clips;
  if (e->d_prepared) { CUDA_TRY(e, cudaMemset(e->d_prepared, 0, (size_t)e->cfg.num_robot * sizeof(int))); e->prep_scan = true; }   // other clips, other episodes
  if (e->d_ahead) CUDA_TRY(e, cudaMemset(e->d_ahead, 0, e->ahead_bytes));   // and other packets
  e->motions_loaded = true;
  return ORL_OK;
}

extern "C" int orl_get_motion_tables(orl_env* e, double* frames_out, double* vels_out, double* duration, double* delta_pos,
                                     double* delta_heading) {
  if (!e) return ORL_ERR_INVALID;
  if (!e->motions_loaded) return fail(e, ORL_ERR_STATE, "motions not loaded");
  if (frames_out) memcpy(frames_out, e->h_frames.data(), e->h_frames.size() * sizeof(double));
  if (vels_out) memcpy(vels_out, e->h_vels.data(), e->h_vels.size() * sizeof(double));
  if (duration) memcpy(duration, e->h_dur.data(), e->h_dur.size() * sizeof(double));
  if (delta_pos) memcpy(delta_pos, e->h_dpos.data(), e->h_dpos.size() * sizeof(double));
  if (delta_heading) memcpy(delta_heading, e->h_dheading.data(), e->h_dheading.size() * sizeof(double));
  return ORL_OK;
}

#define NEED_MOTIONS(e)                                                            \
  do {                                                                             \
    if (!(e)) return ORL_ERR_INVALID;                                              \
    if (!(e)->motions_loaded) return fail(e, ORL_ERR_STATE, "motions not loaded"); \
  } while (0)

static int check_launch(orl_env* e, const char* what) {
  cudaError_t s = cudaGetLastError();
  if (s != cudaSuccess) return fail(e, ORL_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(s));
  e->launches++;
  return ORL_OK;
}

extern "C" int orl_reset(orl_env* e, const uint8_t* mask_dev, float* obs_dev, void* stream) {
  NEED_MOTIONS(e);
  CUDA_TRY(e, cudaSetDevice(e->device));       // the caller's current device may be another GPU of the box
  const int blocks = (e->cfg.num_robot + ORL_ROBOTS_PER_BLOCK - 1) / ORL_ROBOTS_PER_BLOCK;
  if ((uintptr_t)obs_dev & 15) return fail(e, ORL_ERR_INVALID, "observation buffer must be 16-byte aligned");
  DevEnv d = e->dev;
  d.prep_parity = e->prep_parity ^ 1;          // queue_for_preparation appends to list prep_parity ^ 1 = the one the next step launch reads
  if (e->opt) orl_reset_kernel<true><<<blocks, ORL_BLOCK_THREADS, ORL_SMEM_BYTES, (cudaStream_t)stream>>>(d, mask_dev, obs_dev);
  else orl_reset_kernel<false><<<blocks, ORL_BLOCK_THREADS, ORL_SMEM_BYTES, (cudaStream_t)stream>>>(d, mask_dev, obs_dev);
  return check_launch(e, "orl_reset_kernel");
}

static int step_launch(orl_env* e, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, void* stream, bool obs_split);
extern "C" int orl_step(orl_env* e, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, void* stream) {
  return step_launch(e, actions_dev, obs_dev, reward_dev, done_dev, stream, false);
}
static int step_launch(orl_env* e, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, void* stream, bool obs_split) {
  NEED_MOTIONS(e);
  if (!actions_dev || !obs_dev || !reward_dev || !done_dev) return fail(e, ORL_ERR_INVALID, "null buffer");
  if ((uintptr_t)obs_dev & 15) return fail(e, ORL_ERR_INVALID, "observation buffer must be 16-byte aligned");
  CUDA_TRY(e, cudaSetDevice(e->device));
  const int blocks = (e->cfg.num_robot + ORL_ROBOTS_PER_BLOCK - 1) / ORL_ROBOTS_PER_BLOCK;
  DevEnv d = e->dev;
  // leading blocks that prepare next episodes (they get the lowest block indices: scheduled first when the grid needs several waves)
  d.prep_blocks = d.next_ep ? std::min(16, std::max(1, (e->cfg.num_robot + 255) / 256)) : 0;
  if (e->opt) d.ahead = nullptr;                                  // (clip switching moves the clock inside the step)
  if (d.ahead) {
    // Reference packet rows of the next launch: one quad-pass per robot (~13 k cycles per chunk of 8 robots: cold FP64 code),
    // and the launch lasts until the leading blocks are done.  Two cases gain (profiles/r02_step_kernel.md section 9):
    //  * the workers leave ENOUGH SMs idle for the rows to take well under a worker's time (one leading block per 28 chunks:
    //    4096 robots = 128 worker blocks + 20 idle SMs for 512 chunks: 0.0844 -> 0.0831 ms; with 8 blocks the same launch
    //    takes 0.108 ms, with 4 blocks 0.18 ms: it waits for them) -- all idle SMs help;
    //  * the grid needs several waves (16384 robots: 0.2416 -> 0.2320 ms): one leading block per 40 chunks.
    // Leading blocks that share SMs with the worker blocks of the ONLY wave evict the substep loop from those SMs'
    // instruction caches and the launch waits for those workers (8192 robots: 0.1351 -> 0.1375 ms; sleeping until the first
    // workers are done did not help either): no rows there, the workers evaluate the reference themselves.
    const char* force = getenv("ORL_AHEAD");
    const int idle = e->num_sms - blocks, chunks = (e->cfg.num_robot + ORL_ROBOTS_PER_WARP - 1) / ORL_ROBOTS_PER_WARP;
    const int need = (chunks + 27) / 28;
    if (idle >= need) d.prep_blocks = std::min(idle, std::max(32, need));
    else if (blocks > ORL_STEP_MIN_BLOCKS * e->num_sms || (force && atoi(force) == 2))
      d.prep_blocks = std::max(d.prep_blocks, std::min(64, std::max(1, chunks / 40)));
    else d.ahead = nullptr;
  }
  { static const char* ov = getenv("ORL_PREP_BLOCKS"); if (ov && (d.next_ep || d.ahead)) d.prep_blocks = atoi(ov); }   // experiment knob
  if (d.prep_blocks == 0) { d.next_ep = nullptr; d.prepared = nullptr; d.todo = nullptr; d.ahead = nullptr; }
  d.ahead_parity = e->ahead_parity; e->ahead_parity ^= 1;         // this launch reads the half the previous one wrote
  d.prep_parity = e->prep_parity; d.prep_scan = e->prep_scan ? 1 : 0;
  d.obs_split = obs_split ? 1 : 0;
  e->prep_parity ^= 1; e->prep_scan = false;
  const dim3 grid(blocks + d.prep_blocks);
  cudaStream_t st = (cudaStream_t)stream;
  if (e->opt) {
    if (e->dev.ring_q) orl_step_kernel<true, true><<<grid, ORL_BLOCK_THREADS, ORL_SMEM_BYTES, st>>>(d, actions_dev, obs_dev, reward_dev, done_dev);
    else orl_step_kernel<false, true><<<grid, ORL_BLOCK_THREADS, ORL_SMEM_BYTES, st>>>(d, actions_dev, obs_dev, reward_dev, done_dev);
  } else {
    if (e->dev.ring_q) orl_step_kernel<true, false><<<grid, ORL_BLOCK_THREADS, ORL_SMEM_BYTES, st>>>(d, actions_dev, obs_dev, reward_dev, done_dev);
    else orl_step_kernel<false, false><<<grid, ORL_BLOCK_THREADS, ORL_SMEM_BYTES, st>>>(d, actions_dev, obs_dev, reward_dev, done_dev);
  }
  return check_launch(e, "orl_step_kernel");
}

// True when the CUDA runtime can DMA to/from this host pointer directly (cudaMallocHost / cudaHostRegister /
// torch pin_memory): such buffers skip the internal pinned staging copy.
static bool host_ptr_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// Device -> host copy of `bytes`: direct when dst is pinned, else through the pinned staging buffer in chunks so
// that the CPU memcpy of chunk k overlaps the DMA of chunk k+1.  The caller synchronises the stream afterwards
// only in the direct case (the staged case has consumed every chunk on return).
static int d2h(orl_env* e, void* dst, void* staging, const void* src, size_t bytes, bool direct) {
  if (direct) { CUDA_TRY(e, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream)); return ORL_OK; }
  const int NC = bytes >= (1u << 20) ? ORL_D2H_CHUNKS : 1;
  const size_t chunk = ((bytes + NC - 1) / NC + 255) & ~(size_t)255;
  int used = 0;
  for (size_t off = 0; off < bytes; off += chunk, ++used) {
    const size_t nb = bytes - off < chunk ? bytes - off : chunk;
    CUDA_TRY(e, cudaMemcpyAsync((char*)staging + off, (const char*)src + off, nb, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaEventRecord(e->ev_chunk[used], e->stream));
  }
  used = 0;
  for (size_t off = 0; off < bytes; off += chunk, ++used) {
    const size_t nb = bytes - off < chunk ? bytes - off : chunk;
    CUDA_TRY(e, cudaEventSynchronize(e->ev_chunk[used]));
    memcpy((char*)dst + off, (const char*)staging + off, nb);
  }
  return ORL_OK;
}

extern "C" int orl_reset_host(orl_env* e, const uint8_t* mask, float* obs) {
  NEED_MOTIONS(e);
  CUDA_TRY(e, cudaSetDevice(e->device));
  const int N = e->cfg.num_robot;
  if (mask) {
    memcpy(e->p_mask, mask, N);
    CUDA_TRY(e, cudaMemcpyAsync(e->d_mask, e->p_mask, N, cudaMemcpyHostToDevice, e->stream));
  }
  int rc = orl_reset(e, mask ? e->d_mask : nullptr, e->d_obs, e->stream);
  if (rc) return rc;
  if (obs && (rc = d2h(e, obs, e->p_obs, e->d_obs, (size_t)N * ORL_OBS_SIZE * sizeof(float), host_ptr_pinned(obs)))) return rc;
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  return ORL_OK;
}

extern "C" int orl_step_host(orl_env* e, const float* actions, float* obs, float* reward, uint8_t* done) {
  NEED_MOTIONS(e);
  if (!actions || !obs || !reward || !done) return fail(e, ORL_ERR_INVALID, "null buffer");
  CUDA_TRY(e, cudaSetDevice(e->device));
  const int N = e->cfg.num_robot;
  const size_t ab = (size_t)N * 12 * sizeof(float), ob = (size_t)N * ORL_OBS_SIZE * sizeof(float);
  void* chk = nullptr;
  if (e->zc_host[0] == actions && e->zc_host[1] == obs && e->zc_host[2] == reward && e->zc_host[3] == done &&
      cudaHostGetDevicePointer(&chk, obs, 0) == cudaSuccess && chk == e->zc_dev[1]) {
    // same page-locked buffers as in the previous call (the largest one re-checked): their device aliases are known
    int rc0 = step_launch(e, (const float*)e->zc_dev[0], (float*)e->zc_dev[1], (float*)e->zc_dev[2], (uint8_t*)e->zc_dev[3], e->stream, true);
    if (rc0) return rc0;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return ORL_OK;
  }
  cudaGetLastError();                  // (a failed re-check above leaves a sticky-less error behind)
  e->zc_host[0] = nullptr;
  const bool pin_a = host_ptr_pinned(actions), pin_o = host_ptr_pinned(obs), pin_r = host_ptr_pinned(reward), pin_d = host_ptr_pinned(done);
  if (pin_a && pin_o && pin_r && pin_d && !((uintptr_t)obs & 15)) {
    // Zero-copy: every buffer is page-locked, so the step kernel reads the actions and writes observation / reward /
    // done straight through the mapped host pointers.  The PCIe transfers then overlap the kernel's own tail
    // (warps finish at different times) instead of being serialised behind it as separate copies.
    void *da, *dob, *dr, *dd;
    CUDA_TRY(e, cudaHostGetDevicePointer(&da, (void*)actions, 0));
    CUDA_TRY(e, cudaHostGetDevicePointer(&dob, obs, 0));
    CUDA_TRY(e, cudaHostGetDevicePointer(&dr, reward, 0));
    CUDA_TRY(e, cudaHostGetDevicePointer(&dd, done, 0));
    e->zc_host[0] = actions; e->zc_host[1] = obs; e->zc_host[2] = reward; e->zc_host[3] = done;
    e->zc_dev[0] = da; e->zc_dev[1] = dob; e->zc_dev[2] = dr; e->zc_dev[3] = dd;
    int rc0 = step_launch(e, (const float*)da, (float*)dob, (float*)dr, (uint8_t*)dd, e->stream, true);
    if (rc0) return rc0;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return ORL_OK;
  }
  if (pin_a) {
    CUDA_TRY(e, cudaMemcpyAsync(e->d_act, actions, ab, cudaMemcpyHostToDevice, e->stream));
  } else {
    memcpy(e->p_act, actions, ab);
    CUDA_TRY(e, cudaMemcpyAsync(e->d_act, e->p_act, ab, cudaMemcpyHostToDevice, e->stream));
  }
  int rc = orl_step(e, e->d_act, e->d_obs, e->d_rew, e->d_done, e->stream);
  if (rc) return rc;
  // small results first so that they ride along with the first observation chunk
  CUDA_TRY(e, cudaMemcpyAsync(pin_r ? reward : e->p_rew, e->d_rew, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(e, cudaMemcpyAsync(pin_d ? done : e->p_done, e->d_done, (size_t)N, cudaMemcpyDeviceToHost, e->stream));
  if ((rc = d2h(e, obs, e->p_obs, e->d_obs, ob, pin_o))) return rc;
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  if (!pin_r) memcpy(reward, e->p_rew, (size_t)N * sizeof(float));
  if (!pin_d) memcpy(done, e->p_done, (size_t)N);
  return ORL_OK;
}

extern "C" int orl_set_max_episode_steps(orl_env* e, int32_t steps) {
  if (!e) return ORL_ERR_INVALID;
  if (steps <= 0) return fail(e, ORL_ERR_INVALID, "max_episode_steps must be positive");
  e->cfg.max_episode_steps = steps;
  e->dev.cfg.max_episode_steps = steps;
  return ORL_OK;
}

extern "C" int orl_set_ref_state_init_prob(orl_env* e, double prob) {
  if (!e) return ORL_ERR_INVALID;
  e->cfg.ref_state_init_prob = prob;
  e->dev.cfg.rsi_prob = prob;
  if (e->d_prepared) {                 // prepared episodes drew their start with the old probability
    CUDA_TRY(e, cudaSetDevice(e->device));
    CUDA_TRY(e, cudaDeviceSynchronize());
    CUDA_TRY(e, cudaMemset(e->d_prepared, 0, (size_t)e->cfg.num_robot * sizeof(int)));
    e->prep_scan = true;
  }
  return ORL_OK;
}

extern "C" int orl_state_layout_get(const orl_env* e, orl_state_layout* o) {
  if (!e || !o) return ORL_ERR_INVALID;
  o->stride = ORL_STRIDE;
  o->pos = S_POS; o->quat = S_QUAT; o->q = S_Q; o->lin = S_LIN; o->ang = S_ANG; o->qd = S_QD;
  o->filt_x = S_FX; o->filt_y = S_FY; o->prev_filtered = S_PREV; o->ring1 = S_RING1; o->ring2 = S_RING2;
  o->imu_hist = S_IMU; o->act_hist = S_ACT; o->ang_hist = S_ANGH;
  o->ref_pose = S_REFPOSE; o->ref_vel = S_REFVEL; o->origin_pos = S_OPOS; o->origin_rot = S_OROT;
  o->time_offset = S_TOFF; o->prev_phase = S_PPHASE;
  o->clip = S_CLIP; o->warmup = S_WARM; o->substep_counter = S_SUBSTEP; o->step_counter = S_STEP;
  o->env_step_counter = S_ENVSTEP; o->episode = S_EPISODE; o->ep_return = S_EPRET; o->obs_quat = S_OBSQ;
  o->motor_strength = S_STRENGTH;
  o->rand_params = S_RAND;
  o->clip_change_time = S_CHANGET; o->clip_changes = S_NCHANGE;
  return ORL_OK;
}

extern "C" int orl_state_ptr(orl_env* e, float** state_dev) {
  if (!e || !state_dev) return ORL_ERR_INVALID;
  *state_dev = e->d_state;
  return ORL_OK;
}

extern "C" int orl_get_state_host(orl_env* e, float* state) {
  if (!e || !state) return ORL_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaDeviceSynchronize());
  CUDA_TRY(e, cudaMemcpy(state, e->d_state, (size_t)e->cfg.num_robot * ORL_STRIDE * sizeof(float), cudaMemcpyDeviceToHost));
  return ORL_OK;
}

extern "C" int orl_set_state_host(orl_env* e, const float* state) {
  if (!e || !state) return ORL_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaDeviceSynchronize());
  CUDA_TRY(e, cudaMemcpy(e->d_state, state, (size_t)e->cfg.num_robot * ORL_STRIDE * sizeof(float), cudaMemcpyHostToDevice));
  e->prep_scan = true;                 // episode counters may have moved: the next launch checks every robot's prepared record
  if (e->d_ahead) CUDA_TRY(e, cudaMemset(e->d_ahead, 0, e->ahead_bytes));
  return ORL_OK;
}

extern "C" int orl_stats(orl_env* e, double out[4], int32_t reset_after) {
  if (!e || !out) return ORL_ERR_INVALID;
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaDeviceSynchronize());
  CUDA_TRY(e, cudaMemcpy(out, e->d_stats, 4 * sizeof(double), cudaMemcpyDeviceToHost));
#ifdef ORL_DEBUG_PREP
  { double h[8]; cudaMemcpy(h, e->d_stats, sizeof(h), cudaMemcpyDeviceToHost); fprintf(stderr, "[orl prep] resets from prepared records %.0f, computed inline %.0f; warp-steps with reference packets %.0f, without %.0f\n", h[4], h[5], h[6], h[7]); }
#endif
#ifdef ORL_PROFILE_RARE
  { unsigned long long h[24]; cudaMemcpyFromSymbol(h, orl_prof, sizeof(h));
    fprintf(stderr, "[orl rare] entries %llu trips %llu cycles %llu | loop cycles %llu warp-steps %llu | max cycles %llu max trips %llu\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    fprintf(stderr, "[orl rare] cycles hist (8k buckets):"); for (int i = 0; i < 8; ++i) fprintf(stderr, " %llu", h[8 + i]);
    fprintf(stderr, "\n[orl rare] trips hist (16 buckets):"); for (int i = 0; i < 4; ++i) fprintf(stderr, " %llu", h[16 + i]);
    fprintf(stderr, "\n[orl rare] backward pass without rare contacts: %.0f cycles (n=%llu); with: %.0f cycles", h[20] ? (double)h[21] / (double)h[20] : 0.0, h[20], h[0] ? (double)h[2] / (double)h[0] : 0.0);
    fprintf(stderr, "\n[orl rare] per warp-step cycles: stage_in %.0f | prologue %.0f | loop %.0f | epilogue (loop end -> stage_out) %.0f | stage_out %.0f", (double)h[12] / h[4], (double)h[13] / h[4], (double)h[3] / h[4], (double)h[15] / h[4], (double)h[14] / h[4]);
    fprintf(stderr, "\n[orl rare] epilogue sections per warp-step: writeback+sensors %.0f | reward %.0f | reference update %.0f | termination %.0f", (double)h[8] / h[4], (double)h[9] / h[4], (double)h[10] / h[4], (double)h[11] / h[4]);
    fprintf(stderr, "\n[orl rare] reset path: %llu warp-steps with a reset, %.0f cycles each | build_obs %.0f cycles per warp-step\n", h[23],
            h[23] ? (double)h[7] / (double)h[23] : 0.0, h[4] ? (double)h[22] / (double)h[4] : 0.0);
    unsigned long long z[24] = {0}; cudaMemcpyToSymbol(orl_prof, z, sizeof(z)); }
#endif
  if (reset_after) CUDA_TRY(e, cudaMemset(e->d_stats, 0, 8 * sizeof(double)));
  return ORL_OK;
}

extern "C" int orl_motion_lookup(orl_env* e, const int32_t* clip_dev, const double* time_dev, int32_t n, int32_t* f0_dev,
                                 int32_t* f1_dev, int32_t* cycle_dev, double* blend_dev, float* pose_dev, float* vel_dev,
                                 void* stream) {
  NEED_MOTIONS(e);
  if (n <= 0) return ORL_OK;
  orl_motion_lookup_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(e->dev.mt, clip_dev, time_dev, n, f0_dev, f1_dev,
                                                                             cycle_dev, blend_dev, pose_dev, vel_dev);
  return check_launch(e, "orl_motion_lookup_kernel");
}

extern "C" int orl_fk(orl_env* e, const float* pose_dev, int32_t n, float* com_dev, void* stream) {
  if (!e) return ORL_ERR_INVALID;
  if (n <= 0) return ORL_OK;
  orl_fk_kernel<<<(n * 4 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(e->d_model, pose_dev, n, com_dev);
  return check_launch(e, "orl_fk_kernel");
}

static int dyn_call(orl_env* e, int mode, const float* pose, const float* vel, const float* in, int n, float* out, void* stream) {
  if (!e) return ORL_ERR_INVALID;
  if (n <= 0) return ORL_OK;
  if (mode == 0) CUDA_TRY(e, cudaMemsetAsync(out, 0, (size_t)n * 324 * sizeof(float), (cudaStream_t)stream));
  orl_dyn_kernel<<<(n * 4 + ORL_DYN_BLOCK - 1) / ORL_DYN_BLOCK, ORL_DYN_BLOCK, 0, (cudaStream_t)stream>>>(e->d_model, e->dev.cfg.phys, mode, pose, vel, in, e->parity_rand, n, out);
  return check_launch(e, "orl_dyn_kernel");
}
extern "C" int orl_get_latency_rings_host(orl_env* e, float* ring_q, float* ring_b) {
  if (!e) return ORL_ERR_INVALID;
  if (!e->dev.ring_q) return fail(e, ORL_ERR_INVALID, "the observation rings only exist when the randomizer is enabled");
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaDeviceSynchronize());
  const size_t n = (size_t)e->cfg.num_robot * ORL_LAT_RING;
  if (ring_q) CUDA_TRY(e, cudaMemcpy(ring_q, e->dev.ring_q, n * 12 * sizeof(float), cudaMemcpyDeviceToHost));
  if (ring_b) CUDA_TRY(e, cudaMemcpy(ring_b, e->dev.ring_b, n * 8 * sizeof(float), cudaMemcpyDeviceToHost));
  return ORL_OK;
}
extern "C" int orl_set_latency_rings_host(orl_env* e, const float* ring_q, const float* ring_b) {
  if (!e) return ORL_ERR_INVALID;
  if (!e->dev.ring_q) return fail(e, ORL_ERR_INVALID, "the observation rings only exist when the randomizer is enabled");
  CUDA_TRY(e, cudaSetDevice(e->device));
  CUDA_TRY(e, cudaDeviceSynchronize());
  const size_t n = (size_t)e->cfg.num_robot * ORL_LAT_RING;
  if (ring_q) CUDA_TRY(e, cudaMemcpy(e->dev.ring_q, ring_q, n * 12 * sizeof(float), cudaMemcpyHostToDevice));
  if (ring_b) CUDA_TRY(e, cudaMemcpy(e->dev.ring_b, ring_b, n * 8 * sizeof(float), cudaMemcpyHostToDevice));
  return ORL_OK;
}
extern "C" int orl_set_dynamics_rand(orl_env* e, const float* rand_dev) {
  if (!e) return ORL_ERR_INVALID;
  e->parity_rand = rand_dev;
  return ORL_OK;
}
extern "C" int orl_mass_matrix(orl_env* e, const float* pose_dev, int32_t n, float* M_dev, void* stream) {
  return dyn_call(e, 0, pose_dev, nullptr, nullptr, n, M_dev, stream);
}
extern "C" int orl_inverse_dynamics(orl_env* e, const float* pose_dev, const float* vel_dev, const float* acc_dev, int32_t n,
                                    float* tau_dev, void* stream) {
  return dyn_call(e, 1, pose_dev, vel_dev, acc_dev, n, tau_dev, stream);
}
extern "C" int orl_forward_dynamics(orl_env* e, const float* pose_dev, const float* vel_dev, const float* tau_dev, int32_t n,
                                    float* acc_dev, void* stream) {
  return dyn_call(e, 2, pose_dev, vel_dev, tau_dev, n, acc_dev, stream);
}
extern "C" int orl_reward_terms(orl_env* e, float* terms_dev, void* stream) {
  NEED_MOTIONS(e);
  const int n = e->cfg.num_robot;
  orl_reward_terms_kernel<<<(n * 4 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(e->dev, terms_dev);
  return check_launch(e, "orl_reward_terms_kernel");
}

// ---- FP32 FFMA micro-benchmark (roofline denominator) -------------------------------------------------
__global__ void __launch_bounds__(256) orl_ffma_kernel(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true; keeps the chain alive
}

extern "C" int orl_fp32_peak(int32_t device, int32_t repeats, double* tflops_out) {
  if (!tflops_out) return ORL_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return ORL_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ORL_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  float* d = nullptr;
  if (cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)) != cudaSuccess) return ORL_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int r = 0; r < repeats + 1; ++r) {
    cudaEventRecord(e0);
    orl_ffma_kernel<<<blocks, threads>>>(d, iters, 0.999f, 1e-4f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 16.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  if (cudaGetLastError() != cudaSuccess) return ORL_ERR_CUDA;
  *tflops_out = best;
  return ORL_OK;
}

// ---- GAE(lambda) over the interleaved rollout buffers (ppo_imitation.py:68-93, 329-338) -------------------
__global__ void orl_gae_kernel(const float* __restrict__ rew, const float* __restrict__ vpred, const float* __restrict__ nextv,
                               const uint8_t* __restrict__ es, int T, int N, float gamma, float lam, int quirk, int standardize,
                               float* __restrict__ adv, float* __restrict__ ret) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float last = 0.f;
  double sum = 0.0, sum2 = 0.0;
  for (int t = T - 1; t >= 0; --t) {
    const long idx = (long)t * N + i;
    const long e = quirk ? idx + 1 + i : idx + N;
    const float nonterminal = 1.0f - ((e < (long)T * N) ? (float)(es[e] != 0) : 0.0f);
    const float delta = rew[idx] + gamma * nextv[idx] - vpred[idx];
    last = delta + gamma * lam * nonterminal * last;
    adv[idx] = last;
    ret[idx] = last + vpred[idx];
    sum += last; sum2 += (double)last * last;
  }
  if (standardize) {
    const double mean = sum / T;
    const double var = fmax(sum2 / T - mean * mean, 0.0);
    // numpy divides by a zero standard deviation here (T == 1 or a constant advantage: NaN into the PPO loss and the Adam
    // moments); this kernel only centres such a column -- a stated divergence, the reference never meets the case
    const float m = (float)mean, inv = var > 0.0 ? (float)(1.0 / sqrt(var)) : 1.0f;
    for (int t = 0; t < T; ++t) { const long idx = (long)t * N + i; adv[idx] = (adv[idx] - m) * inv; }
  }
}

// MotorModel.convert_to_torque (minitaur_motor.py:104-173): one thread per (robot, motor)
__global__ void orl_motor_kernel(int mode, const float* __restrict__ cmd, const float* __restrict__ q, const float* __restrict__ qd,
                                 const float* __restrict__ kp, const float* __restrict__ kd, const float* __restrict__ strength,
                                 const float* __restrict__ limits, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * ORL_NUM_MOTORS) return;
  const int r = i / ORL_NUM_MOTORS, m = i - r * ORL_NUM_MOTORS;
  const float sr = strength ? strength[i] : 1.0f;
  if (mode == ORL_MOTOR_TORQUE) { out[i] = sr * cmd[i]; return; }                       // :137-140 (no clipping)
  float p, d, des_q, des_qd, add;
  if (mode == ORL_MOTOR_POSITION) { p = kp[m]; d = kd[m]; des_q = cmd[i]; des_qd = 0.f; add = 0.f; }   // :147-152
  else {                                                                                 // HYBRID :153-162
    const float* c = cmd + (size_t)r * (5 * ORL_NUM_MOTORS) + 5 * m;
    des_q = c[0]; p = c[1]; des_qd = c[2]; d = c[3]; add = c[4];
  }
  float t = -1.0f * (p * (q[i] - des_q)) - d * (qd[i] - des_qd) + add;                   // :163-165
  t = sr * t;
  if (limits) t = fminf(fmaxf(t, -1.0f * limits[m]), limits[m]);                         // :167-171
  out[i] = t;
}

extern "C" int orl_motor_convert_to_torque(int32_t mode, const float* commands_dev, const float* motor_angle_dev,
                                           const float* motor_velocity_dev, const float* kp_dev, const float* kd_dev,
                                           const float* strength_dev, const float* torque_limits_dev, int32_t n,
                                           float* out_dev, void* stream) {
  if (mode != ORL_MOTOR_POSITION && mode != ORL_MOTOR_TORQUE && mode != ORL_MOTOR_HYBRID) return ORL_ERR_INVALID;   // PWM: ValueError, :133-135
  if (n <= 0 || !commands_dev || !out_dev) return ORL_ERR_INVALID;
  if (mode != ORL_MOTOR_TORQUE && (!motor_angle_dev || !motor_velocity_dev)) return ORL_ERR_INVALID;
  if (mode == ORL_MOTOR_POSITION && (!kp_dev || !kd_dev)) return ORL_ERR_INVALID;
  const int total = n * ORL_NUM_MOTORS;
  orl_motor_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mode, commands_dev, motor_angle_dev, motor_velocity_dev,
                                                                         kp_dev, kd_dev, strength_dev, torque_limits_dev, n, out_dev);
  return cudaGetLastError() == cudaSuccess ? ORL_OK : ORL_ERR_CUDA;
}

extern "C" int orl_gae(const float* rewards_dev, const float* vpred_dev, const float* nextvpred_dev, const uint8_t* episode_starts_dev,
                       int32_t T, int32_t N, float gamma, float lam, int32_t reference_index_quirk, int32_t standardize,
                       float* adv_dev, float* tdlamret_dev, void* stream) {
  if (!rewards_dev || !vpred_dev || !nextvpred_dev || !episode_starts_dev || !adv_dev || !tdlamret_dev || T <= 0 || N <= 0)
    return ORL_ERR_INVALID;
  orl_gae_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards_dev, vpred_dev, nextvpred_dev, episode_starts_dev, T, N,
                                                                  gamma, lam, reference_index_quirk, standardize, adv_dev, tdlamret_dev);
  return cudaGetLastError() == cudaSuccess ? ORL_OK : ORL_ERR_CUDA;
}
