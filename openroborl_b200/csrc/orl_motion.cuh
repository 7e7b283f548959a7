// Reference-motion frame lookup and interpolation on the device.
// Restates R/envs/quadruped_robot/task/motion_data.py:210-253, 417-525, 591-633, 682-718 and the
// origin-offset / warm-up logic of R/envs/quadruped_robot/task/imitation_task.py:920-1018.
// The index arithmetic is done in FP64 with explicitly un-fused operations in the reference's
// operation order, so (f0, f1, cycle) are bit-exact and blend agrees to the last bit; the
// interpolation itself is FP32.
#pragma once
#include "orl_math.cuh"
#include "../../include/orl.h"

#define ORL_FRAME_STRIDE 20   // 19 (pose) / 18 (vel) floats padded to 20

#define ORL_MAX_CLIPS 32
struct MotionTables {
  const float* frames;        // [sumF][20] post-processed poses
  const float* vels;          // [sumF][20] frame velocities
  // Per-clip scalars travel with the kernel parameters (constant bank): every motion function starts from them, and a
  // round trip to L2 for five scalars was the first of three dependent memory latencies of each reference sample.
  double fd[ORL_MAX_CLIPS];        // frame duration
  double dur[ORL_MAX_CLIPS];       // fd * (F - 1), computed on the host in double like motion_data.py:205-207
  double dheading[ORL_MAX_CLIPS];  // cycle delta heading
  int32_t start[ORL_MAX_CLIPS];    // first row of clip c
  int32_t nframes[ORL_MAX_CLIPS];
  int32_t flags[ORL_MAX_CLIPS];    // bit0 wrap, bit1 cycle pos, bit2 cycle rot
  float dpos[ORL_MAX_CLIPS * 3];   // [c][3] cycle delta pos
  const float* cyc_table;     // [c][ORL_CYCLE_TABLE][3] pre-summed rotating cycle offsets (motion_data.py:607-613)
  const float* warm_pose;     // [c][20] warm-up pose (imitation_task.py:985-1009)
  int32_t n_clips;
};

struct BlendIdx { int f0, f1, cycle; double blend; };

// motion_data.py:682-718 + 234-253
ORL_DI BlendIdx motion_blend_idx(const MotionTables& T, int clip, double time) {
  BlendIdx o;
  const double dur = T.dur[clip], fd = T.fd[clip];
  const int nf = T.nframes[clip];
  const bool wrap = (T.flags[clip] & 1) != 0;
  const double ph = __ddiv_rn(time, dur);
  const double fl = floor(ph);
  int cyc = (int)fl;
  if (!wrap) cyc = min(max(cyc, 0), 1);
  o.cycle = cyc;
  if (!wrap && time <= 0.0) { o.f0 = 0; o.f1 = 0; o.blend = 0.0; }
  else if (!wrap && time >= dur) { o.f0 = nf - 1; o.f1 = nf - 1; o.blend = 0.0; }
  else {
    const double phase = wrap ? __dsub_rn(ph, fl) : fmin(fmax(ph, 0.0), 1.0);
    o.f0 = (int)__dmul_rn(phase, (double)(nf - 1));
    o.f1 = min(o.f0 + 1, nf - 1);
    const double norm_time = __dmul_rn(phase, dur);
    const double t0 = __dmul_rn((double)o.f0, fd), t1 = __dmul_rn((double)o.f1, fd);
    o.blend = __ddiv_rn(__dsub_rn(norm_time, t0), __dsub_rn(t1, t0));
  }
  return o;
}

// motion_data.py:210-232
ORL_DI double motion_phase(const MotionTables& T, int clip, double time) {
  const double ph = __ddiv_rn(time, T.dur[clip]);
  return (T.flags[clip] & 1) ? __dsub_rn(ph, floor(ph)) : fmin(fmax(ph, 0.0), 1.0);
}

// motion_data.py:265-276
ORL_DI bool motion_is_over(const MotionTables& T, int clip, double time) {
  return !(T.flags[clip] & 1) && time >= T.dur[clip];
}

// cycle offsets, motion_data.py:591-633
ORL_DI void motion_cycle_offset(const MotionTables& T, int clip, int cycle, float3& off_pos, float& s, float& c,
                                Q4& off_rot) {
  const int fl = T.flags[clip];
  const float3 dp = v3(T.dpos[clip * 3 + 0], T.dpos[clip * 3 + 1], T.dpos[clip * 3 + 2]);
  off_pos = v3(0.f, 0.f, 0.f);
  s = 0.f; c = 1.f; off_rot = q4(0.f, 0.f, 0.f, 1.f);
  if (fl & 4) {
    const double h = __dmul_rn((double)cycle, T.dheading[clip]);
    float sh, ch;
    sincos_f64arg(0.5 * h, sh, ch);
    off_rot = q4(0.f, 0.f, sh, ch);
    s = 2.f * sh * ch; c = ch * ch - sh * sh;
  }
  if (fl & 2) {
    if (!(fl & 4)) off_pos = (float)cycle * dp;
    else if (cycle > 0) {
      if (cycle < ORL_CYCLE_TABLE) {
        const float* t = T.cyc_table + ((size_t)clip * ORL_CYCLE_TABLE + cycle) * 3;
        off_pos = v3(t[0], t[1], t[2]);
      } else {
#pragma unroll 1              // never taken for clips shorter than ORL_CYCLE_TABLE cycles: keep it small
        for (int i = 0; i < cycle; ++i) {
          float si, ci;
          sincos_f64arg((double)i * T.dheading[clip], si, ci);
          off_pos += qrot_z(si, ci, dp);
        }
      }
    }
  }
}

// Root part of MotionData.calc_frame (pos, rot after the cycle offset), motion_data.py:417-449
struct CycleOffset { float3 pos; float s, c; Q4 rot; };
ORL_DI CycleOffset motion_cycle_offset(const MotionTables& T, int clip, int cycle) {
  CycleOffset o;
  motion_cycle_offset(T, clip, cycle, o.pos, o.s, o.c, o.rot);
  return o;
}
ORL_DI void motion_root(const MotionTables& T, int clip, const BlendIdx& b, const CycleOffset& co, float3& pos, Q4& rot) {
  const float* a = T.frames + (size_t)(T.start[clip] + b.f0) * ORL_FRAME_STRIDE;
  const float* e = T.frames + (size_t)(T.start[clip] + b.f1) * ORL_FRAME_STRIDE;
  const float bl = (float)b.blend, ib = 1.0f - bl;
  pos = v3(ib * a[0] + bl * e[0], ib * a[1] + bl * e[1], ib * a[2] + bl * e[2]);
  rot = qstd(qslerp(q4(a[3], a[4], a[5], a[6]), q4(e[3], e[4], e[5], e[6]), bl));
  pos = qrot_z(co.s, co.c, pos) + co.pos;
  rot = qstd(qmul(co.rot, rot));
}
ORL_DI float motion_joint(const MotionTables& T, int clip, const BlendIdx& b, int j) {
  const float a = T.frames[(size_t)(T.start[clip] + b.f0) * ORL_FRAME_STRIDE + 7 + j];
  const float e = T.frames[(size_t)(T.start[clip] + b.f1) * ORL_FRAME_STRIDE + 7 + j];
  const float bl = (float)b.blend;
  return (1.0f - bl) * a + bl * e;
}
// MotionData.calc_frame_vel, motion_data.py:451-476
ORL_DI void motion_root_vel(const MotionTables& T, int clip, const BlendIdx& b, const CycleOffset& co, float3& lin, float3& ang) {
  const float* a = T.vels + (size_t)(T.start[clip] + b.f0) * ORL_FRAME_STRIDE;
  const float* e = T.vels + (size_t)(T.start[clip] + b.f1) * ORL_FRAME_STRIDE;
  const float bl = (float)b.blend, ib = 1.0f - bl;
  lin = v3(ib * a[0] + bl * e[0], ib * a[1] + bl * e[1], ib * a[2] + bl * e[2]);
  ang = v3(ib * a[3] + bl * e[3], ib * a[4] + bl * e[4], ib * a[5] + bl * e[5]);
  lin = qrot_z(co.s, co.c, lin);
  ang = qrot_z(co.s, co.c, ang);
}
ORL_DI float motion_joint_vel(const MotionTables& T, int clip, const BlendIdx& b, int j) {
  const float a = T.vels[(size_t)(T.start[clip] + b.f0) * ORL_FRAME_STRIDE + 6 + j];
  const float e = T.vels[(size_t)(T.start[clip] + b.f1) * ORL_FRAME_STRIDE + 6 + j];
  const float bl = (float)b.blend;
  return (1.0f - bl) * a + bl * e;
}

// Per-episode reference-frame state of the task (imitation_task.py:137-139, 126-128)
struct RefFrame {
  int clip;
  bool warmup;
  double warmup_time;
  Q4 origin_rot;
  float3 origin_pos;
};

ORL_DI bool ref_in_warmup(const RefFrame& F, double time) { return F.warmup && time >= -F.warmup_time && time < 0.0; }

// One evaluation point on the reference clock: the FP64 index math is done once and shared by the
// root / joint / velocity readers.
struct RefSample { bool warm; BlendIdx b; CycleOffset co; };
ORL_DI RefSample ref_sample(const MotionTables& T, const RefFrame& F, double time) {
  RefSample s;
  s.warm = ref_in_warmup(F, time);
  s.b = motion_blend_idx(T, F.clip, time);
  s.co = motion_cycle_offset(T, F.clip, s.b.cycle);
  return s;
}
// root pose of the sample before the per-episode origin is applied
ORL_DI void ref_root_raw(const MotionTables& T, const RefFrame& F, const RefSample& s, float3& pos, Q4& rot) {
  if (s.warm) {
    const float* w = T.warm_pose + (size_t)F.clip * ORL_FRAME_STRIDE;
    pos = v3(w[0], w[1], w[2]); rot = q4(w[3], w[4], w[5], w[6]);
  } else {
    motion_root(T, F.clip, s.b, s.co, pos, rot);
  }
}
ORL_DI void ref_apply_origin(const RefFrame& F, float3& pos, Q4& rot) {
  rot = qmul(F.origin_rot, rot);
  pos = qrot(F.origin_rot, pos) + F.origin_pos;
}
// ImitationTask._calc_ref_pose root part (imitation_task.py:920-953)
ORL_DI void ref_root(const MotionTables& T, const RefFrame& F, const RefSample& s, bool apply_origin, float3& pos, Q4& rot) {
  ref_root_raw(T, F, s, pos, rot);
  if (apply_origin) ref_apply_origin(F, pos, rot);
}
ORL_DI float ref_joint(const MotionTables& T, const RefFrame& F, const RefSample& s, int j) {
  if (s.warm) return T.warm_pose[(size_t)F.clip * ORL_FRAME_STRIDE + 7 + j];
  return motion_joint(T, F.clip, s.b, j);
}
// ImitationTask._calc_ref_vel (imitation_task.py:955-983, 1011-1018)
// root velocity of the sample before the per-episode origin rotation is applied
ORL_DI void ref_root_vel_raw(const MotionTables& T, const RefFrame& F, const RefSample& s, float3& lin, float3& ang) {
  if (s.warm) { lin = v3(0.f, 0.f, 0.f); ang = v3(0.f, 0.f, 0.f); }
  else motion_root_vel(T, F.clip, s.b, s.co, lin, ang);
}
ORL_DI void ref_root_vel(const MotionTables& T, const RefFrame& F, const RefSample& s, float3& lin, float3& ang) {
  ref_root_vel_raw(T, F, s, lin, ang);
  lin = qrot(F.origin_rot, lin);
  ang = qrot(F.origin_rot, ang);
}
ORL_DI float ref_joint_vel(const MotionTables& T, const RefFrame& F, const RefSample& s, int j) {
  if (s.warm) return 0.f;
  return motion_joint_vel(T, F.clip, s.b, j);
}
