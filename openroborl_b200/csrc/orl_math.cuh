// Small fixed-size linear algebra, quaternion helpers and the Philox generator used by every kernel.
// Quaternions are xyzw (the reference's convention, R/envs/utilities/pose3d.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ORL_DI __device__ __forceinline__

#define ORL_PI_F 3.14159265358979323846f
#define ORL_TWO_PI_F 6.28318530717958647692f

struct M3 { float3 r0, r1, r2; };               // rows
struct S3 { float xx, xy, xz, yy, yz, zz; };    // symmetric 3x3
struct Q4 { float x, y, z, w; };

ORL_DI float3 v3(float x, float y, float z) { return make_float3(x, y, z); }
ORL_DI float3 operator+(float3 a, float3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
ORL_DI float3 operator-(float3 a, float3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
ORL_DI float3 operator-(float3 a) { return v3(-a.x, -a.y, -a.z); }
ORL_DI float3 operator*(float s, float3 a) { return v3(s * a.x, s * a.y, s * a.z); }
ORL_DI float3 operator*(float3 a, float s) { return v3(s * a.x, s * a.y, s * a.z); }
ORL_DI float3& operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
ORL_DI float3& operator-=(float3& a, float3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
ORL_DI float dot(float3 a, float3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
ORL_DI float3 cross(float3 a, float3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
ORL_DI float3 clamp3(float3 v, float m) { return make_float3(fminf(fmaxf(v.x, -m), m), fminf(fmaxf(v.y, -m), m), fminf(fmaxf(v.z, -m), m)); }
ORL_DI float3 fma3(float s, float3 a, float3 b) { return v3(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }
// c + a x b with every product fused
ORL_DI float3 cross_add(float3 a, float3 b, float3 c) {
  return v3(fmaf(a.y, b.z, fmaf(-a.z, b.y, c.x)), fmaf(a.z, b.x, fmaf(-a.x, b.z, c.y)), fmaf(a.x, b.y, fmaf(-a.y, b.x, c.z)));
}
ORL_DI float fast_rcp(float x) { return __fdividef(1.0f, x); }

ORL_DI float3 mul(const M3& R, float3 v) { return v3(dot(R.r0, v), dot(R.r1, v), dot(R.r2, v)); }
// a + R v
ORL_DI float3 madd(const M3& R, float3 v, float3 a) {
  return v3(fmaf(R.r0.x, v.x, fmaf(R.r0.y, v.y, fmaf(R.r0.z, v.z, a.x))), fmaf(R.r1.x, v.x, fmaf(R.r1.y, v.y, fmaf(R.r1.z, v.z, a.y))),
            fmaf(R.r2.x, v.x, fmaf(R.r2.y, v.y, fmaf(R.r2.z, v.z, a.z))));
}
ORL_DI float3 col(const M3& R, int c) {
  return c == 0 ? v3(R.r0.x, R.r1.x, R.r2.x) : (c == 1 ? v3(R.r0.y, R.r1.y, R.r2.y) : v3(R.r0.z, R.r1.z, R.r2.z));
}
ORL_DI float3 rowmul(float3 r, const M3& B) { return fma3(r.x, B.r0, fma3(r.y, B.r1, r.z * B.r2)); }
ORL_DI M3 mul(const M3& A, const M3& B) { M3 C; C.r0 = rowmul(A.r0, B); C.r1 = rowmul(A.r1, B); C.r2 = rowmul(A.r2, B); return C; }
// R * Rz(angle): per row (x,y,z) -> (c x + s y, -s x + c y, z)
ORL_DI M3 mul_rz(const M3& A, float s, float c) {
  M3 C;
  C.r0 = v3(fmaf(c, A.r0.x, s * A.r0.y), fmaf(c, A.r0.y, -s * A.r0.x), A.r0.z);
  C.r1 = v3(fmaf(c, A.r1.x, s * A.r1.y), fmaf(c, A.r1.y, -s * A.r1.x), A.r1.z);
  C.r2 = v3(fmaf(c, A.r2.x, s * A.r2.y), fmaf(c, A.r2.y, -s * A.r2.x), A.r2.z);
  return C;
}
ORL_DI float3 mul(const S3& S, float3 v) {
  return v3(fmaf(S.xx, v.x, fmaf(S.xy, v.y, S.xz * v.z)), fmaf(S.xy, v.x, fmaf(S.yy, v.y, S.yz * v.z)),
            fmaf(S.xz, v.x, fmaf(S.yz, v.y, S.zz * v.z)));
}
// R S R^T
ORL_DI S3 rot_sym(const M3& R, const S3& S) {
  float3 t0 = mul(S, R.r0), t1 = mul(S, R.r1), t2 = mul(S, R.r2);   // (S R^T) columns = S * row_i(R)
  S3 o;
  o.xx = dot(R.r0, t0); o.xy = dot(R.r0, t1); o.xz = dot(R.r0, t2);
  o.yy = dot(R.r1, t1); o.yz = dot(R.r1, t2); o.zz = dot(R.r2, t2);
  return o;
}
ORL_DI S3 operator+(const S3& a, const S3& b) { S3 o = {a.xx + b.xx, a.xy + b.xy, a.xz + b.xz, a.yy + b.yy, a.yz + b.yz, a.zz + b.zz}; return o; }

// ---- quaternions ---------------------------------------------------------------------------
ORL_DI Q4 q4(float x, float y, float z, float w) { Q4 q = {x, y, z, w}; return q; }
// Hamilton product a (x) b  (pybullet_utils.transformations.quaternion_multiply(a, b))
ORL_DI Q4 qmul(Q4 a, Q4 b) {
  return q4(a.x * b.w + a.y * b.z - a.z * b.y + a.w * b.x, -a.x * b.z + a.y * b.w + a.z * b.x + a.w * b.y,
            a.x * b.y - a.y * b.x + a.z * b.w + a.w * b.z, -a.x * b.x - a.y * b.y - a.z * b.z + a.w * b.w);
}
ORL_DI Q4 qconj(Q4 q) { return q4(-q.x, -q.y, -q.z, q.w); }
ORL_DI Q4 qneg(Q4 q) { return q4(-q.x, -q.y, -q.z, -q.w); }
ORL_DI Q4 qstd(Q4 q) { return q.w < 0.f ? qneg(q) : q; }                  // pose3d.standardize_quaternion
ORL_DI float qdot(Q4 a, Q4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
ORL_DI Q4 qnormalize(Q4 q) { float s = rsqrtf(qdot(q, q)); return q4(q.x * s, q.y * s, q.z * s, q.w * s); }
ORL_DI Q4 qaxis_z(float angle) { float s, c; sincosf(0.5f * angle, &s, &c); return q4(0.f, 0.f, s, c); }  // about_axis(angle,[0,0,1])
ORL_DI M3 qmat(Q4 q) {
  M3 R;
  R.r0 = v3(1.f - 2.f * (q.y * q.y + q.z * q.z), 2.f * (q.x * q.y - q.z * q.w), 2.f * (q.x * q.z + q.y * q.w));
  R.r1 = v3(2.f * (q.x * q.y + q.z * q.w), 1.f - 2.f * (q.x * q.x + q.z * q.z), 2.f * (q.y * q.z - q.x * q.w));
  R.r2 = v3(2.f * (q.x * q.z - q.y * q.w), 2.f * (q.y * q.z + q.x * q.w), 1.f - 2.f * (q.x * q.x + q.y * q.y));
  return R;
}
// pose3d.QuaternionRotatePoint for a unit quaternion: v + 2 w (u x v) + 2 u x (u x v)
ORL_DI float3 qrot(Q4 q, float3 v) {
  float3 u = v3(q.x, q.y, q.z);
  float3 t = 2.f * cross(u, v);
  return v + q.w * t + cross(u, t);
}
// rotation about z by a unit (0,0,s,c) quaternion
ORL_DI float3 qrot_z(float s, float c, float3 v) {   // s = sin(angle), c = cos(angle) (full angle)
  return v3(c * v.x - s * v.y, s * v.x + c * v.y, v.z);
}
// pose3d.calc_heading: atan2 of the rotated x axis
ORL_DI float qheading(Q4 q) {
  return atan2f(2.f * (q.x * q.y + q.z * q.w), 1.f - 2.f * (q.y * q.y + q.z * q.z));
}
// sin / cos of MINUS the heading without trigonometry: the heading is atan2(y, x) of the rotated x axis, so
// cos = x / |(x, y)| and sin = -y / |(x, y)| (heading 0 for a vertical axis, like atan2(0, 0))
ORL_DI void qheading_neg_sincos(Q4 q, float& s, float& c) {
  const float y = 2.f * (q.x * q.y + q.z * q.w), x = 1.f - 2.f * (q.y * q.y + q.z * q.z);
  const float n2 = fmaf(x, x, y * y);
  const float inv = n2 > 0.f ? rsqrtf(n2) : 0.f;
  c = n2 > 0.f ? x * inv : 1.f;
  s = -y * inv;
}
// half-angle of a rotation given by (s, c) = (sin a, cos a), a in (-pi, pi]: (sin a/2, cos a/2) without trigonometry;
// each branch takes the square root of the well-conditioned one of (1 + c) / 2, (1 - c) / 2
ORL_DI void half_angle(float s, float c, float& sh, float& ch) {
  if (c >= 0.f) { ch = sqrtf(0.5f * (1.f + c)); sh = 0.5f * s / ch; }
  else { const float a = sqrtf(0.5f * (1.f - c)); sh = s >= 0.f ? a : -a; ch = 0.5f * fabsf(s) / a; }
}
// pose3d.QuaternionToAxisAngle angle + normalize_rotation_angle
ORL_DI float qangle_normalized(Q4 q) {
  float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
  float ang = 2.f * atan2f(n, q.w);
  // n >= 0, so ang is in [0, 2 pi]: normalize_rotation_angle's fmod (motion_util.py) is the identity there, only its
  // "> pi" branch can fire (an inline fmodf was 50 instructions of once-per-step code at each of the two call sites)
  if (ang > ORL_PI_F) ang -= ORL_TWO_PI_F;
  return ang;
}
// pose3d.MapToMinusPiToPi
ORL_DI float wrap_pi(float a) {
  float m = fmodf(a, ORL_TWO_PI_F);
  if (m >= ORL_PI_F) m -= ORL_TWO_PI_F; else if (m < -ORL_PI_F) m += ORL_TWO_PI_F;
  return m;
}
// identical to wrap_pi, without the fmodf loop in the (usual) case |a| < pi
// Compact variant for the substep loop (code size): identical to wrap_pi for |a| < pi; beyond, the nearest-
// multiple reduction differs from it only at a = pi + 2 pi k exactly (maps to +pi instead of -pi).
ORL_DI float wrap_pi_fast(float a) { return fabsf(a) < ORL_PI_F ? a : fmaf(-ORL_TWO_PI_F, rintf(a * (1.0f / ORL_TWO_PI_F)), a); }
// pybullet.getEulerFromQuaternion (roll, pitch only) and yaw
ORL_DI void bullet_euler(Q4 q, float& roll, float& pitch, float& yaw) {
  float sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z, squ = q.w * q.w;
  float sarg = -2.f * (q.x * q.z - q.w * q.y);
  if (sarg <= -0.99999f) { pitch = -0.5f * ORL_PI_F; roll = 0.f; yaw = 2.f * atan2f(q.x, -q.y); }
  else if (sarg >= 0.99999f) { pitch = 0.5f * ORL_PI_F; roll = 0.f; yaw = 2.f * atan2f(-q.x, q.y); }
  else {
    pitch = asinf(sarg);
    roll = atan2f(2.f * (q.y * q.z + q.w * q.x), squ - sqx - sqy + sqz);
    yaw = atan2f(2.f * (q.x * q.y + q.w * q.z), squ + sqx - sqy - sqz);
  }
}
// pybullet.getQuaternionFromEuler (setEulerZYX + normalize)
ORL_DI Q4 bullet_quat_from_euler(float roll, float pitch, float yaw) {
  float sr, cr, sp, cp, sy, cy;       // half angles in [-pi/2, pi/2]: SFU sincos, abs. error < 4e-7
  __sincosf(0.5f * roll, &sr, &cr); __sincosf(0.5f * pitch, &sp, &cp); __sincosf(0.5f * yaw, &sy, &cy);
  return qnormalize(q4(sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                       cr * cp * cy + sr * sp * sy));
}
// transformations.quaternion_slerp (shortest path) for unit inputs.  Near-parallel inputs use the
// normalised lerp (difference from slerp < theta^3/20, theta < 0.03 rad); see DESIGN.md.
ORL_DI Q4 qslerp(Q4 a, Q4 b, float f) {
  if (f == 0.f) return a;
  if (f == 1.f) return b;
  float d = qdot(a, b);
  if (d < 0.f) { d = -d; b = qneg(b); }
  float wa, wb;
  if (d > 0.9995f) {
    wa = 1.f - f; wb = f;
    Q4 r = q4(wa * a.x + wb * b.x, wa * a.y + wb * b.y, wa * a.z + wb * b.z, wa * a.w + wb * b.w);
    return qnormalize(r);
  }
  float ang = acosf(d);
  float isin = 1.f / sinf(ang);
  wa = sinf((1.f - f) * ang) * isin; wb = sinf(f * ang) * isin;
  return q4(wa * a.x + wb * b.x, wa * a.y + wb * b.y, wa * a.z + wb * b.z, wa * a.w + wb * b.w);
}
// sin / cos of a DOUBLE angle of any magnitude (cycle headings: cycle count x heading per cycle): reduction to
// [-pi/4, pi/4] in FP64 (the reference computes these in double; rounding the angle to float first cost up to 1.5e-5 at
// |x| = 300), then degree-9 / degree-8 polynomials in float: absolute error 1e-7, 20 instructions against the ~190 of
// sincosf with its large-argument path.
ORL_DI void sincos_f64arg(double x, float& s, float& c) {
  const int k = __double2int_rn(x * 0.63661977236758134308);
  const double kd = (double)k;
  const float r = (float)fma(-kd, 6.123233995736766e-17, fma(-kd, 1.5707963267948966, x));
  const float r2 = r * r;
  const float sr = fmaf(r * r2, fmaf(r2, fmaf(r2, fmaf(r2, 1.f / 362880.f, -1.f / 5040.f), 1.f / 120.f), -1.f / 6.f), r);
  const float cr = fmaf(r2, fmaf(r2, fmaf(r2, fmaf(r2, 1.f / 40320.f, -1.f / 720.f), 1.f / 24.f), -0.5f), 1.f);
  const bool swap = k & 1;
  const float ss = swap ? cr : sr, cc = swap ? sr : cr;
  s = (k & 2) ? -ss : ss;
  c = ((k + 1) & 2) ? -cc : cc;
}

// ---- Philox4x32-10 (same contract as oracle/philox.py) --------------------------------------
ORL_DI void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
ORL_DI double u01_53(uint32_t hi, uint32_t lo) {
  unsigned long long v = ((unsigned long long)hi << 21) | ((unsigned long long)lo >> 11);
  return (double)v * (1.0 / 9007199254740992.0);
}

// ---- quad (4-lane) reductions: lanes 4r..4r+3 of a warp belong to robot r ---------------------
ORL_DI float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
ORL_DI int quad_or(int v) {
  v |= __shfl_xor_sync(0xffffffffu, v, 1);
  v |= __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
