// Per-leg rigid-body dynamics: one lane per leg, four lanes per robot.
//
// Formulation (DESIGN.md "dynamics kernel"): all spatial quantities of a robot are expressed in the
// frame A = world-aligned axes with origin at the base COM.  Each lane runs forward kinematics, the
// composite-rigid-body algorithm (CRBA) and the recursive Newton-Euler algorithm (RNEA) for its own
// 3-joint leg chain; the four legs only couple through the floating base, so the 18x18 system
//     [ Hbb  B1 B2 B3 B4 ] [a_b ]   [ -p_b ]
//     [ B_l^T    A_ll    ] [qdd_l] = [ tau_l - C_l ]
// is solved by a per-lane 3x3 LDL^T, a 4-lane shuffle reduction of the 6x6 Schur complement and a
// 6x6 LDL^T.  Ground contact is a semi-implicit penalty model: the force at the start-of-substep
// velocity enters RNEA as an external force and its velocity derivative (frozen diagonal D) adds
// dt * J^T D J to the same block pattern.
//
// v2 (profiles/r01_v1_step_kernel.md): model constants come as float4 from shared memory, collision
// points are culled per body with a bounding radius and a warp vote, trigonometry / reciprocals inside
// the substep loop use the SFU, and every sum of products is written as an FMA chain.
#pragma once
#include "orl_math.cuh"
#include "orl_devmodel.h"

struct PhysParams {
  float dt, g;              // g = -gravity_z (>0)
  float kn, bn, inv_d0, bt;
  float klim, dlim, inv_e0;
};

// rigid-body inertia at the frame-A origin: mass, first moment h = m c, rotational inertia about the origin
struct RBI { float m; float3 h; S3 I; };
ORL_DI RBI operator+(const RBI& a, const RBI& b) { RBI o; o.m = a.m + b.m; o.h = a.h + b.h; o.I = a.I + b.I; return o; }
struct SV { float3 a, l; };   // spatial vector: angular, linear
ORL_DI SV operator+(const SV& x, const SV& y) { SV o; o.a = x.a + y.a; o.l = x.l + y.l; return o; }
ORL_DI SV operator-(const SV& x, const SV& y) { SV o; o.a = x.a - y.a; o.l = x.l - y.l; return o; }
ORL_DI float sdot(const SV& x, const SV& y) {
  return fmaf(x.a.x, y.a.x, fmaf(x.a.y, y.a.y, fmaf(x.a.z, y.a.z, fmaf(x.l.x, y.l.x, fmaf(x.l.y, y.l.y, x.l.z * y.l.z)))));
}
// I * [w; v] = [I w + h x v ; m v - h x w]
ORL_DI SV apply(const RBI& B, const SV& v) {
  SV o;
  o.a = cross_add(B.h, v.l, mul(B.I, v.a));
  o.l = cross_add(v.a, B.h, B.m * v.l);
  return o;
}
// acc + v x m   (motion cross product)
ORL_DI SV crm_add(const SV& v, const SV& m, const SV& acc) {
  SV o;
  o.a = cross_add(v.a, m.a, acc.a);
  o.l = cross_add(v.l, m.a, cross_add(v.a, m.l, acc.l));
  return o;
}
// acc + v x* f  (force cross product)
ORL_DI SV crf_add(const SV& v, const SV& f, const SV& acc) {
  SV o;
  o.a = cross_add(v.l, f.l, cross_add(v.a, f.a, acc.a));
  o.l = cross_add(v.a, f.l, acc.l);
  return o;
}

ORL_DI RBI body_inertia(float m, float3 c /* COM rel. origin, frame A */, const S3& Ic /* about COM, frame A */) {
  RBI o; o.m = m; o.h = m * c;
  const float mcc = m * dot(c, c);
  const float3 mc = o.h;
  o.I.xx = fmaf(-mc.x, c.x, Ic.xx + mcc); o.I.yy = fmaf(-mc.y, c.y, Ic.yy + mcc); o.I.zz = fmaf(-mc.z, c.z, Ic.zz + mcc);
  o.I.xy = fmaf(-mc.x, c.y, Ic.xy); o.I.xz = fmaf(-mc.x, c.z, Ic.xz); o.I.yz = fmaf(-mc.y, c.z, Ic.yz);
  return o;
}

// lower-triangular index of a symmetric 6x6 stored row-major, i >= j
#define TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

// expand a rigid-body inertia into the 21 lower-triangular entries of its 6x6 ([ang; lin] ordering)
ORL_DI void rbi_to_tri(const RBI& B, float A[21]) {
  A[TRI(0, 0)] = B.I.xx; A[TRI(1, 0)] = B.I.xy; A[TRI(1, 1)] = B.I.yy;
  A[TRI(2, 0)] = B.I.xz; A[TRI(2, 1)] = B.I.yz; A[TRI(2, 2)] = B.I.zz;
  A[TRI(3, 0)] = 0.f;    A[TRI(3, 1)] = B.h.z;  A[TRI(3, 2)] = -B.h.y;
  A[TRI(4, 0)] = -B.h.z; A[TRI(4, 1)] = 0.f;    A[TRI(4, 2)] = B.h.x;
  A[TRI(5, 0)] = B.h.y;  A[TRI(5, 1)] = -B.h.x; A[TRI(5, 2)] = 0.f;
  A[TRI(3, 3)] = B.m; A[TRI(4, 3)] = 0.f; A[TRI(4, 4)] = B.m; A[TRI(5, 3)] = 0.f; A[TRI(5, 4)] = 0.f; A[TRI(5, 5)] = B.m;
}

// LDL^T solve of a symmetric positive definite 6x6 (lower triangle in A, destroyed); x overwrites b
ORL_DI void ldl6_solve(float A[21], float b[6]) {
  float D[6], Dinv[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float d = A[TRI(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) d = fmaf(-A[TRI(j, k)] * D[k], A[TRI(j, k)], d);
    D[j] = d;
    Dinv[j] = fast_rcp(d);
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      float s = A[TRI(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fmaf(-A[TRI(i, k)] * D[k], A[TRI(j, k)], s);
      A[TRI(i, j)] = s * Dinv[j];
    }
  }
#pragma unroll
  for (int i = 1; i < 6; ++i)
#pragma unroll
    for (int k = 0; k < i; ++k) b[i] = fmaf(-A[TRI(i, k)], b[k], b[i]);
#pragma unroll
  for (int i = 0; i < 6; ++i) b[i] *= Dinv[i];
#pragma unroll
  for (int i = 4; i >= 0; --i)
#pragma unroll
    for (int k = i + 1; k < 6; ++k) b[i] = fmaf(-A[TRI(k, i)], b[k], b[i]);
}

// LDL^T of a symmetric positive definite 3x3: A = [a00 a10 a11 a20 a21 a22] (lower, row-major)
struct LDL3 { float l10, l20, l21, i0, i1, i2; };
ORL_DI LDL3 ldl3(const float A[6]) {
  LDL3 f;
  f.i0 = fast_rcp(A[0]);
  f.l10 = A[1] * f.i0; f.l20 = A[3] * f.i0;
  const float d1 = fmaf(-f.l10, A[1], A[2]);
  f.i1 = fast_rcp(d1);
  const float t = fmaf(-f.l20, A[1], A[4]);
  f.l21 = t * f.i1;
  const float d2 = fmaf(-f.l21, t, fmaf(-f.l20, A[3], A[5]));
  f.i2 = fast_rcp(d2);
  return f;
}
ORL_DI void ldl3_solve(const LDL3& f, float& x0, float& x1, float& x2) {
  x1 = fmaf(-f.l10, x0, x1); x2 = fmaf(-f.l21, x1, fmaf(-f.l20, x0, x2));
  x0 *= f.i0; x1 *= f.i1; x2 *= f.i2;
  x1 = fmaf(-f.l21, x2, x1); x0 = fmaf(-f.l20, x2, fmaf(-f.l10, x1, x0));
}

ORL_DI float3 f4xyz(float4 v) { return v3(v.x, v.y, v.z); }

// forward kinematics of one leg in frame A.  FAST: SFU sincos (abs. error < 4e-7 on [-pi, pi]).
struct LegFK { M3 R[3]; float3 p[3]; };
template <bool FAST>
ORL_DI void leg_fk(const DevLeg& L, const M3& Rb, const float q[3], LegFK& K) {
  M3 Rp = Rb; float3 pp = v3(0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float4 a0 = L.j[k].r0, a1 = L.j[k].r1, a2 = L.j[k].r2;
    M3 F; F.r0 = f4xyz(a0); F.r1 = f4xyz(a1); F.r2 = f4xyz(a2);
    float s, c;
    if (FAST) __sincosf(q[k], &s, &c); else sincosf(q[k], &s, &c);
    K.p[k] = madd(Rp, v3(a0.w, a1.w, a2.w), pp);
    K.R[k] = mul_rz(mul(Rp, F), s, c);
    Rp = K.R[k]; pp = K.p[k];
  }
}
ORL_DI M3 pickR(const LegFK& K, int kb) { return kb == 0 ? K.R[0] : (kb == 1 ? K.R[1] : K.R[2]); }
ORL_DI float3 pickp(const LegFK& K, int kb) { return kb == 0 ? K.p[0] : (kb == 1 ? K.p[1] : K.p[2]); }

// What the caller can ask leg_dynamics to export (parity entry points)
struct DynExport {
  float Hll[6];        // leg block of the mass matrix, lower triangle
  SV Fcol[3];          // base-leg coupling columns
  float Hbb[21];       // base block (full, after the 4-lane reduction)
  float C[3];          // joint bias (incl. gravity, excl. contact)
  float pb[6];         // base bias
};

// Accumulators of the contact model for one lane
struct ContactAcc {
  float Wbb[21], Wll[6];
  SV Wbl[3];
  SV fb_ext;
  bool any;
};

// One active collision point: force at the old velocity -> external force on its body, frozen derivative -> W.
// KB is the body the point is fixed to (-1 = base); r is the contact location relative to the frame-A origin.
template <int KB>
ORL_DI void contact_point(const PhysParams& P, float d, float mu, float3 r, const SV& vk, float3 e_cl, const SV S[3],
                          const float3 p[3], SV f[3], ContactAcc& W) {
  const float3 V = cross_add(vk.a, r, vk.l);
  const float sc = fminf(1.0f, d * P.inv_d0);
  const float fn = fmaxf(0.0f, fmaf(-P.bn * sc, V.z, P.kn * d));
  const float vt2 = fmaf(V.x, V.x, V.y * V.y);
  const float ct = fminf(P.bt, mu * fn * rsqrtf(fmaxf(vt2, 1e-18f)));
  const float cn = fn > 0.f ? fmaf(P.bn, sc, P.kn * P.dt) : 0.f;
  const float3 Dd = v3(P.dt * ct, P.dt * ct, P.dt * cn);
  // force at the old velocity, corrected for the classical/spatial acceleration difference
  const float3 F = v3(fmaf(-Dd.x, e_cl.x, -ct * V.x), fmaf(-Dd.y, e_cl.y, -ct * V.y), fmaf(-Dd.z, e_cl.z, fn));
  SV fe; fe.a = cross(r, F); fe.l = F;
  if (KB < 0) W.fb_ext = W.fb_ext + fe; else f[KB < 0 ? 0 : KB] = f[KB < 0 ? 0 : KB] - fe;
  // Jacobian columns of the contact-point velocity: base w (e_a x r), base v (e_a), joints 0..KB
  float3 J[9];
  J[0] = v3(0.f, -r.z, r.y); J[1] = v3(r.z, 0.f, -r.x); J[2] = v3(-r.y, r.x, 0.f);
  J[3] = v3(1.f, 0.f, 0.f); J[4] = v3(0.f, 1.f, 0.f); J[5] = v3(0.f, 0.f, 1.f);
#pragma unroll
  for (int j = 0; j < 3; ++j) J[6 + j] = (j <= KB) ? cross(S[j].a, r - p[j]) : v3(0.f, 0.f, 0.f);
  float3 DJ[9];
#pragma unroll
  for (int a = 0; a < 9; ++a) DJ[a] = v3(Dd.x * J[a].x, Dd.y * J[a].y, Dd.z * J[a].z);
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) W.Wbb[TRI(a, b)] += dot(DJ[a], J[b]);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j <= KB) {
      W.Wbl[j].a += v3(dot(DJ[0], J[6 + j]), dot(DJ[1], J[6 + j]), dot(DJ[2], J[6 + j]));
      W.Wbl[j].l += v3(dot(DJ[3], J[6 + j]), dot(DJ[4], J[6 + j]), dot(DJ[5], J[6 + j]));
    }
  }
  if (KB >= 0) W.Wll[0] += dot(DJ[6], J[6]);
  if (KB >= 1) { W.Wll[1] += dot(DJ[7], J[6]); W.Wll[2] += dot(DJ[7], J[7]); }
  if (KB >= 2) { W.Wll[3] += dot(DJ[8], J[6]); W.Wll[4] += dot(DJ[8], J[7]); W.Wll[5] += dot(DJ[8], J[8]); }
  W.any = true;
}

// All collision points of body KB of this lane: bounding-radius cull, warp vote, then per-point test.
template <int KB>
ORL_DI void contact_body(const DevLeg& L, const PhysParams& P, int first, int count, float brad, float base_z, const M3& Rk,
                         float3 pk, const SV& vk, float3 e_cl, const SV S[3], const float3 p[3], SV f[3], ContactAcc& W) {
  const float oz = base_z + pk.z;
  if (!__any_sync(0xffffffffu, oz < brad)) return;       // warp-uniform skip: nobody's body KB is near the ground
  if (oz >= brad) return;
  // phase 1: penetration test of every point (independent iterations: the LDS.128 loads pipeline)
  unsigned active = 0u;
#pragma unroll 4
  for (int i = 0; i < count; ++i) {
    const float4 a = L.cp[first + i].a;
    const float z = oz + dot(Rk.r2, f4xyz(a)) - a.w;
    active |= (z < 0.f ? 1u : 0u) << i;
  }
  // phase 2: the (few) penetrating points
  while (active) {
    const int i = __ffs(active) - 1;
    active &= active - 1u;
    const float4 a = L.cp[first + i].a;
    const float z = oz + dot(Rk.r2, f4xyz(a)) - a.w;
    float3 r = madd(Rk, f4xyz(a), pk); r.z -= a.w;
    contact_point<KB>(P, -z, L.cp[first + i].b.x, r, vk, e_cl, S, p, f, W);
  }
}

// One dynamics evaluation for this lane's leg.  Returns the spatial base acceleration (a_b: angular
// alpha, linear a_lin so that d/dt v_com = a_lin + w x v) and this leg's joint accelerations.
// tau: generalised joint forces in URDF joint space (motor + limit + damping torques).
template <bool CONTACT, bool EXPORT, bool FAST>
ORL_DI void leg_dynamics(const DevLeg& L, const DevModel& Mo, const PhysParams& P, const M3& Rb, float base_z, float3 wb,
                         float3 vb, const float q[3], const float qd[3], const float tau[3], float3& alpha_b, float3& alin_b,
                         float qdd[3], DynExport* ex) {
  LegFK K;
  leg_fk<FAST>(L, Rb, q, K);
  // ---- velocities, velocity-product accelerations, per-body inertia and bias force -------------
  SV S[3], v[3], f[3];
  RBI I[3];
  SV vp; vp.a = wb; vp.l = vb;
  SV ap; ap.a = v3(0.f, 0.f, 0.f); ap.l = v3(0.f, 0.f, P.g);    // a0 = -a_gravity
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    S[k].a = col(K.R[k], 2);
    S[k].l = cross(K.p[k], S[k].a);
    SV sq; sq.a = qd[k] * S[k].a; sq.l = qd[k] * S[k].l;
    v[k].a = vp.a + sq.a; v[k].l = vp.l + sq.l;
    const SV avp = crm_add(vp, sq, ap);
    const float4 i0 = L.j[k].i0, i1 = L.j[k].i1, c1 = L.j[k].c1;
    const float3 c = madd(K.R[k], v3(i1.w, c1.x, c1.y), K.p[k]);
    const S3 Ic = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y};
    I[k] = body_inertia(i1.z, c, rot_sym(K.R[k], Ic));
    f[k] = crf_add(v[k], apply(I[k], v[k]), apply(I[k], avp));
    vp = v[k]; ap = avp;
  }
  // ---- contact: explicit force -> external force; implicit part -> W blocks --------------------
  ContactAcc W;
  W.any = false;
  W.fb_ext.a = v3(0.f, 0.f, 0.f); W.fb_ext.l = v3(0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 21; ++i) W.Wbb[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) W.Wll[i] = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) { W.Wbl[k].a = v3(0.f, 0.f, 0.f); W.Wbl[k].l = v3(0.f, 0.f, 0.f); }
  if (CONTACT) {
    const float3 e_cl = cross(wb, vb);     // classical - spatial linear base acceleration
    const int4 cnt = L.cnt;
    SV vbase; vbase.a = wb; vbase.l = vb;
    contact_body<2>(L, P, cnt.x + cnt.y, cnt.z, L.j[2].m1.w, base_z, K.R[2], K.p[2], v[2], e_cl, S, K.p, f, W);
    contact_body<1>(L, P, cnt.x, cnt.y, L.j[1].m1.w, base_z, K.R[1], K.p[1], v[1], e_cl, S, K.p, f, W);
    contact_body<0>(L, P, 0, cnt.x, L.j[0].m1.w, base_z, K.R[0], K.p[0], v[0], e_cl, S, K.p, f, W);
    contact_body<-1>(L, P, cnt.x + cnt.y + cnt.z, cnt.w, __int_as_float(L.meta.y), base_z, Rb, v3(0.f, 0.f, 0.f), vbase, e_cl,
                     S, K.p, f, W);
  }
  // ---- composite inertias / forces, CRBA, RNEA bias ---------------------------------------------
  const RBI Ic2 = I[2], Ic1 = I[1] + Ic2, Ic0 = I[0] + Ic1;
  const SV fc2 = f[2], fc1 = f[1] + fc2, fc0 = f[0] + fc1;
  SV Fk[3];
  Fk[0] = apply(Ic0, S[0]); Fk[1] = apply(Ic1, S[1]); Fk[2] = apply(Ic2, S[2]);
  float A[6];
  A[0] = sdot(S[0], Fk[0]);
  A[1] = sdot(S[0], Fk[1]); A[2] = sdot(S[1], Fk[1]);
  A[3] = sdot(S[0], Fk[2]); A[4] = sdot(S[1], Fk[2]); A[5] = sdot(S[2], Fk[2]);
  const float C0 = sdot(S[0], fc0), C1 = sdot(S[1], fc1), C2 = sdot(S[2], fc2);
  if (EXPORT) {
#pragma unroll
    for (int i = 0; i < 6; ++i) ex->Hll[i] = A[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) ex->Fcol[k] = Fk[k];
    ex->C[0] = C0; ex->C[1] = C1; ex->C[2] = C2;
  }
  // ---- lane-local Schur complement ---------------------------------------------------------------
  SV B[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) B[k] = Fk[k];
  if (CONTACT && W.any) {
#pragma unroll
    for (int i = 0; i < 6; ++i) A[i] += W.Wll[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) B[k] = B[k] + W.Wbl[k];
  }
  const LDL3 fac = ldl3(A);
  float r0 = tau[0] - C0, r1 = tau[1] - C1, r2 = tau[2] - C2;
  ldl3_solve(fac, r0, r1, r2);                 // y_r = A^-1 r
  // Y = A^-1 B^T : 6 right-hand sides, component c of B[k]
  float Bm[6][3], Y[3][6];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    Bm[0][k] = B[k].a.x; Bm[1][k] = B[k].a.y; Bm[2][k] = B[k].a.z;
    Bm[3][k] = B[k].l.x; Bm[4][k] = B[k].l.y; Bm[5][k] = B[k].l.z;
  }
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    float y0 = Bm[c][0], y1 = Bm[c][1], y2 = Bm[c][2];
    ldl3_solve(fac, y0, y1, y2);
    Y[0][c] = y0; Y[1][c] = y1; Y[2][c] = y2;
  }
  float red[27];
  rbi_to_tri(Ic0, red);
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b)
      red[TRI(a, b)] = fmaf(-Bm[a][0], Y[0][b], fmaf(-Bm[a][1], Y[1][b], fmaf(-Bm[a][2], Y[2][b], red[TRI(a, b)])));
  if (CONTACT && W.any) {
#pragma unroll
    for (int i = 0; i < 21; ++i) red[i] += W.Wbb[i];
  }
  {
    const float pbv[6] = {fc0.a.x - W.fb_ext.a.x, fc0.a.y - W.fb_ext.a.y, fc0.a.z - W.fb_ext.a.z,
                          fc0.l.x - W.fb_ext.l.x, fc0.l.y - W.fb_ext.l.y, fc0.l.z - W.fb_ext.l.z};
#pragma unroll
    for (int a = 0; a < 6; ++a) red[21 + a] = fmaf(-Bm[a][0], r0, fmaf(-Bm[a][1], r1, fmaf(-Bm[a][2], r2, -pbv[a])));
  }
  if (EXPORT) {
    // un-modified base block and base bias need their own reduction (no W, no Schur term)
    float t[27];
    rbi_to_tri(Ic0, t);
    t[21] = fc0.a.x; t[22] = fc0.a.y; t[23] = fc0.a.z; t[24] = fc0.l.x; t[25] = fc0.l.y; t[26] = fc0.l.z;
#pragma unroll
    for (int i = 0; i < 27; ++i) t[i] = quad_sum(t[i]);
#pragma unroll
    for (int i = 0; i < 21; ++i) ex->Hbb[i] = t[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) ex->pb[i] = t[21 + i];
  }
#pragma unroll
  for (int i = 0; i < 27; ++i) red[i] = quad_sum(red[i]);
  // ---- base body itself (computed redundantly by the 4 lanes) --------------------------------------
  {
    const float4 b0 = Mo.b0, b1 = Mo.b1;
    const S3 Ib = {b0.y, b0.z, b0.w, b1.x, b1.y, b1.z};
    const S3 IbA = rot_sym(Rb, Ib);
    // COM at the origin: I = diag(IbA, m); bias force = I a0 + v x* I v = [w x (IbA w) ; m (g e_z + w x v)]
    const float3 Iw = mul(IbA, wb);
    const float3 fa = cross(wb, Iw);
    const float3 fl = b0.x * cross_add(wb, vb, v3(0.f, 0.f, P.g));
    red[TRI(0, 0)] += IbA.xx; red[TRI(1, 0)] += IbA.xy; red[TRI(1, 1)] += IbA.yy;
    red[TRI(2, 0)] += IbA.xz; red[TRI(2, 1)] += IbA.yz; red[TRI(2, 2)] += IbA.zz;
    red[TRI(3, 3)] += b0.x; red[TRI(4, 4)] += b0.x; red[TRI(5, 5)] += b0.x;
    red[21] -= fa.x; red[22] -= fa.y; red[23] -= fa.z;
    red[24] -= fl.x; red[25] -= fl.y; red[26] -= fl.z;
    if (EXPORT) {
      ex->Hbb[TRI(0, 0)] += IbA.xx; ex->Hbb[TRI(1, 0)] += IbA.xy; ex->Hbb[TRI(1, 1)] += IbA.yy;
      ex->Hbb[TRI(2, 0)] += IbA.xz; ex->Hbb[TRI(2, 1)] += IbA.yz; ex->Hbb[TRI(2, 2)] += IbA.zz;
      ex->Hbb[TRI(3, 3)] += b0.x; ex->Hbb[TRI(4, 4)] += b0.x; ex->Hbb[TRI(5, 5)] += b0.x;
      ex->pb[0] += fa.x; ex->pb[1] += fa.y; ex->pb[2] += fa.z;
      ex->pb[3] += fl.x; ex->pb[4] += fl.y; ex->pb[5] += fl.z;
    }
  }
  float ab[6] = {red[21], red[22], red[23], red[24], red[25], red[26]};
  ldl6_solve(red, ab);
  alpha_b = v3(ab[0], ab[1], ab[2]);
  alin_b = v3(ab[3], ab[4], ab[5]);
  qdd[0] = r0; qdd[1] = r1; qdd[2] = r2;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    qdd[0] = fmaf(-Y[0][c], ab[c], qdd[0]); qdd[1] = fmaf(-Y[1][c], ab[c], qdd[1]); qdd[2] = fmaf(-Y[2][c], ab[c], qdd[2]);
  }
}
