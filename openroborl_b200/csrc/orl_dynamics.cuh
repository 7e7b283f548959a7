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
#pragma once
#include "orl_math.cuh"
#include "../../include/orl.h"

struct PhysParams {
  float dt, g;              // g = -gravity_z (>0)
  float kn, bn, d0, bt;
  float klim, dlim, e0;
};

// rigid-body inertia at the frame-A origin: mass, first moment h = m c, rotational inertia about the origin
struct RBI { float m; float3 h; S3 I; };
ORL_DI RBI operator+(const RBI& a, const RBI& b) { RBI o; o.m = a.m + b.m; o.h = a.h + b.h; o.I = a.I + b.I; return o; }
struct SV { float3 a, l; };   // spatial vector: angular, linear
ORL_DI SV operator+(const SV& x, const SV& y) { SV o; o.a = x.a + y.a; o.l = x.l + y.l; return o; }
ORL_DI SV operator-(const SV& x, const SV& y) { SV o; o.a = x.a - y.a; o.l = x.l - y.l; return o; }
ORL_DI float sdot(const SV& x, const SV& y) { return dot(x.a, y.a) + dot(x.l, y.l); }
// I * [w; v] = [I w + h x v ; m v - h x w]
ORL_DI SV apply(const RBI& B, const SV& v) { SV o; o.a = mul(B.I, v.a) + cross(B.h, v.l); o.l = B.m * v.l - cross(B.h, v.a); return o; }
// motion cross product  v x m
ORL_DI SV crm(const SV& v, const SV& m) { SV o; o.a = cross(v.a, m.a); o.l = cross(v.a, m.l) + cross(v.l, m.a); return o; }
// force cross product  v x* f
ORL_DI SV crf(const SV& v, const SV& f) { SV o; o.a = cross(v.a, f.a) + cross(v.l, f.l); o.l = cross(v.a, f.l); return o; }

ORL_DI RBI body_inertia(float m, float3 c /* COM rel. origin, frame A */, const S3& Ic /* about COM, frame A */) {
  RBI o; o.m = m; o.h = m * c;
  const float cc = dot(c, c);
  o.I.xx = Ic.xx + m * (cc - c.x * c.x); o.I.yy = Ic.yy + m * (cc - c.y * c.y); o.I.zz = Ic.zz + m * (cc - c.z * c.z);
  o.I.xy = Ic.xy - m * c.x * c.y; o.I.xz = Ic.xz - m * c.x * c.z; o.I.yz = Ic.yz - m * c.y * c.z;
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
  float D[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float d = A[TRI(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= A[TRI(j, k)] * A[TRI(j, k)] * D[k];
    D[j] = d;
    const float inv = 1.0f / d;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      float s = A[TRI(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= A[TRI(i, k)] * A[TRI(j, k)] * D[k];
      A[TRI(i, j)] = s * inv;
    }
  }
#pragma unroll
  for (int i = 1; i < 6; ++i)
#pragma unroll
    for (int k = 0; k < i; ++k) b[i] -= A[TRI(i, k)] * b[k];
#pragma unroll
  for (int i = 0; i < 6; ++i) b[i] /= D[i];
#pragma unroll
  for (int i = 4; i >= 0; --i)
#pragma unroll
    for (int k = i + 1; k < 6; ++k) b[i] -= A[TRI(k, i)] * b[k];
}

// LDL^T of a symmetric positive definite 3x3: A = [a00 a10 a11 a20 a21 a22] (lower, row-major)
struct LDL3 { float l10, l20, l21, i0, i1, i2; };
ORL_DI LDL3 ldl3(const float A[6]) {
  LDL3 f;
  f.i0 = 1.0f / A[0];
  f.l10 = A[1] * f.i0; f.l20 = A[3] * f.i0;
  const float d1 = A[2] - f.l10 * A[1];
  f.i1 = 1.0f / d1;
  const float t = A[4] - f.l20 * A[1];
  f.l21 = t * f.i1;
  const float d2 = A[5] - f.l20 * A[3] - f.l21 * t;
  f.i2 = 1.0f / d2;
  return f;
}
ORL_DI void ldl3_solve(const LDL3& f, float& x0, float& x1, float& x2) {
  x1 -= f.l10 * x0; x2 -= f.l20 * x0 + f.l21 * x1;
  x0 *= f.i0; x1 *= f.i1; x2 *= f.i2;
  x1 -= f.l21 * x2; x0 -= f.l10 * x1 + f.l20 * x2;
}

// forward kinematics of one leg in frame A
struct LegFK { M3 R[3]; float3 p[3]; };
ORL_DI void leg_fk(const orl_leg_model& L, const M3& Rb, const float q[3], LegFK& K) {
  M3 Rp = Rb; float3 pp = v3(0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    M3 F;
    F.r0 = v3(L.R_fix[k][0], L.R_fix[k][1], L.R_fix[k][2]);
    F.r1 = v3(L.R_fix[k][3], L.R_fix[k][4], L.R_fix[k][5]);
    F.r2 = v3(L.R_fix[k][6], L.R_fix[k][7], L.R_fix[k][8]);
    float s, c;
    sincosf(q[k], &s, &c);
    K.p[k] = pp + mul(Rp, v3(L.p_fix[k][0], L.p_fix[k][1], L.p_fix[k][2]));
    K.R[k] = mul_rz(mul(Rp, F), s, c);
    Rp = K.R[k]; pp = K.p[k];
  }
}

// What the caller can ask leg_dynamics to export (parity entry points)
struct DynExport {
  float Hll[6];        // leg block of the mass matrix, lower triangle
  SV Fcol[3];          // base-leg coupling columns
  float Hbb[21];       // base block (full, after the 4-lane reduction)
  float C[3];          // joint bias (incl. gravity, excl. contact)
  float pb[6];         // base bias
};

// One dynamics evaluation for this lane's leg.  Returns the spatial base acceleration (a_b: angular
// alpha, linear a_lin so that d/dt v_com = a_lin + w x v) and this leg's joint accelerations.
// tau: generalised joint forces in URDF joint space (motor + limit + damping torques).
template <bool CONTACT, bool EXPORT>
ORL_DI void leg_dynamics(const orl_leg_model& L, const orl_robot_model& Mo, const PhysParams& P, const M3& Rb,
                         float base_z, float3 wb, float3 vb, const float q[3], const float qd[3], const float tau[3],
                         float3& alpha_b, float3& alin_b, float qdd[3], DynExport* ex) {
  LegFK K;
  leg_fk(L, Rb, q, K);
  // ---- velocities, velocity-product accelerations, per-body inertia and bias force -------------
  SV S[3], v[3], avp[3], f[3];
  RBI I[3];
  SV vp; vp.a = wb; vp.l = vb;
  SV ap; ap.a = v3(0.f, 0.f, 0.f); ap.l = v3(0.f, 0.f, P.g);    // a0 = -a_gravity
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    S[k].a = col(K.R[k], 2);
    S[k].l = cross(K.p[k], S[k].a);
    SV sq; sq.a = qd[k] * S[k].a; sq.l = qd[k] * S[k].l;
    v[k] = vp + sq;
    avp[k] = ap + crm(vp, sq);
    const float3 c = K.p[k] + mul(K.R[k], v3(L.com[k][0], L.com[k][1], L.com[k][2]));
    S3 Ic = {L.inertia[k][0], L.inertia[k][1], L.inertia[k][2], L.inertia[k][3], L.inertia[k][4], L.inertia[k][5]};
    I[k] = body_inertia(L.mass[k], c, rot_sym(K.R[k], Ic));
    f[k] = apply(I[k], avp[k]) + crf(v[k], apply(I[k], v[k]));
    vp = v[k]; ap = avp[k];
  }
  // ---- contact: explicit force -> external force; implicit part -> W blocks --------------------
  float Wbb[21], Wll[6];
  SV Wbl[3];
  SV fb_ext; fb_ext.a = v3(0.f, 0.f, 0.f); fb_ext.l = v3(0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 21; ++i) Wbb[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) Wll[i] = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) { Wbl[k].a = v3(0.f, 0.f, 0.f); Wbl[k].l = v3(0.f, 0.f, 0.f); }
  if (CONTACT) {
    const float3 e_cl = cross(wb, vb);     // classical - spatial linear base acceleration
    for (int i = 0; i < L.n_cp; ++i) {
      const orl_contact_point& cp = L.cp[i];
      const int kb = cp.body;
      const float3 loc = v3(cp.local[0], cp.local[1], cp.local[2]);
      float3 rz_row; float pz;
      if (kb < 0) { rz_row = Rb.r2; pz = 0.f; }
      else if (kb == 0) { rz_row = K.R[0].r2; pz = K.p[0].z; }
      else if (kb == 1) { rz_row = K.R[1].r2; pz = K.p[1].z; }
      else { rz_row = K.R[2].r2; pz = K.p[2].z; }
      const float z = base_z + pz + dot(rz_row, loc) - cp.radius;
      if (z < 0.f) {
        M3 Rk; float3 pk; SV vk;
        if (kb < 0) { Rk = Rb; pk = v3(0.f, 0.f, 0.f); vk.a = wb; vk.l = vb; }
        else if (kb == 0) { Rk = K.R[0]; pk = K.p[0]; vk = v[0]; }
        else if (kb == 1) { Rk = K.R[1]; pk = K.p[1]; vk = v[1]; }
        else { Rk = K.R[2]; pk = K.p[2]; vk = v[2]; }
        float3 r = pk + mul(Rk, loc); r.z -= cp.radius;            // contact location rel. frame-A origin
        const float3 V = vk.l + cross(vk.a, r);
        const float d = -z;
        const float sc = fminf(1.0f, d / P.d0);
        const float fn = fmaxf(0.0f, P.kn * d - P.bn * sc * V.z);
        const float vt = sqrtf(V.x * V.x + V.y * V.y);
        const float ct = fminf(P.bt, cp.mu * fn / fmaxf(vt, 1e-9f));
        const float cn = fn > 0.f ? P.kn * P.dt + P.bn * sc : 0.f;
        const float3 D = v3(ct, ct, cn);
        // force at the old velocity, corrected for the classical/spatial acceleration difference
        float3 F = v3(-ct * V.x, -ct * V.y, fn);
        F = F - P.dt * v3(D.x * e_cl.x, D.y * e_cl.y, D.z * e_cl.z);
        SV fe; fe.a = cross(r, F); fe.l = F;
        if (kb < 0) fb_ext = fb_ext + fe;
        else if (kb == 0) f[0] = f[0] - fe;
        else if (kb == 1) f[1] = f[1] - fe;
        else f[2] = f[2] - fe;
        // Jacobian columns of the contact-point velocity: base w (3), base v (3), joints 0..kb
        float3 J[9];
        J[0] = v3(0.f, -r.z, r.y); J[1] = v3(r.z, 0.f, -r.x); J[2] = v3(-r.y, r.x, 0.f);   // e_a x r
        J[3] = v3(1.f, 0.f, 0.f); J[4] = v3(0.f, 1.f, 0.f); J[5] = v3(0.f, 0.f, 1.f);
#pragma unroll
        for (int j = 0; j < 3; ++j) J[6 + j] = (j <= kb) ? cross(S[j].a, r - K.p[j]) : v3(0.f, 0.f, 0.f);
        const float3 Dd = P.dt * D;
        float3 DJ[9];
#pragma unroll
        for (int a = 0; a < 9; ++a) DJ[a] = v3(Dd.x * J[a].x, Dd.y * J[a].y, Dd.z * J[a].z);
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b) Wbb[TRI(a, b)] += dot(DJ[a], J[b]);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          Wbl[j].a += v3(dot(DJ[0], J[6 + j]), dot(DJ[1], J[6 + j]), dot(DJ[2], J[6 + j]));
          Wbl[j].l += v3(dot(DJ[3], J[6 + j]), dot(DJ[4], J[6 + j]), dot(DJ[5], J[6 + j]));
        }
        Wll[0] += dot(DJ[6], J[6]); Wll[1] += dot(DJ[7], J[6]); Wll[2] += dot(DJ[7], J[7]);
        Wll[3] += dot(DJ[8], J[6]); Wll[4] += dot(DJ[8], J[7]); Wll[5] += dot(DJ[8], J[8]);
      }
    }
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
  float C0 = sdot(S[0], fc0), C1 = sdot(S[1], fc1), C2 = sdot(S[2], fc2);
  if (EXPORT) {
#pragma unroll
    for (int i = 0; i < 6; ++i) ex->Hll[i] = A[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) ex->Fcol[k] = Fk[k];
    ex->C[0] = C0; ex->C[1] = C1; ex->C[2] = C2;
  }
  // ---- lane-local Schur complement ---------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 6; ++i) A[i] += Wll[i];
  SV B[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) B[k] = Fk[k] + Wbl[k];
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
      red[TRI(a, b)] += Wbb[TRI(a, b)] - (Bm[a][0] * Y[0][b] + Bm[a][1] * Y[1][b] + Bm[a][2] * Y[2][b]);
  {
    const float pbv[6] = {fc0.a.x - fb_ext.a.x, fc0.a.y - fb_ext.a.y, fc0.a.z - fb_ext.a.z,
                          fc0.l.x - fb_ext.l.x, fc0.l.y - fb_ext.l.y, fc0.l.z - fb_ext.l.z};
#pragma unroll
    for (int a = 0; a < 6; ++a) red[21 + a] = -pbv[a] - (Bm[a][0] * r0 + Bm[a][1] * r1 + Bm[a][2] * r2);
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
    S3 Ib = {Mo.base_inertia[0], Mo.base_inertia[1], Mo.base_inertia[2], Mo.base_inertia[3], Mo.base_inertia[4], Mo.base_inertia[5]};
    RBI Bb = body_inertia(Mo.base_mass, v3(0.f, 0.f, 0.f), rot_sym(Rb, Ib));
    SV vb6; vb6.a = wb; vb6.l = vb;
    SV a0; a0.a = v3(0.f, 0.f, 0.f); a0.l = v3(0.f, 0.f, P.g);
    SV fbody = apply(Bb, a0) + crf(vb6, apply(Bb, vb6));
    float tb[21];
    rbi_to_tri(Bb, tb);
#pragma unroll
    for (int i = 0; i < 21; ++i) red[i] += tb[i];
    red[21] -= fbody.a.x; red[22] -= fbody.a.y; red[23] -= fbody.a.z;
    red[24] -= fbody.l.x; red[25] -= fbody.l.y; red[26] -= fbody.l.z;
    if (EXPORT) {
#pragma unroll
      for (int i = 0; i < 21; ++i) ex->Hbb[i] += tb[i];
      ex->pb[0] += fbody.a.x; ex->pb[1] += fbody.a.y; ex->pb[2] += fbody.a.z;
      ex->pb[3] += fbody.l.x; ex->pb[4] += fbody.l.y; ex->pb[5] += fbody.l.z;
    }
  }
  float ab[6] = {red[21], red[22], red[23], red[24], red[25], red[26]};
  ldl6_solve(red, ab);
  alpha_b = v3(ab[0], ab[1], ab[2]);
  alin_b = v3(ab[3], ab[4], ab[5]);
  qdd[0] = r0; qdd[1] = r1; qdd[2] = r2;
#pragma unroll
  for (int c = 0; c < 6; ++c) { qdd[0] -= Y[0][c] * ab[c]; qdd[1] -= Y[1][c] * ab[c]; qdd[2] -= Y[2][c] * ab[c]; }
}
