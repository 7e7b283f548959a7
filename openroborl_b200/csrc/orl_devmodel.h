// Device-side packed copy of orl_robot_model: float4-aligned so that a lane fetches its leg's constants
// with LDS.128 from shared memory, contact points grouped per body with a bounding radius for culling.
// Built on the host by pack_model() (orl_env.cu) at orl_create time.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/orl.h"

// collision-point slots per leg on the device: the host's points grouped per body, each group padded to at least
// ORL_CP_UNROLL entries; the penetration mask is one 32-bit word
#define ORL_CP_UNROLL 2
#define ORL_DEV_MAX_CP 32

struct __align__(16) DevJoint {
  float4 r0, r1, r2;   // R_fix rows (xyz) + p_fix component (w)
  float4 i0;           // Ixx Ixy Ixz Iyy        (about the COM, body frame)
  float4 i1;           // Iyz Izz mass com.x
  float4 c1;           // com.y com.z lower upper
  float4 m0;           // damping dir offset kp
  float4 m1;           // kd init_angle motor(int bits) bounding radius of this body's collision points
  float4 i2;           // parallel-axis share of the inertia: xx xy xz yy   (scales with the randomised mass ratio)
  float4 i3;           // ... yz zz, randomizer class (int bits: 0 chassis set, 1 leg set), friction draw index (int bits, -1: none)
};
struct __align__(16) DevCp {
  float4 a;            // local xyz, radius
  float4 b;            // mu, margin, is_foot (int bits), unused
};
struct __align__(16) DevLeg {
  DevJoint j[3];
  float4 ee[2];        // end-effector link COM xyz, body (int bits)
  int4 meta;           // n_ee, base bounding radius (float bits), max margin of non-foot points (float bits), primary cp index
  int4 cnt;            // collision points on body 0, 1, 2 and on the base share; stored in that order in cp[]
  float4 bb[4][2];     // per body (0,1,2,base): body-frame bounding box of its NON-primary points: (centre xyz, max radius), (half extents, -)
  DevCp cp[ORL_DEV_MAX_CP];
};
// The four lanes of a quad read the same field of four different legs with one LDS.128: the legs' copies have to
// fall into different 16-byte bank groups (a 1600-byte DevLeg put legs 0/2 and 1/3 on the same banks and doubled
// the shared-memory wavefronts of every model load, profiles/r01_v4_step_kernel.md).
static_assert((sizeof(DevLeg) / 16) % 8 != 0 && (sizeof(DevLeg) / 16) % 8 != 4, "DevLeg stride maps two legs onto the same shared-memory banks");

struct __align__(16) DevModel {
  float4 b0;           // base mass, Ixx, Ixy, Ixz
  float4 b1;           // Iyy, Iyz, Izz, max_angle_change
  float4 ipos;         // init position
  float4 iquat;        // init orientation
  DevLeg leg[ORL_NUM_LEGS];
  int32_t n_links, pad0, pad1, pad2;
  int32_t link_leg[ORL_MAX_LINKS];
  int32_t link_body[ORL_MAX_LINKS];
  float4 link_com[ORL_MAX_LINKS];
};
