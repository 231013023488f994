// llq_kernels.cuh -- sm_100a kernels of the batched quadruped rollout engine.
//
// Mapping: one environment = 4 adjacent lanes of a warp, one lane per leg (FR, FL, HR, HL);
// a warp therefore advances 8 environments.  Base quantities are replicated on the 4 lanes, the per-leg
// 3-joint chain recursion runs lane-local, and the only cross-lane traffic is
//   * the reduction of the legs' articulated inertia / bias force into the base (27 floats, xor-shuffles),
//   * broadcast of unit-impulse base responses and of each Gauss-Seidel row's impulse (shfl, width 4).
// All ten 2 ms sub-steps of one 50 Hz policy step run inside one launch with the state in registers.
//
// Replaces (reference, relative to src/lifelike/sim_envs/pybullet_envs/):
//   PrimitiveLevelEnv.step                     primitive_level_env/primitive_level_env.py:195-245
//   LeggedRobot.apply_action                   legged_robot/legged_robot.py:119-148
//   pybullet stepSimulation (Bullet btMultiBody ABA + PGS)
//   MotionLib.step/get_states_info(_future)    primitive_level_env/motion_lib.py:65-166
//   _prepare_obs/_compute_reward/_check_terminate   primitive_level_env.py:276-426
#pragma once
#ifndef LLQ_BARRIERS
#define LLQ_BARRIERS 1   // CTA barriers per sub-step that keep the warps on the same code stretch (instruction-cache sharing)
#endif
#include "llq_math.cuh"
#include <cuda_pipeline.h>
#include <stdint.h>

namespace llq {

constexpr int kObsDim = 207, kObsDimEpmc = 916, kObsDimSepmc = 965, kPropDim = 33, kActDim = 12, kStateDim = 37, kAuxDim = 18;
template <int ENV> struct ObsW { static constexpr int value = (ENV == 1 || ENV == 3) ? kObsDimEpmc : (ENV == 2 ? kObsDimSepmc : kObsDim); };
constexpr int kMaxBoxes = 36, kMaxCand = 12;   // ENV 3 = EPMC corridor (elements 1-3): static boxes per env, contact candidates per step
constexpr int kNewObs = 120;
constexpr int kRowFloats = 153;  // per-lane floats of the constraint-row workspace in shared memory (Yc 18 | Yl 18 | Ul 9 | Acl 36 | Alc 36 | All 36)  // floats staged per env: prop 33 | action 12 | future 72 (+3 pad)

struct DampItem { float m; float c[3]; float Ic[6]; };
struct JointConst {
  float r[3];
  float m; float h[3]; float I[6];
  int nd; DampItem d[2];
  float lower, upper, jdamp; int haslim;
};
struct LegConst { JointConst j[3]; float foot[3]; float foot_r; float pad[4]; };
struct BaseConst { float qI[4]; float m; float h[3]; float I[6]; int nd; DampItem d[3]; };
struct alignas(16) ModelConst {
  BaseConst base; LegConst leg[4];
  float push_R[9]; float push_c[3];   // FR hip link: inertial-frame rotation (link <- inertial) and CoM, for applyExternalForce(LINK_FRAME) (PR:73-77)
  float init_state[37]; float pad_[3]; // EPMC episode start state (LR:115-117, utils/constants.py:103-116)
  // detection proxies for the PMC hurdle plate: wheel (knee) centre in the thigh frame + radius, hip radius, body-box corners (base coords)
  float wheel_off[4][3]; float wheel_r[4]; float hip_r[4]; float corner[8][3];
  float handle[2][4];                  // SEPMC: front / hind handle centre (base coords) + radius (LR:150-156)
};

struct MocapFrame { double x, y, z, pad; float quat[4]; float q[12]; };  // 96 B, 16-byte aligned

struct StepParams {
  int n_envs, substeps, solver_iters;
  float dt, kp, kd, max_tau, gz, mu, erp, jerp, slop, warm, breaking, kl, ka, vmax, max_imp;
  float w_jp, w_jv, w_ee, w_pose, w_vel;   // already normalised to sum 1
  double sim_dt, frame_dt;
  int margin;
  // EPMC (PGE / PR)
  int max_steps, cmd_freq_lo, cmd_freq_hi, push_start_count, push_interval, push_duration, push_enabled;
  float mu_ground, fr_lo, fr_hi, ph_lo, ph_hi, pv_lo, pv_hi, ts_lo, ts_hi;
  // PMC hurdle plates (PLE:173-193)
  int has_ob; float ob_hx, ob_hy, ob_hz;
  // EPMC corridor (BSE)
  int element_id; float ww_lo, ww_hi, wg_lo, wg_hi, hg_lo, hg_hi;
  // knee-wheel ground contact (llq_config.knee_contacts / link_friction)
  int knee; float mu_wheel;
  float aux_r;      // EPMC elements 1-3: radius of the auxiliary edge cylinders (0 = none)
};

struct EnvArrays {      // SoA device arrays, N envs
  double* pos;          // [3][N]
  float* st;            // [34][N]: quat4 lin3 ang3 q12 qd12
  double* time;         // [N]
  int* clip;            // [N]
  float* reward_sum;    // [N]
  int* episode_steps;   // [N]
  long long* episode;   // [N]
  float* warm;          // [4][N]
  float* obs;           // [N][207] (history carry)
  float* kin;           // [37][N]
  float* foot_pos;      // [12][N]
  float* done_reward;   // [N] reward_sum at termination
  unsigned char* done;  // [N]
  float* reward;        // [N]
  unsigned long long* counters;  // [8]
  double* aux;          // [18][N] EPMC bookkeeping (include/llq.h LLQ_F_AUX)
  int* ob_id;           // [N] active hurdle plate
  float* boxes;         // [N][36][6] EPMC corridor: centre xyz, half extents xyz (walls first)
  int* nbox;            // [N]
};

struct MocapDev { const MocapFrame* frames; const int* clip_off; int n_clips; const double* ob_table; const int* ob_off; };

#define FULL 0xffffffffu

LLQ_DI float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
LLQ_DI float gsum4(float v) {  // sum over the 4 lanes of an env, result on all 4
  v += __shfl_xor_sync(FULL, v, 1);
  v += __shfl_xor_sync(FULL, v, 2);
  return v;
}
LLQ_DI V3 ld3(const float* p) { return V3{p[0], p[1], p[2]}; }
LLQ_DI Sym3 ldsym(const float* p) { return Sym3{p[0], p[1], p[2], p[3], p[4], p[5]}; }

// spatial motion / force vectors (angular, linear) at a link origin, link coordinates
struct SV { V3 a, l; };

// bias force of a rigid body (composite m, h, I about the frame origin) moving with (w, v):  v x* (I v)  plus
// Bullet's per-link damping  m v_c (k + k|v_c|),  I_c w (k + k|w|)  for each original URDF link in the composite.
template <int ND> LLQ_DI SV bias_force(float m, V3 h, Sym3 I, int nd, const DampItem* d, V3 w, V3 v, float kl, float ka) {
  V3 hl = fma3(m, v, cross(w, h));
  V3 ha = mul(I, w) + cross(h, v);
  SV p;
  p.a = cross(w, ha) + cross(v, hl);
  p.l = cross(w, hl);
  float wn = norm3(w);
#pragma unroll
  for (int t = 0; t < ND; t++) {
    if (t < nd) {
      V3 c = ld3(d[t].c);
      V3 vc = v + cross(w, c);
      V3 f = (d[t].m * (kl + kl * norm3(vc))) * vc;
      V3 n = (ka + ka * wn) * mul(ldsym(d[t].Ic), w);
      p.l = p.l + f;
      p.a = p.a + n + cross(c, f);
    }
  }
  return p;
}

// articulated inertia blocks: f_ang = A w + B v ; f_lin = B^T w + C v
struct ABI { Sym3 A; M3 B; Sym3 C; };

LLQ_DI ABI rigid_abi(float m, V3 h, Sym3 I) {
  ABI r; r.A = I; r.B = skew(h); r.C = Sym3{m, 0.f, 0.f, m, 0.f, m};
  return r;
}

// per-joint cache kept for the whole sub-step
struct JC { float c, s; V3 Ua, Ul; float Dinv, u; };

// Reduce a 1-dof joint about coordinate axis AX (sign SG) out of (I, p), then express the result in the parent frame
// (rotation E = Rot(AX, angle) with (c, s), origin offset r).  cor = velocity-product acceleration of the link.
template <int AX, int SG> LLQ_DI void joint_reduce(ABI& I, SV& p, SV cor, float tau, V3 r, JC& jc) {
  const float sg = (float)SG;
  V3 Ua = sg * col(I.A, AX), Ul = sg * row(I.B, AX);
  float D = diag(I.A, AX), Dinv = 1.0f / D;
  float u = tau - sg * comp(p.a, AX);
  jc.Ua = Ua; jc.Ul = Ul; jc.Dinv = Dinv; jc.u = u;
  I.A = sub_outer(I.A, Ua, Dinv);
  I.B = sub_outer(I.B, Ua, Ul, Dinv);
  I.C = sub_outer(I.C, Ul, Dinv);
  float ud = u * Dinv;
  V3 pa = p.a + mul(I.A, cor.a) + mul(I.B, cor.l) + ud * Ua;
  V3 pl = p.l + tmul(I.B, cor.a) + mul(I.C, cor.l) + ud * Ul;
  // rotate into parent axes
  Sym3 Ar = rot_sym<AX>(I.A, jc.c, jc.s), Cr = rot_sym<AX>(I.C, jc.c, jc.s);
  M3 Br = rot_mat<AX>(I.B, jc.c, jc.s);
  V3 par = rot<AX>(pa, jc.c, jc.s), plr = rot<AX>(pl, jc.c, jc.s);
  // translate by r:  C' = C ; B' = B + rx C ; A' = A - K - K^T - (rx C) rx,  K = B rx (rows of B crossed with r)
  V3 g0 = cross(r, col(Cr, 0)), g1 = cross(r, col(Cr, 1)), g2 = cross(r, col(Cr, 2));   // columns of G = rx C
  M3 G = M3{g0.x, g1.x, g2.x, g0.y, g1.y, g2.y, g0.z, g1.z, g2.z};
  V3 k0 = cross(row(Br, 0), r), k1 = cross(row(Br, 1), r), k2 = cross(row(Br, 2), r);   // rows of K
  V3 h0 = cross(row(G, 0), r), h1 = cross(row(G, 1), r), h2 = cross(row(G, 2), r);      // rows of G rx (symmetric)
  I.A = Sym3{Ar.xx - 2.f * k0.x - h0.x, Ar.xy - k0.y - k1.x - h0.y, Ar.xz - k0.z - k2.x - h0.z,
             Ar.yy - 2.f * k1.y - h1.y, Ar.yz - k1.z - k2.y - h1.z, Ar.zz - 2.f * k2.z - h2.z};
  I.B = Br + G;
  I.C = Cr;
  p.a = par + cross(r, plr);
  p.l = plr;
}

// motion transform parent -> child:  a_c = E^T a_p ; l_c = E^T (l_p + a_p x r)
template <int AX> LLQ_DI SV xmotion(SV vp, V3 r, float c, float s) {
  SV o; o.a = rotT<AX>(vp.a, c, s); o.l = rotT<AX>(vp.l + cross(vp.a, r), c, s);
  return o;
}
// force transform child -> parent:  n_p = E n_c + r x (E f_c) ; f_p = E f_c
template <int AX> LLQ_DI SV xforce(SV fc, V3 r, float c, float s) {
  SV o; o.l = rot<AX>(fc.l, c, s); o.a = rot<AX>(fc.a, c, s) + cross(r, o.l);
  return o;
}

// Unit-impulse response (Bullet calcAccelerationDeltasMultiDof): up pass on the owning lane.
// F3: spatial force applied at the shank origin (shank coords); t1..t3: joint torques.  Returns -Z0 (the force the
// base sees, base coords) and the per-joint u's.
LLQ_DI SV response_up(const JC (&jc)[3], const V3 (&r)[3], SV F3, float t1, float t2, float t3, float (&u)[3]) {
  SV Z = SV{neg(F3.a), neg(F3.l)};
  u[2] = t3 - (-1.f) * Z.a.y;                               // S3 = (0,-1,0 | 0)
  float ud = u[2] * jc[2].Dinv;
  Z.a = fma3(ud, jc[2].Ua, Z.a); Z.l = fma3(ud, jc[2].Ul, Z.l);
  Z = xforce<1>(Z, r[2], jc[2].c, jc[2].s);
  u[1] = t2 - (-1.f) * Z.a.y;
  ud = u[1] * jc[1].Dinv;
  Z.a = fma3(ud, jc[1].Ua, Z.a); Z.l = fma3(ud, jc[1].Ul, Z.l);
  Z = xforce<1>(Z, r[1], jc[1].c, jc[1].s);
  u[0] = t1 - Z.a.x;                                        // S1 = (1,0,0 | 0)
  ud = u[0] * jc[0].Dinv;
  Z.a = fma3(ud, jc[0].Ua, Z.a); Z.l = fma3(ud, jc[0].Ul, Z.l);
  Z = xforce<0>(Z, r[0], jc[0].c, jc[0].s);
  return SV{neg(Z.a), neg(Z.l)};
}
// down pass on every lane: base acceleration a0 (base coords) -> joint accelerations of this lane's leg
LLQ_DI void response_down(const JC (&jc)[3], const V3 (&r)[3], SV a0, float u1, float u2, float u3, float (&qdd)[3]) {
  SV a = xmotion<0>(a0, r[0], jc[0].c, jc[0].s);
  qdd[0] = (u1 - dot(jc[0].Ua, a.a) - dot(jc[0].Ul, a.l)) * jc[0].Dinv;
  a.a.x += qdd[0];
  a = xmotion<1>(a, r[1], jc[1].c, jc[1].s);
  qdd[1] = (u2 - dot(jc[1].Ua, a.a) - dot(jc[1].Ul, a.l)) * jc[1].Dinv;
  a.a.y -= qdd[1];
  a = xmotion<1>(a, r[2], jc[2].c, jc[2].s);
  qdd[2] = (u3 - dot(jc[2].Ua, a.a) - dot(jc[2].Ul, a.l)) * jc[2].Dinv;
}

// ---------------------------------------------------------------------------------------------------------------
// Mocap interpolation (motion_lib.py:88-166), fp32 except positions/time (fp64)
struct KinBase { double px, py, pz; Q4 q; V3 lin, ang; };

LLQ_DI Q4 ldq(const float* p) { float4 v = *reinterpret_cast<const float4*>(p); return Q4{v.x, v.y, v.z, v.w}; }

LLQ_DI KinBase mocap_base(const MocapFrame* fc, const MocapFrame* fn, double frac, double frame_dt) {
  KinBase k;
  double cx = fc->x, cy = fc->y, cz = fc->z, nx = fn->x, ny = fn->y, nz = fn->z;
  k.px = cx + frac * (nx - cx); k.py = cy + frac * (ny - cy); k.pz = cz + frac * (nz - cz);
  float inv = (float)(1.0 / frame_dt);
  k.lin = V3{(float)(nx - cx) * inv, (float)(ny - cy) * inv, (float)(nz - cz) * inv};
  Q4 qc = qnormalize(ldq(fc->quat)), qn = qnormalize(ldq(fn->quat));
  V3 rv = q_rotvec(qmul(qconj(qc), qn));
  k.q = qmul(qc, rotvec_q((float)frac * rv));
  V3 rw = q_rotvec(qmul(qn, qconj(qc)));
  float angle = norm3(rw);
  float sc = angle / (angle + 1e-8f) * inv;
  k.ang = sc * rw;
  return k;
}

// Foot (link *4) world position for a robot state given in the pybullet base-inertial convention:
// R_bp = world <- B' (URDF body axes), p = base CoM.  q = this leg's joint angles.
LLQ_DI V3 foot_in_base(const LegConst& L, float q1, float q2, float q3) {
  float c1, s1, c2, s2, c3, s3;
  llq_sincosf(q1, &s1, &c1); llq_sincosf(-q2, &s2, &c2); llq_sincosf(-q3, &s3, &c3);
  V3 p = rot<1>(ld3(L.foot), c3, s3) + ld3(L.j[2].r);
  p = rot<1>(p, c2, s2) + ld3(L.j[1].r);
  p = rot<0>(p, c1, s1) + ld3(L.j[0].r);
  return p;
}

// Philox4x32-10
LLQ_DI void philox4x32_10(unsigned (&c)[4], unsigned k0, unsigned k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    unsigned n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// distance test of a sphere (centre w in world, radius r) against the yawed plate centred at (ox, oy, 0)
LLQ_DI bool plate_hit(V3 w, float r, float cy, float sy, float hx, float hy, float hz, float thr) {
  float bx = cy * w.x + sy * w.y, by = -sy * w.x + cy * w.y;
  float qx = bx - clampf(bx, -hx, hx), qy = by - clampf(by, -hy, hy), qz = w.z - clampf(w.z, -hz, hz);
  return sqrtf(qx * qx + qy * qy + qz * qz) - r < thr;
}

// four uniforms of stream `stream` (1 = EPMC reset, 2 = push randomiser, 3 = joystick command), draw `index` (matches the oracle)
LLQ_DI void stream_uniforms(unsigned long long seed, long long gid, long long episode, unsigned stream, unsigned index, double (&u)[4]) {
  unsigned c[4] = {(unsigned)gid, ((unsigned)((unsigned long long)gid >> 32) & 0x00FFFFFFu) | (stream << 24), (unsigned)episode, index};
  philox4x32_10(c, (unsigned)seed, (unsigned)(seed >> 32));
#pragma unroll
  for (int i = 0; i < 4; i++) u[i] = ((double)c[i] + 0.5) * (1.0 / 4294967296.0);
}
LLQ_DI void epmc_randomize_push(const StepParams& P, unsigned long long seed, long long gid, long long ep, int& push_draws, float (&pf)[3]) {   // PR:89-99
  double u[4];
  stream_uniforms(seed, gid, ep, 2, (unsigned)push_draws++, u);
  double sn, cs;
  sincos(2.0 * 3.14159265358979323846 * u[0], &sn, &cs);
  double h = (double)P.ph_lo + u[1] * ((double)P.ph_hi - (double)P.ph_lo);
  pf[0] = (float)(h * cs); pf[1] = (float)(h * sn); pf[2] = (float)((double)P.pv_lo + u[2] * ((double)P.pv_hi - (double)P.pv_lo));
}


LLQ_DI void push_force_of_draw(const StepParams& P, unsigned long long seed, long long gid, long long ep, int index, float (&pf)[3]) {
  int d = index;
  epmc_randomize_push(P, seed, gid, ep, d, pf);
}

// ---------------------------------------------------------------------------------------------------------------
// SEPMC (ChaseTagGameEnv, max_game/chase_tag_game_env.py = CTG; arena = max_game/bullet_static_entities.py:863-902):
// robots 2p and 2p+1 are the two agents of pair p (adjacent 4-lane groups of one warp).
constexpr float kWallIn = 2.495f;    // inner faces of the four 0.01 m walls centred at +-2.5
// closest hit fraction of the segment o -> o + d against the arena's static boxes (ground slab, 4 walls, flag), or -1
LLQ_DI float ray_box1(V3 o, V3 d, V3 lo, V3 hi, float best) {
  if (o.x > lo.x && o.x < hi.x && o.y > lo.y && o.y < hi.y && o.z > lo.z && o.z < hi.z) return best;   // starts inside: no hit
  float t0 = 0.f, t1 = 1.f;
  bool hit = true, entered = false;
  const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    if (dd[ax] == 0.f) { if (oo[ax] < l[ax] || oo[ax] > h[ax]) hit = false; continue; }
    const float inv = 1.0f / dd[ax];
    float ta = (l[ax] - oo[ax]) * inv, tb = (h[ax] - oo[ax]) * inv;
    if (ta > tb) { const float t = ta; ta = tb; tb = t; }
    if (ta > t0) { t0 = ta; entered = true; }
    t1 = fminf(t1, tb);
    if (t0 > t1) hit = false;
  }
  if (hit && entered && (best < 0.f || t0 < best)) best = t0;
  return best;
}
// Fast paths of ray_arena for origins inside the arena (|x|, |y| < 2.49): the four walls are then hit on their inner faces, the
// first one crossed is always a valid hit of that face in xy, so only its z range needs checking; same slab arithmetic.
LLQ_DI float ray_arena_inside(V3 o, V3 d, float fx, float fy) {
  float best = -1.f;
  if (d.z < 0.f && o.z > 0.f) {                                    // ground slab top
    const float t = (0.f - o.z) * (1.0f / d.z);
    if (t <= 1.f) best = t;
  }
  float tw = 2.f; 
  if (d.x != 0.f) { const float t = ((d.x > 0.f ? kWallIn : -kWallIn) - o.x) * (1.0f / d.x); tw = t; }
  if (d.y != 0.f) { const float t = ((d.y > 0.f ? kWallIn : -kWallIn) - o.y) * (1.0f / d.y); tw = fminf(tw, t); }
  if (tw <= 1.f) {
    const float z = fmaf(tw, d.z, o.z);
    if (z >= 0.f && z <= 2.f && (best < 0.f || tw < best)) best = tw;
  }
  // the flag is a 0.1 m column: only rays whose line passes within its circumscribed radius (in xy) need the slab test
  const float cx = fx - o.x, cy = fy - o.y, cr = d.x * cy - d.y * cx;
  if (cr * cr > 0.00501f * (d.x * d.x + d.y * d.y)) return best;
  return ray_box1(o, d, V3{fx - 0.05f, fy - 0.05f, 0.f}, V3{fx + 0.05f, fy + 0.05f, 0.5f}, best);
}
LLQ_DI float ray_arena(V3 o, V3 d, float fx, float fy) {
  float b = -1.f;
  b = ray_box1(o, d, V3{-100.f, -100.f, -10.f}, V3{100.f, 100.f, 0.f}, b);
  b = ray_box1(o, d, V3{-2.5f, 2.495f, 0.f}, V3{2.5f, 2.505f, 2.f}, b);
  b = ray_box1(o, d, V3{-2.5f, -2.505f, 0.f}, V3{2.5f, -2.495f, 2.f}, b);
  b = ray_box1(o, d, V3{2.495f, -2.5f, 0.f}, V3{2.505f, 2.5f, 2.f}, b);
  b = ray_box1(o, d, V3{-2.505f, -2.5f, 0.f}, V3{-2.495f, 2.5f, 2.f}, b);
  b = ray_box1(o, d, V3{fx - 0.05f, fy - 0.05f, 0.f}, V3{fx + 0.05f, fy + 0.05f, 0.5f}, b);
  return b;
}
LLQ_DI float flag_dist(V3 c, float fx, float fy) {    // distance of a point to the flag box (CTG:163-190)
  const float qx = c.x - clampf(c.x, fx - 0.05f, fx + 0.05f), qy = c.y - clampf(c.y, fy - 0.05f, fy + 0.05f), qz = c.z - clampf(c.z, 0.f, 0.5f);
  return sqrtf(qx * qx + qy * qy + qz * qz);
}
// points of this lane's leg in base coordinates (B' axes about the base reference point)
LLQ_DI void leg_points(const ModelConst& M, const LegConst& L, int k, const float (&q)[3], V3& hip, V3& wheel, V3& foot) {
  float c1, s1, c2, s2, c3, s3;
  llq_sincosf(q[0], &s1, &c1); llq_sincosf(-q[1], &s2, &c2); llq_sincosf(-q[2], &s3, &c3);
  hip = ld3(L.j[0].r);
  const V3 p2 = hip + rot<0>(ld3(L.j[1].r), c1, s1);
  wheel = p2 + rot<0>(rot<1>(ld3(M.wheel_off[k]), c2, s2), c1, s1);
  V3 f = rot<1>(ld3(L.foot), c3, s3) + ld3(L.j[2].r);
  f = rot<1>(f, c2, s2) + ld3(L.j[1].r);
  foot = rot<0>(f, c1, s1) + hip;
}
// force of push-randomiser draw `index` (PR:89-99)

// staging row of SEPMC (kNewObs floats per robot): prop 33 | action 12 | R 9 (45) | pos 3 (54) | flag xy 2 (57) | yaw 1 (59) | pad 2 |
// small vectors 52 (62): percept_vec 5, oppo_info 15, oppo_info_cheat 15, flag_info 7, flag_info_cheat 7, with_flag 2, control_spd 1
struct PairState { int with_flag, flag_draws, visible, sw; double flag_x, flag_y; };

// End-of-step pair logic shared by the step and the reset kernels (CTG:495-596, 472-493): visibility, flag switch, the small
// observation vectors.  Every lane of both robots runs it; `snew` is the robot's staging row, `spart` the partner's.
template <int PX = 4>   // lane distance of the partner robot: 4 (one robot = 4 lanes) or 16 (one robot = a half-warp)
LLQ_DI void sepmc_pair_tail(const ModelConst& M, const LegConst& L, int k, int robot, float* snew, const float* spart, double px, double py,
                            double pz, Q4 qp, Q4 qb, V3 vw, V3 ww, const float (&q)[3], bool touch_own, float fix_spd, unsigned long long seed,
                            long long pair_gid, long long epi, PairState& S) {
  const M3 Rp = qmat(qp);
  const V3 pos = V3{(float)px, (float)py, (float)pz};
  // own convex points (LR:150-156) in world coordinates -> staging row, read by the partner
  {
    V3 hip, wheel, foot;
    leg_points(M, L, k, q, hip, wheel, foot);
    const V3 fw = pos + mul(Rp, foot), ww_ = pos + mul(Rp, wheel);
    snew[3 * k] = fw.x; snew[3 * k + 1] = fw.y; snew[3 * k + 2] = fw.z;
    snew[12 + 3 * k] = ww_.x; snew[13 + 3 * k] = ww_.y; snew[14 + 3 * k] = ww_.z;
    if (k < 2) {
      const V3 hw = pos + mul(Rp, V3{M.handle[k][0], M.handle[k][1], M.handle[k][2]});
      snew[24 + 3 * k] = hw.x; snew[25 + 3 * k] = hw.y; snew[26 + 3 * k] = hw.z;
    }
  }
  __syncwarp();
  // partner's root state
  const double ox = __shfl_xor_sync(FULL, px, PX), oy = __shfl_xor_sync(FULL, py, PX), oz = __shfl_xor_sync(FULL, pz, PX);
  const Q4 oq = Q4{__shfl_xor_sync(FULL, qb.x, PX), __shfl_xor_sync(FULL, qb.y, PX), __shfl_xor_sync(FULL, qb.z, PX), __shfl_xor_sync(FULL, qb.w, PX)};
  const V3 ov = V3{__shfl_xor_sync(FULL, vw.x, PX), __shfl_xor_sync(FULL, vw.y, PX), __shfl_xor_sync(FULL, vw.z, PX)};
  const V3 oww = V3{__shfl_xor_sync(FULL, ww.x, PX), __shfl_xor_sync(FULL, ww.y, PX), __shfl_xor_sync(FULL, ww.z, PX)};
  const bool touch_other = __shfl_xor_sync(FULL, touch_own ? 1 : 0, PX) != 0;
  const V3 opos = V3{(float)ox, (float)oy, (float)oz};
  const float fx = (float)S.flag_x, fy = (float)S.flag_y;
  // visibility (CTG:472-493): the root segment is cast from robot 0 to robot 1 for both agents
  const V3 ra = robot == 0 ? pos : opos, rb = robot == 0 ? opos : pos;
  bool vis = ray_arena(ra, rb - ra, fx, fy) < 0.f;
  {
    const V3 head = V3{snew[24], snew[25], snew[26]};
    const V3 tf = V3{spart[3 * k], spart[3 * k + 1], spart[3 * k + 2]}, tw = V3{spart[12 + 3 * k], spart[13 + 3 * k], spart[14 + 3 * k]};
    bool any = ray_arena(head, tf - head, fx, fy) < 0.f || ray_arena(head, tw - head, fx, fy) < 0.f;
    if (k < 2) {
      const V3 th = V3{spart[24 + 3 * k], spart[25 + 3 * k], spart[26 + 3 * k]};
      any = any || ray_arena(head, th - head, fx, fy) < 0.f;
    }
    int a = any ? 1 : 0;
    a |= __shfl_xor_sync(FULL, a, 1);
    a |= __shfl_xor_sync(FULL, a, 2);
    vis = vis || a != 0;
  }
  const Q4 q1 = qnormalize(qb);
  const M3 Rq = qmat(q1);
  {
    // cos of the bearing of the opponent against visible_angle = pi; in fp64 so that |cos| <= 1 holds unless fp64 itself rounds over
    const double c = (double)Rq.a00, s_ = (double)Rq.a10, n = sqrt(c * c + s_ * s_);
    const double dx = ox - px, dy = oy - py;
    const double cv = ((c / n) * dx + (s_ / n) * dy) / sqrt(dx * dx + dy * dy);
    vis = vis && cv >= -1.0;
  }
  __syncwarp();   // convex points consumed; the row is free for the observation staging
  S.visible = vis ? 1 : 0;
  // flag switch (CTG:573-581): the robot without the flag touches it
  const int wf_old = S.with_flag;
  const double ffx = S.flag_x, ffy = S.flag_y;
  S.sw = 0;
  if ((wf_old && touch_other) || (!wf_old && touch_own)) {
    S.with_flag = 1 - wf_old;
    S.sw = 1;
    double u[4];
    stream_uniforms(seed, pair_gid, epi, 4, (unsigned)S.flag_draws, u);
    S.flag_draws += 1;
    S.flag_x = -2.0 + 4.0 * u[0]; S.flag_y = -2.0 + 4.0 * u[1];
  }
  if (k == 0) {
    const V3 wl = tmul(Rq, ww), vl = tmul(Rq, vw);
    snew[24] = wl.x; snew[25] = wl.y; snew[26] = wl.z; snew[27] = vl.x; snew[28] = vl.y; snew[29] = vl.z;
    snew[30] = Rq.a20; snew[31] = Rq.a21; snew[32] = Rq.a22;
    snew[45] = Rq.a00; snew[46] = Rq.a01; snew[47] = Rq.a02; snew[48] = Rq.a10; snew[49] = Rq.a11; snew[50] = Rq.a12;
    snew[51] = Rq.a20; snew[52] = Rq.a21; snew[53] = Rq.a22;
    snew[54] = pos.x; snew[55] = pos.y; snew[56] = pos.z;
    snew[57] = (float)ffx; snew[58] = (float)ffy;                       // the flag where it stood during this step
    const float yaw = atan2f(Rq.a10, Rq.a00);
    snew[59] = yaw;
    float sy, cy;
    llq_sincosf(yaw, &sy, &cy);
    float* v = snew + 62;
    v[0] = pos.x; v[1] = pos.y; v[2] = pos.z; v[3] = cy; v[4] = sy;                          // percept_vec
    const M3 Ro = qmat(qnormalize(oq));
    const float yawo = atan2f(Ro.a10, Ro.a00);
    float sd, cd;
    llq_sincosf(yawo - yaw, &sd, &cd);
    const V3 dl = tmul(Rq, V3{(float)(ox - px), (float)(oy - py), (float)(oz - pz)}), ovl = tmul(Rq, ov), owl = tmul(Rq, oww);
    const float oppo[15] = {vis ? 1.f : 0.f, opos.x, opos.y, opos.z, dl.x, dl.y, dl.z, cd, sd, ovl.x, ovl.y, ovl.z, owl.x, owl.y, owl.z};
#pragma unroll
    for (int t = 0; t < 15; t++) { v[5 + t] = vis ? oppo[t] : 0.f; v[20 + t] = oppo[t]; }
    const V3 fl = tmul(Rq, V3{(float)(ffx - px), (float)(ffy - py), (float)(0.25 - pz)});
    const float fi[7] = {1.f, (float)ffx, (float)ffy, 0.25f, fl.x, fl.y, fl.z};
#pragma unroll
    for (int t = 0; t < 7; t++) { v[35 + t] = fi[t]; v[42 + t] = fi[t]; }
    v[49] = (float)S.with_flag; v[50] = (float)(1 - S.with_flag);                              // CTG:584, after a possible switch
    v[51] = fix_spd;
  }
}


// ---------------------------------------------------------------------------------------------------------------
// EPMC corridor (elements 1-3; BSE = max_game_elements/bullet_static_entities.py).
// 64-bit mask of the env's boxes whose xy footprint comes within `margin` of (px, py) (zsel: whose z range contains pz);
// the 4 lanes of an env scan interleaved quarters and combine.
LLQ_DI unsigned long long box_mask(const float* boxes, int nb, int k, float px, float py, float pz, float margin, bool zsel) {
  unsigned long long m = 0ull;
  for (int j = k; j < nb; j += 4) {
    const float* b = boxes + 6 * j;
    const bool hit = zsel ? fabsf(b[2] - pz) <= b[5] : (fabsf(b[0] - px) <= b[3] + margin && fabsf(b[1] - py) <= b[4] + margin);
    if (hit) m |= 1ull << j;
  }
  unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
  lo |= __shfl_xor_sync(FULL, lo, 1); hi |= __shfl_xor_sync(FULL, hi, 1);
  lo |= __shfl_xor_sync(FULL, lo, 2); hi |= __shfl_xor_sync(FULL, hi, 2);
  return ((unsigned long long)hi << 32) | lo;
}
// closest hit fraction against the ground slab and the boxes selected by `mask`
LLQ_DI float ray_boxlist(V3 o, V3 d, const float* boxes, unsigned long long mask) {
  float best = ray_box1(o, d, V3{-100.f, -100.f, -10.f}, V3{100.f, 100.f, 0.f}, -1.f);
  while (mask) {
    const int j = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    const float* b = boxes + 6 * j;
    best = ray_box1(o, d, V3{b[0] - b[3], b[1] - b[4], b[2] - b[5]}, V3{b[0] + b[3], b[1] + b[4], b[2] + b[5]}, best);
  }
  return best;
}
struct TerrainRng {
  unsigned long long seed; long long gid, ep; int k; double u[4];
  LLQ_DI double next() {
    if ((k & 3) == 0) stream_uniforms(seed, gid, ep, 5, (unsigned)(k >> 2), u);
    const double v = (k & 3) == 0 ? u[0] : ((k & 3) == 1 ? u[1] : ((k & 3) == 2 ? u[2] : u[3]));
    k++;
    return v;
  }
  LLQ_DI double uniform(double lo, double hi) { return lo + next() * (hi - lo); }
  LLQ_DI int randint(int lo, int hi) { return lo + (int)floor(next() * (double)(hi - lo)); }
};
LLQ_DI void put_box(float* boxes, int& nb, bool wr, double cx, double cy, double cz, double lx, double ly, double lz) {
  if (nb < kMaxBoxes && wr) {
    float* b = boxes + 6 * nb;
    b[0] = (float)cx; b[1] = (float)cy; b[2] = (float)cz; b[3] = (float)(lx / 2); b[4] = (float)(ly / 2); b[5] = (float)(lz / 2);
  }
  if (nb < kMaxBoxes) nb++;
}
// reset(): _generate_random_width_walls + _create_hurdles / _create_holes / _create_cubes(easy) (BSE:170-263, 308-500); returns the
// number of boxes, writes them when `wr`, and the target x (target y = 0)
LLQ_DI int generate_corridor(const StepParams& P, unsigned long long seed, long long gid, long long ep, float* boxes, bool wr, double& tgx) {
  TerrainRng R{seed, gid, ep, 0, {0.0, 0.0, 0.0, 0.0}};
  int nb = 0;
  const double width = R.uniform((double)P.ww_lo, (double)P.ww_hi), gap = R.uniform((double)P.wg_lo, (double)P.wg_hi);
  put_box(boxes, nb, wr, 5.0, gap / 2.0 + width / 2.0, 1.0, 200.0, width, 2.0);
  put_box(boxes, nb, wr, 5.0, -(gap / 2.0 + width / 2.0), 1.0, 200.0, width, 2.0);
  double cur = 0.0;
  tgx = 8.0;
  if (P.element_id == 1 || P.element_id == 2) {
    const int n = R.randint(1, 10);
    for (int pass = 0; pass < 2; pass++) {
      for (int i = 0; i < n; i++) {
        if (P.element_id == 1) {
          const double h = R.uniform(0.05, 0.15), d = R.uniform(1.0, 3.0);
          put_box(boxes, nb, wr, cur + d / 2, 0.0, h / 2, 0.1, gap, h);
          cur += d + 0.1;
        } else {
          const double d = R.uniform(1.0, 3.0), g = R.uniform((double)P.hg_lo, (double)P.hg_hi);
          put_box(boxes, nb, wr, cur + d / 2, 0.0, 0.3 / 2 + g, 0.1, gap, 0.3);
          cur += d + 0.1;
        }
      }
      if (pass == 0) tgx = cur + R.uniform(-1.0, 1.0);
    }
  } else {
    const int ns = R.randint(1, 5);
    for (int pass = 0; pass < 2; pass++) {
      for (int i = 0; i < ns; i++) {
        cur += R.uniform(0.0, 1.0);
        put_box(boxes, nb, wr, 1.75 + cur, 0.0, 0.25 / 2, 0.5, gap, 0.25);
        put_box(boxes, nb, wr, 1.0 + cur, 0.0, 0.1 / 2, 0.5, gap, 0.1);
        cur += 1.75 + 0.25;
        put_box(boxes, nb, wr, cur + 0.5, 0.0, 0.25 / 2, 0.5, gap, 0.25);
        put_box(boxes, nb, wr, cur + 1.25, 0.0, 0.1 / 2, 0.5, gap, 0.1);
        cur += 3.0;
      }
      if (pass == 0) tgx = cur + R.uniform(-3.0, 3.0);
    }
  }
  return nb;
}
// stage the perception context of an EPMC-corridor row: yaw and the three candidate masks (as raw bits)
LLQ_DI void stage_corridor_masks(float* snew, const float* boxes, int nb, int k, float px, float py, float pz, float yaw) {
  const unsigned long long m2 = box_mask(boxes, nb, k, px, py, pz, 1.36f, false);   // 2.4 x 1.2 footprint, any yaw
  const unsigned long long mf = box_mask(boxes, nb, k, px, py, pz, 3.35f, false);   // 3 m rays starting up to 0.27 m off the base
  const unsigned long long m1 = box_mask(boxes, nb, k, px, py, pz, 0.f, true);      // horizontal rays at the base height
  if (k == 0) {
    snew[61] = yaw;
    snew[62] = __uint_as_float((unsigned)m2); snew[63] = __uint_as_float((unsigned)(m2 >> 32));
    snew[64] = __uint_as_float((unsigned)mf); snew[65] = __uint_as_float((unsigned)(mf >> 32));
    snew[66] = __uint_as_float((unsigned)m1); snew[67] = __uint_as_float((unsigned)(m1 >> 32));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Shared tail: given the dynamic robot state (pybullet convention) and the mocap cursor, build the new prop / future
// into the staging row `snew` (120 floats per env) and return the pieces the reward needs.
struct ObsCtx {
  KinBase kb;          // kinematic (mocap) base
  float kq[3], kqd[3]; // kinematic joints of this lane's leg
};

LLQ_DI ObsCtx build_obs_new(const MocapDev& mc, const StepParams& P, const ModelConst& M, int lane4, int clip, int frame_id,
                            double frac, double px, double py, double pz, Q4 qb, V3 lin, V3 ang, const float (&q)[3],
                            const float (&qd)[3], float* snew) {
  ObsCtx o;
  const MocapFrame* f0 = mc.frames + mc.clip_off[clip] + frame_id;
  o.kb = mocap_base(f0, f0 + 1, frac, P.frame_dt);
  float inv = (float)(1.0 / P.frame_dt), fr = (float)frac;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float c = f0->q[3 * lane4 + i], n = f0[1].q[3 * lane4 + i];
    o.kq[i] = fmaf(fr, n - c, c);
    o.kqd[i] = (n - c) * inv;
  }
  qb = qnormalize(qb);
  M3 Rb = qmat(qb);
  // prop (PLE:247-260): joint_pos | joint_vel | R^T w | R^T v | R[2,:]
#pragma unroll
  for (int i = 0; i < 3; i++) { snew[3 * lane4 + i] = q[i]; snew[12 + 3 * lane4 + i] = qd[i]; }
  if (lane4 == 0) {
    V3 wl = tmul(Rb, ang), vl = tmul(Rb, lin);
    snew[24] = wl.x; snew[25] = wl.y; snew[26] = wl.z;
    snew[27] = vl.x; snew[28] = vl.y; snew[29] = vl.z;
    snew[30] = Rb.a20; snew[31] = Rb.a21; snew[32] = Rb.a22;
  }
  // future target `lane4` (ML:75-86, PLE:299-317)
  {
    const double tf = lane4 == 0 ? 1. / 30. : (lane4 == 1 ? 1. / 15. : (lane4 == 2 ? 1. / 3. : 1.));
    double t = P.frame_dt * frac + tf;
    int fid = (int)floor(t / P.frame_dt);
    double ffrac = t / P.frame_dt - fid;
    const MocapFrame* g0 = f0 + fid;
    KinBase kf = mocap_base(g0, g0 + 1, ffrac, P.frame_dt);
    V3 dp = tmul(Rb, V3{(float)(kf.px - px), (float)(kf.py - py), (float)(kf.pz - pz)});
    V3 rv = q_rotvec(qnormalize(qmul(qconj(qb), qnormalize(kf.q))));
    float angle = norm3(rv);
    float sc = angle / (angle + 1e-8f);
    float* o18 = snew + 45 + 18 * lane4;
    o18[0] = dp.x; o18[1] = dp.y; o18[2] = dp.z;
    o18[3] = sc * rv.x; o18[4] = sc * rv.y; o18[5] = sc * rv.z;
    float ff = (float)ffrac;
#pragma unroll
    for (int j = 0; j < 12; j++) { float c = g0->q[j], n = g0[1].q[j]; o18[6 + j] = fmaf(ff, n - c, c); }
  }
  return o;
}

// Cooperative, coalesced emission of the 8 observation rows owned by this warp.
// mode 0 (step):  prop = [old[33:99], new] ; prop_a = [old[12:36], act] ; future = new
// mode 1 (reset): prop = [new, new, new]  ; prop_a = 0                 ; future = new      (PLE:282-290)
// `do_row` (bit e of a warp-uniform mask) selects which of the 8 rows are written.
constexpr int kHist = 90;   // per-env history carry: prop[33:99] (66) | prop_a[12:36] (24)

// staging row (kNewObs floats per env).  PMC: prop 33 | action 12 | future 72.
// EPMC: prop 33 | action 12 | R (world<-base inertial, row major) 9 | pos 3 | target 3 | |base_pos| 1   (perception is evaluated while the row is written)
template <int ENV, int EPW = 8>   // EPW = envs per warp (8 with 4 lanes per env, 2 with 16)
LLQ_DI void emit_obs_rows(float* obs, float* obs2, long long obs2_ld, const float* snew_warp, const float* hist_warp, int env0, int n_envs,
                          int mode, unsigned row_mask, const float* boxes_all = nullptr) {
  constexpr int OW = ObsW<ENV>::value;
  const int lane = threadIdx.x & 31;
#pragma unroll 4
  for (int base = 0; base < EPW * OW; base += 32) {
    int idx = base + lane;
    int e = idx / OW, j = idx - e * OW;
    bool ok = idx < EPW * OW && (env0 + e) < n_envs && ((row_mask >> e) & 1u);
    float v = 0.f;
    if (ok) {
      const float* sn = snew_warp + e * kNewObs;
      const float* hs = hist_warp + e * kHist;
      if (j < 99) {
        if (mode == 1) v = sn[j % kPropDim];
        else v = j < 66 ? hs[j] : sn[j - 66];
      } else if (j < 135) {
        int a = j - 99;
        if (mode == 1) v = 0.f;
        else v = a < 24 ? hs[66 + a] : sn[kPropDim + a - 24];
      } else if (ENV == 0) {
        v = sn[45 + (j - 135)];
      } else if (ENV == 3) {
        // EPMC corridor perception against the ground slab and the env's candidate boxes (PGE:374-447)
        const V3 pos = V3{sn[54], sn[55], sn[56]};
        const float* bxs = boxes_all + (size_t)(env0 + e) * (6 * kMaxBoxes);
        if (j < 460) {
          const unsigned long long m = ((unsigned long long)__float_as_uint(sn[63]) << 32) | __float_as_uint(sn[62]);
          const int t = j - 135, a = t / 13, b = t - a * 13;
          const float gx = a == 24 ? 1.2f : -1.2f + (float)a * (2.4f / 24.0f), gy = b == 12 ? 0.6f : -0.6f + (float)b * (1.2f / 12.0f);
          const float x = fmaf(sn[45], gx, fmaf(sn[46], gy, pos.x)), y = fmaf(sn[48], gx, fmaf(sn[49], gy, pos.y));
          const float f = ray_boxlist(V3{x, y, 10.f}, V3{0.f, 0.f, -20.f}, bxs, m);
          v = f < 0.f ? 0.f : fmaf(f, -20.f, 10.f);
          if (f >= 0.f && fabsf(v) < 2e-6f) v = 0.f;
        } else if (j < 588) {
          const unsigned long long m = ((unsigned long long)__float_as_uint(sn[67]) << 32) | __float_as_uint(sn[66]);
          const float ang = sn[61] + 6.283185307179586f * (float)(j - 460) * (1.0f / 128.0f);
          float sa, ca;
          llq_sincosf(ang, &sa, &ca);
          const float f = ray_boxlist(pos, V3{20.f * ca, 20.f * sa, 0.f}, bxs, m);
          v = f < 0.f ? sn[60] : f * 20.f * sqrtf(ca * ca + sa * sa);
        } else if (j < 913) {
          const unsigned long long m = ((unsigned long long)__float_as_uint(sn[65]) << 32) | __float_as_uint(sn[64]);
          const int t = j - 588, a = t / 13, b = t - a * 13;
          const float y = a == 24 ? 0.25f : -0.25f + (float)a * (0.5f / 24.0f), z = b == 12 ? 0.1f : -0.3f + (float)b * (0.4f / 12.0f);
          const V3 from = V3{fmaf(sn[46], y, fmaf(sn[47], z, pos.x)), fmaf(sn[49], y, fmaf(sn[50], z, pos.y)), fmaf(sn[52], y, fmaf(sn[53], z, pos.z))};
          const V3 d = V3{3.f * sn[45], 3.f * sn[48], 3.f * sn[51]};
          const float f = ray_boxlist(from, d, bxs, m);
          v = (f < 0.f ? 1.f : f) * norm3(d);
        } else {
          v = sn[57 + (j - 913)];
        }
      } else if (ENV == 2) {
        // SEPMC perception against ground slab, walls and flag (CTG:598-638, PGE:22-54)
        const V3 pos = V3{sn[54], sn[55], sn[56]};
        const float fx = sn[57], fy = sn[58];
        if (j < 460) {                             // percept_2d: down rays over the 25 x 13 grid in the full base frame, value = hit z
          const int t = j - 135, a = t / 13, b = t - a * 13;
          const float gx = a == 24 ? 1.2f : -1.2f + (float)a * (2.4f / 24.0f), gy = b == 12 ? 0.6f : -0.6f + (float)b * (1.2f / 12.0f);
          const float x = fmaf(sn[45], gx, fmaf(sn[46], gy, pos.x)), y = fmaf(sn[48], gx, fmaf(sn[49], gy, pos.y));
          // a vertical ray sees the highest top among the boxes whose footprint holds (x, y): flag 0.5, walls 2, ground 0
          const bool in_x = fabsf(x) <= 2.5f, in_y = fabsf(y) <= 2.5f;
          const bool wall = (in_x && fabsf(fabsf(y) - 2.5f) <= 0.005f) || (in_y && fabsf(fabsf(x) - 2.5f) <= 0.005f);
          const bool flag = fabsf(x - fx) <= 0.05f && fabsf(y - fy) <= 0.05f;
          v = wall ? 2.0f : (flag ? 0.5f : 0.0f);
          if (!(fabsf(x) < 99.f && fabsf(y) < 99.f)) {                    // off the slab: the general test decides
            const float f = ray_arena(V3{x, y, 10.f}, V3{0.f, 0.f, -20.f}, fx, fy);
            v = f < 0.f ? 0.f : fmaf(f, -20.f, 10.f);
          }
        } else if (j < 588) {                      // percept_1d: 128 horizontal rays of 20 m; a miss reports |ray_from|
          const float ang = sn[59] + 6.283185307179586f * (float)(j - 460) * (1.0f / 128.0f);
          float sa, ca;
          llq_sincosf(ang, &sa, &ca);
          const V3 d = V3{20.f * ca, 20.f * sa, 0.f};
          const bool inside = fabsf(pos.x) < 2.49f && fabsf(pos.y) < 2.49f;
          const float f = inside ? ray_arena_inside(pos, d, fx, fy) : ray_arena(pos, d, fx, fy);
          v = f < 0.f ? norm3(pos) : f * 20.f * sqrtf(ca * ca + sa * sa);
        } else if (j < 913) {                      // percept_front: 25 x 13 rays of 3 m along body +x; a miss reports 3
          const int t = j - 588, a = t / 13, b = t - a * 13;
          const float y = a == 24 ? 0.25f : -0.25f + (float)a * (0.5f / 24.0f), z = b == 12 ? 0.1f : -0.3f + (float)b * (0.4f / 12.0f);
          const V3 from = V3{fmaf(sn[46], y, fmaf(sn[47], z, pos.x)), fmaf(sn[49], y, fmaf(sn[50], z, pos.y)), fmaf(sn[52], y, fmaf(sn[53], z, pos.z))};
          const V3 d = V3{3.f * sn[45], 3.f * sn[48], 3.f * sn[51]};
          const bool inside = fabsf(from.x) < 2.49f && fabsf(from.y) < 2.49f && from.z > 0.f;
          const float f = inside ? ray_arena_inside(from, d, fx, fy) : ray_arena(from, d, fx, fy);
          v = (f < 0.f ? 1.f : f) * norm3(d);
        } else {
          v = sn[62 + (j - 913)];
        }
      } else if (j < 460) {
        v = 0.f;                                   // percep_2d: every down-ray hits the slab top, hit z = 0 (PGE:431-447)
      } else if (j < 588) {
        v = sn[60];                                // percep_1d: horizontal rays miss => |ray_from| (PGE:49-53,388-394)
      } else if (j < 913) {                        // percep_front (PGE:409-429) against the ground slab
        int t = j - 588, i = t / 13, jj = t - i * 13;
        float y = i == 24 ? 0.25f : -0.25f + (float)i * (0.5f / 24.0f);
        float z = jj == 12 ? 0.1f : -0.3f + (float)jj * (0.4f / 12.0f);
        float fz = fmaf(sn[52], y, fmaf(sn[53], z, sn[56]));          // from.z = R[2,1] y + R[2,2] z + pos.z
        float dz = 3.0f * sn[51];                                     // (to - from).z = 3 R[2,0]
        float len = 3.0f * sqrtf(sn[45] * sn[45] + sn[48] * sn[48] + sn[51] * sn[51]);
        float tz = fz + dz;
        v = (fz > 0.f && tz < 0.f) ? len * (fz / (fz - tz)) : len;
      } else {
        v = sn[57 + (j - 913)];
      }
      obs[(size_t)(env0 + e) * OW + j] = v;
      if (obs2) obs2[(size_t)(env0 + e) * obs2_ld + j] = v;
    }
  }
}

// Asynchronous (cp.async) prefetch issued at kernel start; consumed after the ten sub-steps, so DRAM latency is hidden.
template <int ENV, int EPW = 8>
LLQ_DI void prefetch_history(const float* obs, float* hist_warp, int env0, int n_envs) {
  constexpr int OW = ObsW<ENV>::value;
  const int lane = threadIdx.x & 31;
  for (int idx = lane; idx < EPW * kHist; idx += 32) {
    int e = idx / kHist, t = idx - e * kHist;
    int env = env0 + e < n_envs ? env0 + e : n_envs - 1;
    int j = t < 66 ? 33 + t : 99 + 12 + (t - 66);
    __pipeline_memcpy_async(hist_warp + idx, obs + (size_t)env * OW + j, 4);
  }
}
LLQ_DI void prefetch_model(const ModelConst* gmodel, ModelConst* smodel, int nthreads) {
  static_assert(sizeof(ModelConst) % 16 == 0, "ModelConst must be a multiple of 16 bytes");
  const float4* src = reinterpret_cast<const float4*>(gmodel);
  float4* dst = reinterpret_cast<float4*>(smodel);
  for (int i = threadIdx.x; i < (int)(sizeof(ModelConst) / 16); i += nthreads) __pipeline_memcpy_async(dst + i, src + i, 16);
}

// ---------------------------------------------------------------------------------------------------------------
// The fused policy-step kernel.
template <int BLOCK, int ENV>
__global__ void __launch_bounds__(BLOCK) pmc_step_kernel(EnvArrays E, MocapDev mc, StepParams P, const ModelConst* __restrict__ gmodel,
                                                         const float* __restrict__ actions, float* obs2, long long obs2_ld,
                                                         int* __restrict__ winner, unsigned long long seed, long long gid0, int record) {
  __shared__ __align__(16) ModelConst M;
  __shared__ __align__(16) float s_new[BLOCK / 4][kNewObs];
  __shared__ __align__(16) float s_hist[BLOCK / 4][kHist];
  extern __shared__ float rows_sm[];     // kRowFloats * BLOCK floats: per-lane constraint-row workspace
  const int tid = threadIdx.x;
  const int N = P.n_envs;
  prefetch_model(gmodel, &M, BLOCK);
  __pipeline_commit();
  prefetch_history<ENV>(E.obs, &s_hist[(tid & ~31) >> 2][0], (blockIdx.x * BLOCK + (tid & ~31)) >> 2, N);
  __pipeline_commit();
  __pipeline_wait_prior(1);              // model constants have landed; the history copy stays in flight
  __syncthreads();

  const int gtid = blockIdx.x * BLOCK + threadIdx.x;
  const int env_raw = gtid >> 2;
  const int env = env_raw < N ? env_raw : N - 1;   // surplus lanes shadow the last env (they must join the shuffles)
  const bool valid = env_raw < N;
  const int k = threadIdx.x & 3;                   // leg
  const LegConst& L = M.leg[k];
  const V3 r[3] = {ld3(L.j[0].r), ld3(L.j[1].r), ld3(L.j[2].r)};

  // ---- load state (SoA, coalesced over envs; base entries are broadcast within the 4 lanes)
  double px = E.pos[env], py = E.pos[N + env], pz = E.pos[2 * N + env];
  const float* st = E.st;
  Q4 qb = Q4{st[env], st[N + env], st[2 * N + env], st[3 * N + env]};
  V3 vw = V3{st[4 * N + env], st[5 * N + env], st[6 * N + env]};
  V3 ww = V3{st[7 * N + env], st[8 * N + env], st[9 * N + env]};
  float q[3], qd[3], act[3], tgt[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    q[i] = st[(10 + 3 * k + i) * N + env];
    qd[i] = st[(22 + 3 * k + i) * N + env];
    act[i] = actions[(size_t)env * kActDim + 3 * k + i];
    tgt[i] = clampf(q[i] + act[i], -3.0f, 3.0f);           // PLE:200, LR:126-127
  }
  float warm = E.warm[k * N + env];               // remembered normal impulse of this leg's contact; negative: it is the knee wheel's
  bool warm_wheel = warm < 0.f;
  warm = fabsf(warm);
  double time = E.time[env];
  const int clip = ENV == 0 ? E.clip[env] : 0;
  int frame_id = 0; double frame_frac = 0.0;
  int ob_id = 0; bool ob_hit = false;
  if (ENV == 0 && P.has_ob) ob_id = E.ob_id[env];
  // ---- EPMC bookkeeping (replicated on the 4 lanes): joystick command, push randomiser, per-episode friction
  int counter = 0, cmd_freq = 1, push_count = 0, push_draws = 0, cmd_draws = 0;
  double tgx = 0.0, tgy = 0.0, total_spd = 0.0, max_spd = 0.0, target_angle = 0.0, last_len = 0.0;
  float target_spd = 0.f, pf[3] = {0.f, 0.f, 0.f}, mu_env = P.mu;
  long long epi = 0;
  // ---- SEPMC bookkeeping (pair state replicated on both robots): CTG / PR
  PairState PS = {0, 0, 1, 0, 0.0, 0.0};
  float fix_spd = 0.f;
  bool touch_own = false, tag = false;
  const int robot = env & 1;
  const long long pair_gid = gid0 + (env & ~1);
  if (ENV == 2) {
    const double* A = E.aux;
    counter = (int)A[env]; PS.with_flag = (int)A[N + env]; PS.flag_x = A[2 * N + env]; PS.flag_y = A[3 * N + env];
    fix_spd = (float)A[4 * N + env]; total_spd = A[7 * N + env]; max_spd = A[8 * N + env]; push_count = (int)A[9 * N + env];
    pf[0] = (float)A[10 * N + env]; pf[1] = (float)A[11 * N + env]; pf[2] = (float)A[12 * N + env];
    mu_env = P.mu_ground * (float)A[13 * N + env]; push_draws = (int)A[14 * N + env]; PS.flag_draws = (int)A[15 * N + env];
    epi = E.episode[env] - 1;
  }
  double init_len = 1.0;
  if (ENV == 1 || ENV == 3) {
    const double* A = E.aux;
    if (ENV == 3) init_len = A[17 * N + env];
    counter = (int)A[env]; cmd_freq = (int)A[N + env]; tgx = A[2 * N + env]; tgy = A[3 * N + env];
    target_spd = (float)A[4 * N + env]; target_angle = A[5 * N + env]; last_len = A[6 * N + env]; total_spd = A[7 * N + env];
    max_spd = A[8 * N + env]; push_count = (int)A[9 * N + env]; pf[0] = (float)A[10 * N + env]; pf[1] = (float)A[11 * N + env];
    pf[2] = (float)A[12 * N + env]; mu_env = P.mu_ground * (float)A[13 * N + env]; push_draws = (int)A[14 * N + env];
    cmd_draws = (int)A[15 * N + env];
    epi = E.episode[env] - 1;                             // streams of the running episode (the reset advanced the counter)
    if (counter % cmd_freq == 0) {                        // PGE:302-317, element_id 0
      double u[4];
      stream_uniforms(seed, gid0 + env, epi, 3, (unsigned)cmd_draws++, u);
      if (ENV == 1) {
        target_angle = 2.0 * 3.14159265358979323846 * u[0];
        double sn, cs;
        sincos(target_angle, &sn, &cs);
        tgx = px + cs * 100.0; tgy = py + sn * 100.0;
        last_len = sqrt((px - tgx) * (px - tgx) + (py - tgy) * (py - tgy));
      }
      target_spd = (float)((double)P.ts_lo + u[1] * ((double)P.ts_hi - (double)P.ts_lo));
    }
    if (ENV == 3) target_angle = atan2(tgy - py, tgx - px);            // PGE:318-323 (plotting only)
  }
  // ---- EPMC corridor: the boxes the feet can reach during this step -> shared memory (<= kMaxCand per env)
  int n_cand = 0;
  float* s_cand = nullptr;
  if (ENV == 3) {
    s_cand = &s_new[threadIdx.x >> 2][0];              // the staging row is free until the tail: 8 x 6 floats
    const float* bxs = E.boxes + (size_t)env * (6 * kMaxBoxes);
    unsigned long long m = box_mask(bxs, E.nbox[env], k, (float)px, (float)py, (float)pz, 0.6f, false);
    int c = 0;
    while (m && c < kMaxCand) {
      const int j = __ffsll((long long)m) - 1;
      m &= m - 1;
      if ((c & 3) == k) {
#pragma unroll
        for (int t = 0; t < 6; t++) s_cand[6 * c + t] = bxs[6 * j + t];
      }
      c++;
    }
    n_cand = c;
    __syncwarp();
  }
  // base orientation: pybullet speaks in the base inertial frame; dynamics run in URDF body axes B' = inertial * qI^-1
  const Q4 qI = Q4{M.base.qI[0], M.base.qI[1], M.base.qI[2], M.base.qI[3]};
  Q4 qp = qmul(qnormalize(qb), qconj(qI));
  unsigned long long n_contact_rows = 0, n_limit_rows = 0;
  bool bad = false;

  const V3 bh = ld3(M.base.h); const Sym3 bI = ldsym(M.base.I); const float bm = M.base.m;

  for (int sub = 0; sub < P.substeps; sub++) {
    // keep the warps of a CTA on the same stretch of code so that they share instruction-cache lines (the sub-step body
    // is far larger than the SM's instruction cache)
    if (BLOCK > 32) __syncthreads();
    const float dt = P.dt;
    // ---------------- PD actuator (LR:138-141) + joint damping (pybullet applyJointDamping)
    float tau[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      float t = fmaf(P.kp, tgt[i] - q[i], P.kd * (0.f - qd[i]));
      tau[i] = clampf(t, -P.max_tau, P.max_tau) - L.j[i].jdamp * qd[i];
    }
    // ---------------- EPMC push randomiser (PR:56-87): counters in sub-steps, force lasts one sub-step
    bool push_on = false;
    if (ENV == 2 && P.push_enabled) {
      // two robots (PR:79-87): inside the window each robot gets a freshly randomised force every sub-step: robot 0 the current
      // draw, robot 1 the next one, and one more draw is consumed
      push_count += 1;
      if (push_count > 0) {
        if (push_count % P.push_interval == 0) { push_draws += 1; push_count = 0; }
        if (push_count < P.push_duration) {
          push_force_of_draw(P, seed, pair_gid, epi, push_draws - 1 + robot, pf);
          push_draws += 2;
          push_on = true;
        }
      }
    }
    if ((ENV == 1 || ENV == 3) && P.push_enabled) {
      push_count += 1;
      if (push_count > 0) {
        if (push_count % P.push_interval == 0) { epmc_randomize_push(P, seed, gid0 + env, epi, push_draws, pf); push_count = 0; }
        push_on = push_count < P.push_duration;
      }
    }
    // ---------------- kinematics
    const M3 R = qmat(qp);                       // world <- B'
    JC jc[3];
    llq_sincosf(q[0], &jc[0].s, &jc[0].c);
    llq_sincosf(-q[1], &jc[1].s, &jc[1].c);
    llq_sincosf(-q[2], &jc[2].s, &jc[2].c);
    // ---------------- ABA pass 1: velocities, velocity products, bias forces (link coords, link origins)
    SV v0; v0.a = tmul(R, ww); v0.l = tmul(R, vw);
    SV v1 = xmotion<0>(v0, r[0], jc[0].c, jc[0].s);
    SV vj = SV{V3{qd[0], 0.f, 0.f}, V3{0.f, 0.f, 0.f}};
    v1.a.x += qd[0];
    SV c1 = SV{cross(v1.a, vj.a), cross(v1.l, vj.a)};
    SV v2 = xmotion<1>(v1, r[1], jc[1].c, jc[1].s);
    vj.a = V3{0.f, -qd[1], 0.f};
    v2.a.y -= qd[1];
    SV c2 = SV{cross(v2.a, vj.a), cross(v2.l, vj.a)};
    SV v3 = xmotion<1>(v2, r[2], jc[2].c, jc[2].s);
    vj.a = V3{0.f, -qd[2], 0.f};
    v3.a.y -= qd[2];
    SV c3 = SV{cross(v3.a, vj.a), cross(v3.l, vj.a)};
#if LLQ_BARRIERS >= 5
    if (BLOCK > 32) __syncthreads();
#endif
    // ---------------- ABA pass 2: articulated inertia, leaf -> root inside the lane
    ABI IA = rigid_abi(L.j[2].m, ld3(L.j[2].h), ldsym(L.j[2].I));
    SV pA = bias_force<2>(L.j[2].m, ld3(L.j[2].h), ldsym(L.j[2].I), L.j[2].nd, L.j[2].d, v3.a, v3.l, P.kl, P.ka);
    joint_reduce<1, -1>(IA, pA, c3, tau[2], r[2], jc[2]);
    {
      ABI I2 = rigid_abi(L.j[1].m, ld3(L.j[1].h), ldsym(L.j[1].I));
      SV p2 = bias_force<2>(L.j[1].m, ld3(L.j[1].h), ldsym(L.j[1].I), L.j[1].nd, L.j[1].d, v2.a, v2.l, P.kl, P.ka);
      IA.A = IA.A + I2.A; IA.B = IA.B + I2.B; IA.C = IA.C + I2.C; pA.a = pA.a + p2.a; pA.l = pA.l + p2.l;
    }
    joint_reduce<1, -1>(IA, pA, c2, tau[1], r[1], jc[1]);
    {
      ABI I1 = rigid_abi(L.j[0].m, ld3(L.j[0].h), ldsym(L.j[0].I));
      SV p1 = bias_force<2>(L.j[0].m, ld3(L.j[0].h), ldsym(L.j[0].I), L.j[0].nd, L.j[0].d, v1.a, v1.l, P.kl, P.ka);
      if (ENV != 0 && push_on && k == 0) {
        // applyExternalForce(link 0 = FR hip, LINK_FRAME): force given in the hip's inertial frame, applied at its CoM
        const V3 fl = V3{M.push_R[0] * pf[0] + M.push_R[1] * pf[1] + M.push_R[2] * pf[2], M.push_R[3] * pf[0] + M.push_R[4] * pf[1] + M.push_R[5] * pf[2],
                         M.push_R[6] * pf[0] + M.push_R[7] * pf[1] + M.push_R[8] * pf[2]};
        p1.a = p1.a - cross(ld3(M.push_c), fl);
        p1.l = p1.l - fl;
      }
      IA.A = IA.A + I1.A; IA.B = IA.B + I1.B; IA.C = IA.C + I1.C; pA.a = pA.a + p1.a; pA.l = pA.l + p1.l;
    }
    joint_reduce<0, 1>(IA, pA, c1, tau[0], r[0], jc[0]);
#if LLQ_BARRIERS >= 3
    if (BLOCK > 32) __syncthreads();
#endif
    // ---------------- base: sum the four legs (xor shuffles), add the base body, factorise
    float m6[21], z0[6];
    {
      // packed lower triangle of [[A, B], [B^T, C]] : rows 0-2 = A, rows 3-5 = [B^T, C]
      m6[tri(0, 0)] = IA.A.xx; m6[tri(1, 0)] = IA.A.xy; m6[tri(1, 1)] = IA.A.yy;
      m6[tri(2, 0)] = IA.A.xz; m6[tri(2, 1)] = IA.A.yz; m6[tri(2, 2)] = IA.A.zz;
      m6[tri(3, 0)] = IA.B.a00; m6[tri(3, 1)] = IA.B.a10; m6[tri(3, 2)] = IA.B.a20;
      m6[tri(4, 0)] = IA.B.a01; m6[tri(4, 1)] = IA.B.a11; m6[tri(4, 2)] = IA.B.a21;
      m6[tri(5, 0)] = IA.B.a02; m6[tri(5, 1)] = IA.B.a12; m6[tri(5, 2)] = IA.B.a22;
      m6[tri(3, 3)] = IA.C.xx; m6[tri(4, 3)] = IA.C.xy; m6[tri(4, 4)] = IA.C.yy;
      m6[tri(5, 3)] = IA.C.xz; m6[tri(5, 4)] = IA.C.yz; m6[tri(5, 5)] = IA.C.zz;
      z0[0] = pA.a.x; z0[1] = pA.a.y; z0[2] = pA.a.z; z0[3] = pA.l.x; z0[4] = pA.l.y; z0[5] = pA.l.z;
#pragma unroll
      for (int i = 0; i < 21; i++) m6[i] = gsum4(m6[i]);
#pragma unroll
      for (int i = 0; i < 6; i++) z0[i] = gsum4(z0[i]);
      SV pb = bias_force<3>(bm, bh, bI, M.base.nd, M.base.d, v0.a, v0.l, P.kl, P.ka);
      M3 hx = skew(bh);
      m6[tri(0, 0)] += bI.xx; m6[tri(1, 0)] += bI.xy; m6[tri(1, 1)] += bI.yy;
      m6[tri(2, 0)] += bI.xz; m6[tri(2, 1)] += bI.yz; m6[tri(2, 2)] += bI.zz;
      m6[tri(3, 0)] += hx.a00; m6[tri(3, 1)] += hx.a10; m6[tri(3, 2)] += hx.a20;
      m6[tri(4, 0)] += hx.a01; m6[tri(4, 1)] += hx.a11; m6[tri(4, 2)] += hx.a21;
      m6[tri(5, 0)] += hx.a02; m6[tri(5, 1)] += hx.a12; m6[tri(5, 2)] += hx.a22;
      m6[tri(3, 3)] += bm; m6[tri(4, 4)] += bm; m6[tri(5, 5)] += bm;
      z0[0] += pb.a.x; z0[1] += pb.a.y; z0[2] += pb.a.z; z0[3] += pb.l.x; z0[4] += pb.l.y; z0[5] += pb.l.z;
    }
    const Chol6 ch = chol6(m6);
    float a0[6];
    {
      float b[6];
#pragma unroll
      for (int i = 0; i < 6; i++) b[i] = -z0[i];
      chol6_solve(ch, b, a0);                   // acceleration relative to free fall (gravity as a fictitious base acceleration)
    }
    // ---------------- ABA pass 3
    float qdd[3];
    {
      SV a = xmotion<0>(SV{V3{a0[0], a0[1], a0[2]}, V3{a0[3], a0[4], a0[5]}}, r[0], jc[0].c, jc[0].s);
      a.a = a.a + c1.a; a.l = a.l + c1.l;
      qdd[0] = (jc[0].u - dot(jc[0].Ua, a.a) - dot(jc[0].Ul, a.l)) * jc[0].Dinv;
      a.a.x += qdd[0];
      a = xmotion<1>(a, r[1], jc[1].c, jc[1].s);
      a.a = a.a + c2.a; a.l = a.l + c2.l;
      qdd[1] = (jc[1].u - dot(jc[1].Ua, a.a) - dot(jc[1].Ul, a.l)) * jc[1].Dinv;
      a.a.y -= qdd[1];
      a = xmotion<1>(a, r[2], jc[2].c, jc[2].s);
      a.a = a.a + c3.a; a.l = a.l + c3.l;
      qdd[2] = (jc[2].u - dot(jc[2].Ua, a.a) - dot(jc[2].Ul, a.l)) * jc[2].Dinv;
    }
    // ---------------- velocity prediction  v* = clamp(v + a dt)   (btMultiBody::applyDeltaVeeMultiDof)
    {
      V3 wd = mul(R, V3{a0[0], a0[1], a0[2]});
      V3 vd = mul(R, V3{a0[3], a0[4], a0[5]} + cross(v0.a, v0.l));
      vd.z += P.gz;
      ww = V3{clampf(fmaf(wd.x, dt, ww.x), -P.vmax, P.vmax), clampf(fmaf(wd.y, dt, ww.y), -P.vmax, P.vmax), clampf(fmaf(wd.z, dt, ww.z), -P.vmax, P.vmax)};
      vw = V3{clampf(fmaf(vd.x, dt, vw.x), -P.vmax, P.vmax), clampf(fmaf(vd.y, dt, vw.y), -P.vmax, P.vmax), clampf(fmaf(vd.z, dt, vw.z), -P.vmax, P.vmax)};
#pragma unroll
      for (int i = 0; i < 3; i++) qd[i] = clampf(fmaf(qdd[i], dt, qd[i]), -P.vmax, P.vmax);
    }
    // predicted velocity in base coordinates (generalised velocity used by the constraint rows)
    const V3 wbs = tmul(R, ww), vbs = tmul(R, vw);

#if LLQ_BARRIERS >= 2
    if (BLOCK > 32) __syncthreads();
#endif
    // ---------------- leg kinematics and the ABA's per-joint vectors, re-expressed in base coordinates about the base origin
    const float kc1 = jc[0].c, ks1 = jc[0].s, kc2 = jc[1].c, ks2 = jc[1].s;
    const float kc23 = kc2 * jc[2].c - ks2 * jc[2].s, ks23 = ks2 * jc[2].c + kc2 * jc[2].s;
    const V3 p1 = r[0];
    const V3 p2 = p1 + rot<0>(r[1], kc1, ks1);
    const V3 p3 = p2 + rot<0>(rot<1>(r[2], kc2, ks2), kc1, ks1);
    const V3 fb = p3 + rot<0>(rot<1>(ld3(L.foot), kc23, ks23), kc1, ks1);     // foot centre
    const V3 n2 = V3{0.f, -kc1, -ks1};                                       // axis of joints 2, 3 (= -E1 e_y)
    V3 Sa[3], Sl[3], Ua[3], Ul[3];
    Sa[0] = V3{1.f, 0.f, 0.f}; Sl[0] = cross(p1, Sa[0]);
    Sa[1] = n2; Sl[1] = cross(p2, n2);
    Sa[2] = n2; Sl[2] = cross(p3, n2);
    Ul[0] = rot<0>(jc[0].Ul, kc1, ks1); Ua[0] = rot<0>(jc[0].Ua, kc1, ks1) + cross(p1, Ul[0]);
    Ul[1] = rot<0>(rot<1>(jc[1].Ul, kc2, ks2), kc1, ks1); Ua[1] = rot<0>(rot<1>(jc[1].Ua, kc2, ks2), kc1, ks1) + cross(p2, Ul[1]);
    Ul[2] = rot<0>(rot<1>(jc[2].Ul, kc23, ks23), kc1, ks1); Ua[2] = rot<0>(rot<1>(jc[2].Ua, kc23, ks23), kc1, ks1) + cross(p3, Ul[2]);
    const float Di[3] = {jc[0].Dinv, jc[1].Dinv, jc[2].Dinv};

    // ---------------- PMC hurdle plate: getContactPoints (PLE:343) reports the manifolds built on the last sub-step's pre-step poses
    if (ENV == 0 && P.has_ob && sub == P.substeps - 1) {
      const int o0 = mc.ob_off[clip], n_ob = mc.ob_off[clip + 1] - o0;
      if (n_ob > 0) {
        const double* ob = mc.ob_table + (size_t)(o0 + ob_id) * 4;
        float sy, cy;
        llq_sincosf((float)ob[3], &sy, &cy);
        const V3 org = V3{(float)(px - ob[1]), (float)(py - ob[2]), (float)pz};      // base position relative to the plate centre
        const V3 wh = p2 + rot<0>(rot<1>(ld3(M.wheel_off[k]), kc2, ks2), kc1, ks1);
        bool hit = plate_hit(org + mul(R, fb), L.foot_r, cy, sy, P.ob_hx, P.ob_hy, P.ob_hz, P.breaking);
        hit = hit || plate_hit(org + mul(R, wh), M.wheel_r[k], cy, sy, P.ob_hx, P.ob_hy, P.ob_hz, P.breaking);
        hit = hit || plate_hit(org + mul(R, p1), M.hip_r[k], cy, sy, P.ob_hx, P.ob_hy, P.ob_hz, P.breaking);
        hit = hit || plate_hit(org + mul(R, ld3(M.corner[2 * k])), 0.f, cy, sy, P.ob_hx, P.ob_hy, P.ob_hz, P.breaking);
        hit = hit || plate_hit(org + mul(R, ld3(M.corner[2 * k + 1])), 0.f, cy, sy, P.ob_hx, P.ob_hy, P.ob_hz, P.breaking);
        ob_hit = hit;
      }
    }
    // ---------------- SEPMC: getContactPoints() (CTG:426-456) = manifolds of the last sub-step, built on its pre-step poses
    if (ENV == 2 && sub == P.substeps - 1) {
      float* srow = &s_new[threadIdx.x >> 2][0];
      const float* prow = &s_new[(threadIdx.x >> 2) ^ 1][0];
      const V3 pw = V3{(float)px, (float)py, (float)pz};
      const V3 wh = pw + mul(R, p2 + rot<0>(rot<1>(ld3(M.wheel_off[k]), kc2, ks2), kc1, ks1));
      const V3 hp = pw + mul(R, p1), ft = pw + mul(R, fb);
      const V3 c0 = pw + mul(R, ld3(M.corner[2 * k])), c1_ = pw + mul(R, ld3(M.corner[2 * k + 1]));
      float* o = srow + 18 * k;
      o[0] = ft.x; o[1] = ft.y; o[2] = ft.z; o[3] = wh.x; o[4] = wh.y; o[5] = wh.z; o[6] = hp.x; o[7] = hp.y; o[8] = hp.z;
      o[9] = c0.x; o[10] = c0.y; o[11] = c0.z; o[12] = c1_.x; o[13] = c1_.y; o[14] = c1_.z;
      if (k < 2) { const V3 hd = pw + mul(R, V3{M.handle[k][0], M.handle[k][1], M.handle[k][2]}); o[15] = hd.x; o[16] = hd.y; o[17] = hd.z; }
      __syncwarp();
      const float fx = (float)PS.flag_x, fy = (float)PS.flag_y;
      // the robot's "body" links (legs + wheels, CTG:427) are represented by its hip and wheel spheres
      bool tch = flag_dist(hp, fx, fy) - M.hip_r[k] < P.breaking || flag_dist(wh, fx, fy) - M.wheel_r[k] < P.breaking;
      bool tg = false;
#pragma unroll 1
      for (int j = 0; j < 4; j++) {
        const float* pj = prow + 18 * j;
        const float rj[6] = {M.leg[j].foot_r, M.wheel_r[j], M.hip_r[j], 0.f, 0.f, M.handle[j & 1][3]};
#pragma unroll
        for (int t = 0; t < 6; t++) {
          if (t == 5 && j >= 2) continue;
          const V3 c = V3{pj[3 * t], pj[3 * t + 1], pj[3 * t + 2]};
          tg = tg || norm3(hp - c) - M.hip_r[k] - rj[t] < P.breaking || norm3(wh - c) - M.wheel_r[k] - rj[t] < P.breaking;
        }
      }
      int bits = (tch ? 1 : 0) | (tg ? 2 : 0);
      bits |= __shfl_xor_sync(FULL, bits, 1);
      bits |= __shfl_xor_sync(FULL, bits, 2);
      const int other = __shfl_xor_sync(FULL, bits, 4);
      touch_own = (bits & 1) != 0;
      tag = ((robot == 0 ? bits : other) & 2) != 0;               // only robot 0's body counts (CTG:464)
      __syncwarp();
    }
    // ---------------- collision: foot sphere vs plane z = 0 on the pre-step pose
    V3 nb = V3{R.a20, R.a21, R.a22};                     // world z in base coords
    int plane = 0;                                       // SEPMC: 0 ground, 1..4 walls with normals -x, +x, -y, +y; EPMC corridor: 5 = a box
    V3 nworld = V3{0.f, 0.f, 1.f};                       // plane 5: contact normal in world coordinates
    // The foot clearance feeds Bullet's speculative-contact target (-penetration/dt): a 1e-7 m rounding error becomes
    // 5e-5 m/s.  Evaluate just this scalar (height of the foot centre) in fp64 from the fp32 joint sines/cosines.
    float dist;
    {
      const double qx = qp.x, qy = qp.y, qz = qp.z, qw = qp.w;
      const double nx = 2.0 * (qx * qz - qy * qw), ny = 2.0 * (qy * qz + qx * qw), nz = 1.0 - 2.0 * (qx * qx + qy * qy);
      const double dc1 = kc1, ds1 = ks1, dc2 = kc2, ds2 = ks2, dc3 = jc[2].c, ds3 = jc[2].s;
      // foot in shank frame -> thigh frame -> hip frame -> base (same chain as fb, in double)
      double x = L.foot[0], y = L.foot[1], z = L.foot[2], t;
      t = dc3 * x + ds3 * z; z = -ds3 * x + dc3 * z; x = t;            // Ry(theta3)
      x += (double)r[2].x; y += (double)r[2].y; z += (double)r[2].z;
      t = dc2 * x + ds2 * z; z = -ds2 * x + dc2 * z; x = t;            // Ry(theta2)
      x += (double)r[1].x; y += (double)r[1].y; z += (double)r[1].z;
      t = dc1 * y - ds1 * z; z = ds1 * y + dc1 * z; y = t;             // Rx(q1)
      x += (double)r[0].x; y += (double)r[0].y; z += (double)r[0].z;
      dist = (float)(pz + nx * x + ny * y + nz * z - (double)L.foot_r);
      if (ENV == 3) {
        // EPMC corridor: sphere vs the candidate boxes, in fp64 like the ground clearance; one contact per foot, the deepest
        const double wx = px + (1.0 - 2.0 * (qy * qy + qz * qz)) * x + 2.0 * (qx * qy - qz * qw) * y + 2.0 * (qx * qz + qy * qw) * z;
        const double wy = py + 2.0 * (qx * qy + qz * qw) * x + (1.0 - 2.0 * (qx * qx + qz * qz)) * y + 2.0 * (qy * qz - qx * qw) * z;
        const double wz = pz + nx * x + ny * y + nz * z;
        for (int c = 0; c < n_cand; c++) {
          const float* b = s_cand + 6 * c;
          const double p0 = wx - (double)b[0], p1 = wy - (double)b[1], p2 = wz - (double)b[2];
          const double h0 = b[3], h1 = b[4], h2 = b[5];
          const double c0 = fmin(fmax(p0, -h0), h0), c1 = fmin(fmax(p1, -h1), h1), c2 = fmin(fmax(p2, -h2), h2);
          double db; V3 nn;
          if (c0 != p0 || c1 != p1 || c2 != p2) {
            const double v0 = p0 - c0, v1 = p1 - c1, v2 = p2 - c2;
            const double len = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
            db = len - (double)L.foot_r;
            nn = V3{(float)(v0 / len), (float)(v1 / len), (float)(v2 / len)};
          } else {                                 // centre inside the box: leave through the nearest face
            double best = h0 - p0; nn = V3{1.f, 0.f, 0.f};
            if (h0 + p0 < best) { best = h0 + p0; nn = V3{-1.f, 0.f, 0.f}; }
            if (h1 - p1 < best) { best = h1 - p1; nn = V3{0.f, 1.f, 0.f}; }
            if (h1 + p1 < best) { best = h1 + p1; nn = V3{0.f, -1.f, 0.f}; }
            if (h2 - p2 < best) { best = h2 - p2; nn = V3{0.f, 0.f, 1.f}; }
            if (h2 + p2 < best) { best = h2 + p2; nn = V3{0.f, 0.f, -1.f}; }
            db = -best - (double)L.foot_r;
          }
          if ((float)db < dist) { dist = (float)db; plane = 5; nworld = nn; }
        }
      }
      if (ENV == 2) {
        // the arena walls (BSG:863-902) as four more half-spaces; one contact per foot, the deepest
        const double wx = px + (1.0 - 2.0 * (qy * qy + qz * qz)) * x + 2.0 * (qx * qy - qz * qw) * y + 2.0 * (qx * qz + qy * qw) * z;
        const double wy = py + 2.0 * (qx * qy + qz * qw) * x + (1.0 - 2.0 * (qx * qx + qz * qz)) * y + 2.0 * (qy * qz - qx * qw) * z;
        const double lim = (double)kWallIn - (double)L.foot_r;
        const float d1 = (float)(lim - wx), d2 = (float)(lim + wx), d3 = (float)(lim - wy), d4 = (float)(lim + wy);
        if (d1 < dist) { dist = d1; plane = 1; }
        if (d2 < dist) { dist = d2; plane = 2; }
        if (d3 < dist) { dist = d3; plane = 3; }
        if (d4 < dist) { dist = d4; plane = 4; }
      }
    }
    // knee wheel vs ground (fp64 clearance like the foot); one contact per leg: the deeper of {foot, knee wheel}
    bool onwheel = false;
    V3 cb = fb;                                           // centre of the contact sphere, base coordinates
    float crad = L.foot_r;
    const V3 wheel_b = p2 + rot<0>(rot<1>(ld3(M.wheel_off[k]), kc2, ks2), kc1, ks1);
    // cheap fp32 screen first: the fp64 clearance is only needed when the wheel is within 1 cm of becoming the leg's contact
    if (P.knee && (float)pz + dot(nb, wheel_b) - M.wheel_r[k] < fmaxf(dist, P.breaking) + 0.01f) {
      const double qx = qp.x, qy = qp.y, qz = qp.z, qw = qp.w;
      const double nx = 2.0 * (qx * qz - qy * qw), ny = 2.0 * (qy * qz + qx * qw), nz = 1.0 - 2.0 * (qx * qx + qy * qy);
      const double dc1 = kc1, ds1 = ks1, dc2 = kc2, ds2 = ks2;
      double x = M.wheel_off[k][0], y = M.wheel_off[k][1], z = M.wheel_off[k][2], t;
      t = dc2 * x + ds2 * z; z = -ds2 * x + dc2 * z; x = t;            // Ry(theta2)
      x += (double)r[1].x; y += (double)r[1].y; z += (double)r[1].z;
      t = dc1 * y - ds1 * z; z = ds1 * y + dc1 * z; y = t;             // Rx(q1)
      x += (double)r[0].x; y += (double)r[0].y; z += (double)r[0].z;
      const float dw = (float)(pz + nx * x + ny * y + nz * z - (double)M.wheel_r[k]);
      if (dw < dist) {
        dist = dw; onwheel = true; plane = 0;
        cb = wheel_b;
        crad = M.wheel_r[k];
      }
    }
    const bool contact = dist < P.breaking;
    if (!contact || warm_wheel != onwheel) warm = 0.f;   // another manifold point: no warm start
    warm_wheel = contact && onwheel;
    // joint-limit rows (btMultiBodyJointLimitConstraint: a row exists only while the limit is violated)
    float limdir[3];
    unsigned mymask = contact ? 7u : 0u;   // bits 0-2: contact rows n,t1,t2 ; bits 3-5: limit rows of joints 0-2
#pragma unroll
    for (int i = 0; i < 3; i++) {
      limdir[i] = 0.f;
      if (L.j[i].haslim) {
        if (q[i] - L.j[i].lower <= 0.f) limdir[i] = 1.f;
        else if (L.j[i].upper - q[i] <= 0.f) limdir[i] = -1.f;
      }
      if (limdir[i] != 0.f) mymask |= 8u << i;
    }
    const unsigned envmask = __shfl_sync(FULL, mymask, 0, 4) | (__shfl_sync(FULL, mymask, 1, 4) << 6) |
                             (__shfl_sync(FULL, mymask, 2, 4) << 12) | (__shfl_sync(FULL, mymask, 3, 4) << 18);
    const unsigned warpmask = __reduce_or_sync(FULL, envmask);
    const bool any_con_warp = (warpmask & 0x1C71C7u) != 0;   // bits 0-2 of each 6-bit group
    const bool any_lim_warp = (warpmask & 0xE38E38u) != 0;   // bits 3-5 of each 6-bit group
    float dvb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dvl[3] = {0.f, 0.f, 0.f};

    if (warpmask) {
      // Own rows: 0..2 = contact (n, t1, t2) -- hot path, kept in registers, loops unrolled;
      //           3..5 = violated joint limits -- rare path, kept in shared memory, loops rolled.
      // For each row we keep its image under the ABA's L^-1 factor: y = L0^-1 Fhat (base part, 6) and u (joint part, 3).
      // Then  J_r M^-1 J_s^T = y_r.y_s + sum_i u_ri u_si / D_i  (second term only for rows on the same leg), so no
      // per-row down passes are needed; PGS only tracks b_r = J_r . (delta v) of the own rows.
      float yc[3][6], uc[3][3], Ac[3][12];
      float bq[3] = {0.f, 0.f, 0.f}, rhs[3] = {0.f, 0.f, 0.f}, invd[3] = {0.f, 0.f, 0.f}, lam[3] = {0.f, 0.f, 0.f};
      float bl[3] = {0.f, 0.f, 0.f}, rhsl[3] = {0.f, 0.f, 0.f}, invdl[3] = {0.f, 0.f, 0.f}, laml[3] = {0.f, 0.f, 0.f};
      // shared-memory workspace of this lane (element e of lane tid at ws[e * BLOCK])
      float* ws = rows_sm + tid;
      constexpr int W_YC = 0, W_YL = 18, W_UL = 36, W_ACL = 45, W_ALC = 81, W_ALL = 117;   // kRowFloats = 153
      const int g0 = tid & ~3;
      if (any_con_warp) {
        // directions (world): n = +z, t1 = -y, t2 = +x   (btPlaneSpace1 of the plane normal), in base coords
        V3 dirs[3] = {nb, neg(V3{R.a10, R.a11, R.a12}), V3{R.a00, R.a01, R.a02}};
        if (ENV == 2 && plane != 0) {                     // btPlaneSpace1 of the wall normals
          const V3 r0 = V3{R.a00, R.a01, R.a02}, r1 = V3{R.a10, R.a11, R.a12};
          dirs[2] = nb;                                   // t2 = +z for every wall
          if (plane == 1) { dirs[0] = neg(r0); dirs[1] = neg(r1); }
          else if (plane == 2) { dirs[0] = r0; dirs[1] = r1; }
          else if (plane == 3) { dirs[0] = neg(r1); dirs[1] = r0; }
          else { dirs[0] = r1; dirs[1] = neg(r0); }
        }
        if (ENV == 3 && plane == 5) {                     // general normal: btPlaneSpace1 in world axes, then into base coordinates
          const V3 n = nworld;
          V3 t1, t2;
          if (fabsf(n.z) > 0.70710678118654752f) {
            const float a = n.y * n.y + n.z * n.z, kk = rsqrtf(a);
            t1 = V3{0.f, -n.z * kk, n.y * kk};
            t2 = V3{a * kk, -n.x * t1.z, n.x * t1.y};
          } else {
            const float a = n.x * n.x + n.y * n.y, kk = rsqrtf(a);
            t1 = V3{-n.y * kk, n.x * kk, 0.f};
            t2 = V3{-n.z * t1.y, n.z * t1.x, a * kk};
          }
          dirs[0] = tmul(R, n); dirs[1] = tmul(R, t1); dirs[2] = tmul(R, t2);
        }
        const V3 Pc = cb - crad * dirs[0];                // contact point on the sphere surface
#pragma unroll
        for (int d = 0; d < 3; d++) {
          const V3 db = dirs[d];
          V3 Ga = cross(Pc, db), Gl = db;                 // spatial force of a unit impulse, about the base origin
          const float rel0 = dot(Ga, wbs) + dot(Gl, vbs);
          uc[d][2] = onwheel ? 0.f : dot(Sa[2], Ga) + dot(Sl[2], Gl);   // the wheel sits on the thigh: the shank joint does not move it
          const float j2 = dot(Sa[1], Ga) + dot(Sl[1], Gl), j1 = dot(Sa[0], Ga) + dot(Sl[0], Gl);
          const float rel = rel0 + j1 * qd[0] + j2 * qd[1] + uc[d][2] * qd[2];
          float g = uc[d][2] * Di[2];
          Ga = fma3(-g, Ua[2], Ga); Gl = fma3(-g, Ul[2], Gl);
          uc[d][1] = dot(Sa[1], Ga) + dot(Sl[1], Gl);
          g = uc[d][1] * Di[1];
          Ga = fma3(-g, Ua[1], Ga); Gl = fma3(-g, Ul[1], Gl);
          uc[d][0] = dot(Sa[0], Ga) + dot(Sl[0], Gl);
          g = uc[d][0] * Di[0];
          Ga = fma3(-g, Ua[0], Ga); Gl = fma3(-g, Ul[0], Gl);
          { const float bb[6] = {Ga.x, Ga.y, Ga.z, Gl.x, Gl.y, Gl.z}; chol6_fwd(ch, bb, yc[d]); }
          float dg = uc[d][0] * uc[d][0] * Di[0] + uc[d][1] * uc[d][1] * Di[1] + uc[d][2] * uc[d][2] * Di[2];
#pragma unroll
          for (int t = 0; t < 6; t++) { dg = fmaf(yc[d][t], yc[d][t], dg); ws[(W_YC + d * 6 + t) * BLOCK] = yc[d][t]; }
          invd[d] = contact ? 1.0f / dg : 0.f;
          if (d == 0) {   // btMultiBodyConstraintSolver::setupMultiBodyContactConstraint
            float pen = dist + P.slop, poserr = 0.f, velerr = -rel;
            if (pen > 0.f) velerr -= pen / dt; else poserr = -pen * P.erp / dt;
            rhs[0] = (poserr + velerr) * invd[0];
            lam[0] = contact ? P.warm * warm : 0.f;
          } else {
            rhs[d] = -rel * invd[d];
          }
        }
        if (contact) n_contact_rows += 3;
      }
      if (any_lim_warp) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const float dir = limdir[i];
          float u[3] = {0.f, 0.f, 0.f};
          u[i] = dir;
          float g = dir * Di[i];
          V3 Ga = (-g) * Ua[i], Gl = (-g) * Ul[i];
#pragma unroll
          for (int m = i - 1; m >= 0; m--) {
            u[m] = dot(Sa[m], Ga) + dot(Sl[m], Gl);
            g = u[m] * Di[m];
            Ga = fma3(-g, Ua[m], Ga); Gl = fma3(-g, Ul[m], Gl);
          }
          float y[6];
          { const float bb[6] = {Ga.x, Ga.y, Ga.z, Gl.x, Gl.y, Gl.z}; chol6_fwd(ch, bb, y); }
          float dg = u[0] * u[0] * Di[0] + u[1] * u[1] * Di[1] + u[2] * u[2] * Di[2];
#pragma unroll
          for (int t = 0; t < 6; t++) { dg = fmaf(y[t], y[t], dg); ws[(W_YL + i * 6 + t) * BLOCK] = y[t]; }
#pragma unroll
          for (int m = 0; m < 3; m++) ws[(W_UL + i * 3 + m) * BLOCK] = u[m];
          {   // branch-free: an absent row keeps invdl = rhsl = 0
            const bool act = dir != 0.f;
            const float rel = dir * qd[i];
            const float pen = dir > 0.f ? q[i] - L.j[i].lower : L.j[i].upper - q[i];
            invdl[i] = act ? 1.0f / dg : 0.f;
            const float poserr = pen > -0.04f ? -pen * P.jerp / dt : 0.f;   // split-impulse threshold quirk
            rhsl[i] = act ? (poserr - rel) * invdl[i] : 0.f;
            n_limit_rows += act ? 1 : 0;
          }
        }
      }
      __syncwarp();
      // ---- Delassus blocks.  contact x contact: registers, unrolled (the y's of the other legs come through smem)
      if (any_con_warp) {
        float ut[3][3];   // same-leg joint term  sum_i u_ri u_si / D_i  of the own contact rows
#pragma unroll
        for (int rr = 0; rr < 3; rr++)
#pragma unroll
          for (int sr = 0; sr < 3; sr++)
            ut[rr][sr] = uc[rr][0] * uc[sr][0] * Di[0] + uc[rr][1] * uc[sr][1] * Di[1] + uc[rr][2] * uc[sr][2] * Di[2];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          {   // all four legs, always: a warp whose 8 envs all have leg j in the air is a 2 % case, and the uniform skip cost more than it saved
            const float* Yj = rows_sm + g0 + j;
            const bool own = (j == k);
#pragma unroll
            for (int sr = 0; sr < 3; sr++) {
              float ys[6];
#pragma unroll
              for (int t = 0; t < 6; t++) ys[t] = Yj[(W_YC + sr * 6 + t) * BLOCK];
#pragma unroll
              for (int rr = 0; rr < 3; rr++) {
                float acc = own ? ut[rr][sr] : 0.f;
#pragma unroll
                for (int t = 0; t < 6; t++) acc = fmaf(yc[rr][t], ys[t], acc);
                Ac[rr][3 * j + sr] = acc;
              }
            }
          }
        }
      }
      // blocks that involve limit rows: shared memory, rolled loops (rare path)
      if (any_lim_warp) {
#pragma unroll 1
        for (int jl = 0; jl < 4; jl++) {
          const bool own = (jl == k);
#pragma unroll 1
          for (int il = 0; il < 3; il++) {
            if (!((warpmask >> (6 * jl + 3 + il)) & 1u)) continue;       // source: limit row (jl, il)
            float ys[6], us[3];
#pragma unroll
            for (int t = 0; t < 6; t++) ys[t] = rows_sm[(W_YL + il * 6 + t) * BLOCK + g0 + jl];
#pragma unroll
            for (int m = 0; m < 3; m++) us[m] = rows_sm[(W_UL + il * 3 + m) * BLOCK + g0 + jl] * (own ? Di[m] : 0.f);
#pragma unroll
            for (int rr = 0; rr < 3; rr++) {                              // targets: own contact rows, own limit rows
              float a1 = uc[rr][0] * us[0] + uc[rr][1] * us[1] + uc[rr][2] * us[2];
              float a2 = ws[(W_UL + rr * 3 + 0) * BLOCK] * us[0] + ws[(W_UL + rr * 3 + 1) * BLOCK] * us[1] + ws[(W_UL + rr * 3 + 2) * BLOCK] * us[2];
#pragma unroll
              for (int t = 0; t < 6; t++) { a1 = fmaf(yc[rr][t], ys[t], a1); a2 = fmaf(ws[(W_YL + rr * 6 + t) * BLOCK], ys[t], a2); }
              ws[(W_ACL + rr * 12 + 3 * jl + il) * BLOCK] = any_con_warp ? a1 : 0.f;
              ws[(W_ALL + rr * 12 + 3 * jl + il) * BLOCK] = a2;
            }
          }
        }
        // (own limit row rr) x (contact row sr of leg jl) is the transpose of the entry lane jl just wrote for (its contact row sr) x
        // (limit slot (k, rr)): fetch it from that lane's workspace instead of recomputing 36 nine-term dot products
        if (any_con_warp) {
          __syncwarp();
          const float* other = rows_sm + g0 + (W_ACL + 3 * k) * BLOCK;
#pragma unroll 1
          for (int jl = 0; jl < 4; jl++) {
#pragma unroll
            for (int sr = 0; sr < 3; sr++)
#pragma unroll
              for (int rr = 0; rr < 3; rr++) {
                const float v = other[(sr * 12 + rr) * BLOCK + jl];
                ws[(W_ALC + rr * 12 + 3 * jl + sr) * BLOCK] = ((warpmask >> (6 * k + 3 + rr)) & 1u) ? v : 0.f;
              }
          }
        }
      }
      // ---- projected Gauss-Seidel (btMultiBodyConstraintSolver::solveSingleIteration order: limits, normals, friction)
      // an impulse dl_ on contact column (j_, d_) of the env: own contact rows from registers, own limit rows from smem
#define LLQ_APPLY_C(j_, d_, dl_)                                                                   \
      {                                                                                            \
        bq[0] = fmaf(Ac[0][3 * (j_) + (d_)], dl_, bq[0]);                                          \
        bq[1] = fmaf(Ac[1][3 * (j_) + (d_)], dl_, bq[1]);                                          \
        bq[2] = fmaf(Ac[2][3 * (j_) + (d_)], dl_, bq[2]);                                          \
        if (any_lim_warp) {                                                                        \
          bl[0] = fmaf(ws[(W_ALC + 0 * 12 + 3 * (j_) + (d_)) * BLOCK], dl_, bl[0]);                \
          bl[1] = fmaf(ws[(W_ALC + 1 * 12 + 3 * (j_) + (d_)) * BLOCK], dl_, bl[1]);                \
          bl[2] = fmaf(ws[(W_ALC + 2 * 12 + 3 * (j_) + (d_)) * BLOCK], dl_, bl[2]);                \
        }                                                                                          \
      }
      if (any_con_warp) {   // warm start of the normal rows
#pragma unroll
        for (int j = 0; j < 4; j++) {
          {
            const float l0 = __shfl_sync(FULL, lam[0], j, 4);      // 0 for feet without contact
            LLQ_APPLY_C(j, 0, l0)
          }
        }
      }
      const float mu = ENV != 0 ? mu_env : P.mu;
#pragma unroll 1
      for (int it = 0; it < P.solver_iters; it++) {
        if (any_lim_warp) {
#pragma unroll 1
          for (int j = 0; j < 4; j++) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
              if (!((warpmask >> (6 * j + 3 + i)) & 1u)) continue;
              // branch-free: every lane evaluates its own row i, only the owner of an existing row keeps the result; the
              // coefficients of a warp-active slot are finite in every env (images of absent rows are zero), so applying
              // dl = 0 there is exact
              const bool mine = (j == k) && limdir[i] != 0.f;
              const float dlc = rhsl[i] - bl[i] * invdl[i];
              const float sum = laml[i] + dlc;
              const bool lo = sum < 0.f, hi = sum > P.max_imp;
              float dl = lo ? -laml[i] : (hi ? P.max_imp - laml[i] : dlc);
              const float ln = lo ? 0.f : (hi ? P.max_imp : sum);
              dl = mine ? dl : 0.f;
              laml[i] = mine ? ln : laml[i];
              dl = __shfl_sync(FULL, dl, j, 4);
              {
                const int col = 3 * j + i;
                bl[0] = fmaf(ws[(W_ALL + 0 * 12 + col) * BLOCK], dl, bl[0]);
                bl[1] = fmaf(ws[(W_ALL + 1 * 12 + col) * BLOCK], dl, bl[1]);
                bl[2] = fmaf(ws[(W_ALL + 2 * 12 + col) * BLOCK], dl, bl[2]);
                if (any_con_warp) {
                  bq[0] = fmaf(ws[(W_ACL + 0 * 12 + col) * BLOCK], dl, bq[0]);
                  bq[1] = fmaf(ws[(W_ACL + 1 * 12 + col) * BLOCK], dl, bq[1]);
                  bq[2] = fmaf(ws[(W_ACL + 2 * 12 + col) * BLOCK], dl, bq[2]);
                }
              }
            }
          }
        }
        if (any_con_warp) {
          // branch-free row updates: only lane j of an env owns foot j's rows, everybody else contributes dl = 0 (the A
          // columns of feet without contact are finite, so applying a zero impulse is exact)
#pragma unroll
          for (int j = 0; j < 4; j++) {   // normal rows, feet in order FR FL HR HL
            {
              const bool mine = (j == k) && contact;
              const float dlc = rhs[0] - bq[0] * invd[0];
              const float sum = lam[0] + dlc;
              const bool lo = sum < 0.f, hi = sum > 1e10f;
              float dl = lo ? -lam[0] : (hi ? 1e10f - lam[0] : dlc);
              const float ln = lo ? 0.f : (hi ? 1e10f : sum);
              dl = mine ? dl : 0.f;
              lam[0] = mine ? ln : lam[0];
              dl = __shfl_sync(FULL, dl, j, 4);
              LLQ_APPLY_C(j, 0, dl)
            }
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {   // friction pairs with the implicit cone (resolveConeFrictionConstraintRows)
            {
              const bool mine = (j == k) && contact;
              float sa = lam[1] + (rhs[1] - bq[1] * invd[1]), sb = lam[2] + (rhs[2] - bq[2] * invd[2]);
              const float limit = (onwheel ? P.mu_wheel : mu) * lam[0];
              const float r2 = sa * sa + sb * sb;
              const bool clip = r2 >= limit * limit && r2 > 0.f;
              const float sc = clip ? limit * rsqrtf(r2) : 1.0f;
              sa = clip ? sa * sc : sa; sb = clip ? sb * sc : sb;
              float da = mine ? sa - lam[1] : 0.f, db = mine ? sb - lam[2] : 0.f;
              lam[1] = mine ? sa : lam[1]; lam[2] = mine ? sb : lam[2];
              da = __shfl_sync(FULL, da, j, 4);
              db = __shfl_sync(FULL, db, j, 4);
              LLQ_APPLY_C(j, 1, da)
              LLQ_APPLY_C(j, 2, db)
            }
          }
        }
      }
#undef LLQ_APPLY_C
      if (contact) warm = lam[0];
      // ---- total impulse -> velocity change: one back substitution and one down pass (base coordinates)
      float Yt[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, wt[3] = {0.f, 0.f, 0.f};
      if (any_con_warp) {      // lam = 0 on feet without contact: no per-lane branch needed
#pragma unroll
        for (int rr = 0; rr < 3; rr++) {
#pragma unroll
          for (int t = 0; t < 6; t++) Yt[t] = fmaf(lam[rr], yc[rr][t], Yt[t]);
#pragma unroll
          for (int i = 0; i < 3; i++) wt[i] = fmaf(lam[rr], uc[rr][i], wt[i]);
        }
      }
      if (any_lim_warp) {
#pragma unroll
        for (int rr = 0; rr < 3; rr++) {
          {                    // laml = 0 on absent rows
#pragma unroll
            for (int t = 0; t < 6; t++) Yt[t] = fmaf(laml[rr], ws[(W_YL + rr * 6 + t) * BLOCK], Yt[t]);
#pragma unroll
            for (int i = 0; i < 3; i++) wt[i] = fmaf(laml[rr], ws[(W_UL + rr * 3 + i) * BLOCK], wt[i]);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 6; t++) Yt[t] = gsum4(Yt[t]);
      chol6_bwd(ch, Yt, dvb);
      {
        V3 aa = V3{dvb[0], dvb[1], dvb[2]}, al = V3{dvb[3], dvb[4], dvb[5]};
        dvl[0] = (wt[0] - dot(Ua[0], aa) - dot(Ul[0], al)) * Di[0];
        aa = fma3(dvl[0], Sa[0], aa); al = fma3(dvl[0], Sl[0], al);
        dvl[1] = (wt[1] - dot(Ua[1], aa) - dot(Ul[1], al)) * Di[1];
        aa = fma3(dvl[1], Sa[1], aa); al = fma3(dvl[1], Sl[1], al);
        dvl[2] = (wt[2] - dot(Ua[2], aa) - dot(Ul[2], al)) * Di[2];
      }
      __syncwarp();
    }

#if LLQ_BARRIERS >= 4
    if (BLOCK > 32) __syncthreads();
#endif
    // ---------------- apply the impulses, clamp, integrate (btMultiBody::stepPositionsMultiDof)
    {
      V3 dw = mul(R, V3{dvb[0], dvb[1], dvb[2]}), dv = mul(R, V3{dvb[3], dvb[4], dvb[5]});
      ww = V3{clampf(ww.x + dw.x, -P.vmax, P.vmax), clampf(ww.y + dw.y, -P.vmax, P.vmax), clampf(ww.z + dw.z, -P.vmax, P.vmax)};
      vw = V3{clampf(vw.x + dv.x, -P.vmax, P.vmax), clampf(vw.y + dv.y, -P.vmax, P.vmax), clampf(vw.z + dv.z, -P.vmax, P.vmax)};
#pragma unroll
      for (int i = 0; i < 3; i++) {
        qd[i] = clampf(qd[i] + dvl[i], -P.vmax, P.vmax);
        q[i] = fmaf(qd[i], dt, q[i]);
      }
      px += (double)vw.x * P.sim_dt; py += (double)vw.y * P.sim_dt; pz += (double)vw.z * P.sim_dt;
      float fa = norm3(ww);
      float sc;
      sc = 0.5f * dt - dt * dt * dt * 0.020833333333f * fa * fa;      // used below 1e-3 rad/s (btMultiBody's series)
      float sh, ch;
      llq_sincosf(0.5f * fa * dt, &sh, &ch);
      if (!(fa < 0.001f)) sc = sh / fa;
      Q4 dq = Q4{sc * ww.x, sc * ww.y, sc * ww.z, ch};
      qp = qnormalize(qmul(dq, qp));
    }
    bad = bad || !(fabsf(qd[0]) <= P.vmax) || !(fabsf(ww.x) <= P.vmax) || !(fabsf(vw.x) <= P.vmax);
    // ---------------- mocap clock (PLE:208-210): sampled with the time *before* the increment
    if (ENV == 0) {
      frame_id = (int)floor(time / P.frame_dt);
      frame_frac = (time - frame_id * P.frame_dt) / P.frame_dt;
      // a cursor past the clip's playable range (auto_reset off and a finished env stepped on, or a clock set through
      // llq_set_field) stays on the clip's last playable frame instead of walking into the next clip; the reference raises there
      // (the last policy step of an episode legitimately runs up to 2.4 frames past the "ended" threshold nf - margin - 1; the bound
      // is the last cursor whose 1 s future window (122 frames) still lies inside the clip)
      const int last = mc.clip_off[clip + 1] - mc.clip_off[clip] - P.margin + 2;
      if (frame_id > last) { frame_id = last; frame_frac = 0.0; }
      if (frame_id < 0) { frame_id = 0; frame_frac = 0.0; }
    }
    time += P.sim_dt;
  }

  // ================= end of the policy step: observation, reward, termination =================
  bool done = false;
  float rew_out = 0.f;
  if (ENV == 0) {
  qb = qmul(qp, qI);                                 // back to the pybullet (inertial-frame) convention
  float* snew = &s_new[threadIdx.x >> 2][0];
  ObsCtx oc = build_obs_new(mc, P, M, k, clip, frame_id, frame_frac, px, py, pz, qb, vw, ww, q, qd, snew);
#pragma unroll
  for (int i = 0; i < 3; i++) snew[kPropDim + 3 * k + i] = act[i];

  // reward (PLE:350-426)
  float djp = 0.f, djv = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) { float a = q[i] - oc.kq[i], b = qd[i] - oc.kqd[i]; djp = fmaf(a, a, djp); djv = fmaf(b, b, djv); }
  V3 fd, fk;
  {
    M3 Rp = qmat(qp);
    V3 f = mul(Rp, foot_in_base(L, q[0], q[1], q[2]));
    fd = V3{(float)px + f.x, (float)py + f.y, (float)pz + f.z};
    Q4 kqp = qmul(qnormalize(oc.kb.q), qconj(qI));
    V3 g = mul(qmat(kqp), foot_in_base(L, oc.kq[0], oc.kq[1], oc.kq[2]));
    // difference of foot positions, formed in double for the base offset
    fk = V3{(float)(oc.kb.px - px) + g.x - f.x, (float)(oc.kb.py - py) + g.y - f.y, (float)(oc.kb.pz - pz) + g.z - f.z};
  }
  float dee = dot(fk, fk);
  djp = gsum4(djp); djv = gsum4(djv); dee = gsum4(dee);
  float dpx = (float)(px - oc.kb.px), dpy = (float)(py - oc.kb.py), dpz = (float)(pz - oc.kb.pz);
  float dp = dpx * dpx + dpy * dpy + dpz * dpz;
  V3 dvl3 = vw - oc.kb.lin, dva3 = ww - oc.kb.ang;
  Q4 q1 = qnormalize(qb), q2 = qnormalize(oc.kb.q);
  float angle = norm3(q_rotvec(qnormalize(qmul(q2, qconj(q1)))));
  float rew = P.w_jp * expf(-1.0f * djp) + P.w_jv * expf(-0.1f * djv) + P.w_ee * expf(-40.0f * dee) +
              P.w_pose * expf(-20.0f * dp - 10.0f * angle * angle) + P.w_vel * expf(-2.0f * dot(dvl3, dvl3) - 0.2f * dot(dva3, dva3));
  // termination (PLE:337-348, LR:158-179, ML:168-172)
  M3 Rq = qmat(q1);
  float left_z = Rq.a02 * Rq.a10 - Rq.a12 * Rq.a00;
  bool fall = left_z > 0.70710678118654752f || left_z < -0.70710678118654752f || Rq.a22 < 0.5f;
  int nf = mc.clip_off[clip + 1] - mc.clip_off[clip];
  bool ended = frame_id >= nf - P.margin - 1;
  bool diff = fabsf(angle) > 1.0f || dp > 1.0f;
  {
    int bi = bad ? 1 : 0;
    bi |= __shfl_xor_sync(FULL, bi, 1);
    bi |= __shfl_xor_sync(FULL, bi, 2);
    bad = bi != 0;
  }
  if (bad || !isfinite(rew)) { rew = 0.f; bad = true; }
  if (P.has_ob) {
    int oh = ob_hit ? 1 : 0;
    oh |= __shfl_xor_sync(FULL, oh, 1);
    oh |= __shfl_xor_sync(FULL, oh, 2);
    ob_hit = oh != 0;
    const int o0 = mc.ob_off[clip], n_ob = mc.ob_off[clip + 1] - o0;                 // PLE:262-268 hand-over to the next plate
    while (ob_id < n_ob - 1 && time > mc.ob_table[(size_t)(o0 + ob_id) * 4] + 0.5) ob_id++;
  }
  done = fall || ended || diff || ob_hit || bad;                                     // PLE:347

  // ---- write back state (SoA)
  if (valid) {
    float* sw = E.st;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      sw[(10 + 3 * k + i) * N + env] = q[i];
      sw[(22 + 3 * k + i) * N + env] = qd[i];
      E.kin[(13 + 3 * k + i) * N + env] = oc.kq[i];
      E.kin[(25 + 3 * k + i) * N + env] = oc.kqd[i];
    }
    E.warm[k * N + env] = warm_wheel ? -warm : warm;
    E.foot_pos[(3 * k) * N + env] = fd.x; E.foot_pos[(3 * k + 1) * N + env] = fd.y; E.foot_pos[(3 * k + 2) * N + env] = fd.z;
    if (k == 0) {
      E.pos[env] = px; E.pos[N + env] = py; E.pos[2 * N + env] = pz;
      sw[env] = qb.x; sw[N + env] = qb.y; sw[2 * N + env] = qb.z; sw[3 * N + env] = qb.w;
      sw[4 * N + env] = vw.x; sw[5 * N + env] = vw.y; sw[6 * N + env] = vw.z;
      sw[7 * N + env] = ww.x; sw[8 * N + env] = ww.y; sw[9 * N + env] = ww.z;
      E.time[env] = time;
      if (P.has_ob) E.ob_id[env] = ob_id;
      float rs = E.reward_sum[env] + rew;
      E.reward_sum[env] = rs;
      E.episode_steps[env] += 1;
      E.reward[env] = rew; rew_out = rew;
      E.done[env] = done ? 1 : 0;
      E.kin[env] = (float)oc.kb.px; E.kin[N + env] = (float)oc.kb.py; E.kin[2 * N + env] = (float)oc.kb.pz;
      E.kin[3 * N + env] = oc.kb.q.x; E.kin[4 * N + env] = oc.kb.q.y; E.kin[5 * N + env] = oc.kb.q.z; E.kin[6 * N + env] = oc.kb.q.w;
      E.kin[7 * N + env] = oc.kb.lin.x; E.kin[8 * N + env] = oc.kb.lin.y; E.kin[9 * N + env] = oc.kb.lin.z;
      E.kin[10 * N + env] = oc.kb.ang.x; E.kin[11 * N + env] = oc.kb.ang.y; E.kin[12 * N + env] = oc.kb.ang.z;
      if (done) {
        E.done_reward[env] = rs;
        atomicMax(&winner[clip], env);       // highest finished env index owns the clip's slot this step (PLE:236)
      }
    }
  }
  } else if (ENV == 2) {
    // ---------------- SEPMC tail (CTG:378-424, 458-470, 495-596, 640-652)
    qb = qmul(qp, qI);
    float* snew = &s_new[threadIdx.x >> 2][0];
    const float* spart = &s_new[(threadIdx.x >> 2) ^ 1][0];
    sepmc_pair_tail(M, L, k, robot, snew, spart, px, py, pz, qp, qb, vw, ww, q, touch_own, fix_spd, seed, pair_gid, epi, PS);
#pragma unroll
    for (int i = 0; i < 3; i++) { snew[3 * k + i] = q[i]; snew[12 + 3 * k + i] = qd[i]; snew[kPropDim + 3 * k + i] = act[i]; }
    const float spd = sqrtf(vw.x * vw.x + vw.y * vw.y);              // stat_spd (CTG:368-373)
    total_spd += (double)spd;
    if ((double)spd > max_spd) max_spd = (double)spd;
    counter += 1;
    const M3 Rq = qmat(qnormalize(qb));
    const float left_z = Rq.a02 * Rq.a10 - Rq.a12 * Rq.a00;
    int fall = (left_z > 0.70710678118654752f || left_z < -0.70710678118654752f || Rq.a22 < 0.5f) ? 1 : 0;
    const int fall_other = __shfl_xor_sync(FULL, fall, 4);
    if (robot == 1) fall = fall_other;                                  // only robot 0's fall ends the episode (CTG:462)
    {
      int bi = bad ? 1 : 0;
      bi |= __shfl_xor_sync(FULL, bi, 1);
      bi |= __shfl_xor_sync(FULL, bi, 2);
      bi |= __shfl_xor_sync(FULL, bi, 4);
      bad = bi != 0;
    }
    done = fall != 0 || counter >= P.max_steps || tag || bad;
    // rewards (CTG:640-652, 412-419): +-1 on a flag switch, +-1 on a tag; with_flag after the switch
    const int wf0 = robot == 0 ? PS.with_flag : 1 - PS.with_flag;       // does robot 0 hold the flag
    float rew = (float)PS.sw * ((wf0 != 0) == (robot == 0) ? 1.f : -1.f);
    if (done && tag) rew += (wf0 != 0) == (robot == 0) ? 1.f : -1.f;
    if (bad) rew = 0.f;
    V3 fd;
    {
      V3 f = mul(qmat(qp), foot_in_base(L, q[0], q[1], q[2]));
      fd = V3{(float)px + f.x, (float)py + f.y, (float)pz + f.z};
    }
    if (valid) {
      float* sw = E.st;
#pragma unroll
      for (int i = 0; i < 3; i++) { sw[(10 + 3 * k + i) * N + env] = q[i]; sw[(22 + 3 * k + i) * N + env] = qd[i]; }
      E.warm[k * N + env] = warm_wheel ? -warm : warm;
      E.foot_pos[(3 * k) * N + env] = fd.x; E.foot_pos[(3 * k + 1) * N + env] = fd.y; E.foot_pos[(3 * k + 2) * N + env] = fd.z;
      if (k == 0) {
        E.pos[env] = px; E.pos[N + env] = py; E.pos[2 * N + env] = pz;
        sw[env] = qb.x; sw[N + env] = qb.y; sw[2 * N + env] = qb.z; sw[3 * N + env] = qb.w;
        sw[4 * N + env] = vw.x; sw[5 * N + env] = vw.y; sw[6 * N + env] = vw.z;
        sw[7 * N + env] = ww.x; sw[8 * N + env] = ww.y; sw[9 * N + env] = ww.z;
        E.time[env] = time;
        E.reward_sum[env] += rew;
        E.episode_steps[env] += 1;
        E.reward[env] = rew; rew_out = rew;
        E.done[env] = done ? 1 : 0;
        double* A = E.aux;
        A[env] = counter; A[N + env] = PS.with_flag; A[2 * N + env] = PS.flag_x; A[3 * N + env] = PS.flag_y; A[5 * N + env] = PS.visible;
        A[6 * N + env] = PS.sw; A[7 * N + env] = total_spd; A[8 * N + env] = max_spd; A[9 * N + env] = push_count;
        A[10 * N + env] = pf[0]; A[11 * N + env] = pf[1]; A[12 * N + env] = pf[2]; A[14 * N + env] = push_draws; A[15 * N + env] = PS.flag_draws;
        A[17 * N + env] = touch_own ? 1.0 : 0.0;
      }
    }
  } else {
    // ---------------- EPMC tail (PGE:334-358, 360-372, 479-502)
    qb = qmul(qp, qI);
    float* snew = &s_new[threadIdx.x >> 2][0];
    const Q4 q1 = qnormalize(qb);
    const M3 Rq = qmat(q1);
#pragma unroll
    for (int i = 0; i < 3; i++) { snew[3 * k + i] = q[i]; snew[12 + 3 * k + i] = qd[i]; snew[kPropDim + 3 * k + i] = act[i]; }
    counter += 1;
    const double dx = tgx - px, dy = tgy - py;
    const double plen = sqrt(dx * dx + dy * dy);
    if (k == 0) {
      V3 wl = tmul(Rq, ww), vl = tmul(Rq, vw);
      snew[24] = wl.x; snew[25] = wl.y; snew[26] = wl.z; snew[27] = vl.x; snew[28] = vl.y; snew[29] = vl.z;
      snew[30] = Rq.a20; snew[31] = Rq.a21; snew[32] = Rq.a22;
      snew[45] = Rq.a00; snew[46] = Rq.a01; snew[47] = Rq.a02; snew[48] = Rq.a10; snew[49] = Rq.a11; snew[50] = Rq.a12;
      snew[51] = Rq.a20; snew[52] = Rq.a21; snew[53] = Rq.a22;
      snew[54] = (float)px; snew[55] = (float)py; snew[56] = (float)pz;
      V3 d = tmul(Rq, V3{(float)dx, (float)dy, (float)(0.0 - pz)});
      float n2 = sqrtf(d.x * d.x + d.y * d.y);
      snew[57] = d.x / n2; snew[58] = d.y / n2; snew[59] = target_spd;
      snew[60] = (float)sqrt(px * px + py * py + pz * pz);
    }
    const float left_z = Rq.a02 * Rq.a10 - Rq.a12 * Rq.a00;
    const bool fall = left_z > 0.70710678118654752f || left_z < -0.70710678118654752f || Rq.a22 < 0.5f;
    const bool reach = plen < 0.5, timeup = counter >= P.max_steps;
    const float ux = (float)(dx / plen), uy = (float)(dy / plen);
    const float spd = fabsf(vw.x * ux + vw.y * uy);
    total_spd += (double)spd;
    if ((double)spd > max_spd) max_spd = (double)spd;
    const float yaw = atan2f(Rq.a10, Rq.a00);
    float sy, cy;
    llq_sincosf(yaw, &sy, &cy);
    float rew = expf(-fabsf(spd - target_spd)) * expf((cy * ux + sy * uy - 1.0f) * 5.0f) / (float)P.max_steps;
    if (ENV == 3) {                                                    // _compute_avg_spd_reward (PGE:504-539)
      const float reward_rot = expf((cy * ux + sy * uy - 1.0f) * 5.0f);
      const float reward_dist = (float)((plen - last_len) / init_len);
      last_len = plen;
      rew = reward_rot / (float)P.max_steps * 0.1f * 2.0f - reward_dist * 0.1f;
      if (reach) rew += expf(-fabsf((float)(total_spd / (double)counter) - target_spd));
      stage_corridor_masks(snew, E.boxes + (size_t)env * (6 * kMaxBoxes), E.nbox[env], k, (float)px, (float)py, (float)pz, yaw);
    }
    {
      int bi = bad ? 1 : 0;
      bi |= __shfl_xor_sync(FULL, bi, 1);
      bi |= __shfl_xor_sync(FULL, bi, 2);
      bad = bi != 0;
    }
    if (bad || !isfinite(rew)) { rew = 0.f; bad = true; }
    done = fall || timeup || reach || bad;
    V3 fd;
    {
      V3 f = mul(qmat(qp), foot_in_base(L, q[0], q[1], q[2]));
      fd = V3{(float)px + f.x, (float)py + f.y, (float)pz + f.z};
    }
    if (valid) {
      float* sw = E.st;
#pragma unroll
      for (int i = 0; i < 3; i++) { sw[(10 + 3 * k + i) * N + env] = q[i]; sw[(22 + 3 * k + i) * N + env] = qd[i]; }
      E.warm[k * N + env] = warm_wheel ? -warm : warm;
      E.foot_pos[(3 * k) * N + env] = fd.x; E.foot_pos[(3 * k + 1) * N + env] = fd.y; E.foot_pos[(3 * k + 2) * N + env] = fd.z;
      if (k == 0) {
        E.pos[env] = px; E.pos[N + env] = py; E.pos[2 * N + env] = pz;
        sw[env] = qb.x; sw[N + env] = qb.y; sw[2 * N + env] = qb.z; sw[3 * N + env] = qb.w;
        sw[4 * N + env] = vw.x; sw[5 * N + env] = vw.y; sw[6 * N + env] = vw.z;
        sw[7 * N + env] = ww.x; sw[8 * N + env] = ww.y; sw[9 * N + env] = ww.z;
        E.time[env] = time;
        E.reward_sum[env] += rew;
        E.episode_steps[env] += 1;
        E.reward[env] = rew; rew_out = rew;
        E.done[env] = done ? 1 : 0;
        double* A = E.aux;
        A[env] = counter; A[N + env] = cmd_freq; A[2 * N + env] = tgx; A[3 * N + env] = tgy; A[4 * N + env] = target_spd;
        A[5 * N + env] = target_angle; A[6 * N + env] = last_len; A[7 * N + env] = total_spd; A[8 * N + env] = max_spd;
        A[9 * N + env] = push_count; A[10 * N + env] = pf[0]; A[11 * N + env] = pf[1]; A[12 * N + env] = pf[2];
        A[14 * N + env] = push_draws; A[15 * N + env] = cmd_draws;
      }
    }
  }
  // record mode (llq_set_option "record"): the trajectory columns action 12 | reward | done behind the observation of the slab row
  // (the kernel writes the whole record, no column copies afterwards); neglogp / value belong to the policy kernel
  // record == 2: into the slab row before the one that receives the observation (the [T+1, N, ld] layout of parallel/rollout.py, where
  // row t holds obs_t and this step's a_t | r_t | done_t while obs_{t+1} goes to row t+1)
  if (record && obs2 && valid) {
    float* row = obs2 + (size_t)env * obs2_ld + ObsW<ENV>::value - (record == 2 ? (long long)N * obs2_ld : 0ll);
#pragma unroll
    for (int i = 0; i < 3; i++) row[3 * k + i] = act[i];
    if (k == 0) { row[12] = rew_out; row[13] = done ? 1.f : 0.f; }
  }
  // counters: one atomic per warp
  {
    unsigned long long cr = n_contact_rows, lr = n_limit_rows;
    if (!valid) { cr = 0; lr = 0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { cr += __shfl_xor_sync(FULL, cr, o); lr += __shfl_xor_sync(FULL, lr, o); }
    unsigned dm = __ballot_sync(FULL, valid && k == 0 && done);
    if ((threadIdx.x & 31) == 0) {
      if (cr) atomicAdd(&E.counters[2], cr);
      if (lr) atomicAdd(&E.counters[3], lr);
      if (dm) atomicAdd(&E.counters[1], (unsigned long long)__popc(dm));
    }
  }
  // ---- observation rows of this warp (history shift + new prop / action / future), coalesced
  __pipeline_wait_prior(0);
  __syncwarp();
  const int warp_env0 = (blockIdx.x * BLOCK + (threadIdx.x & ~31)) >> 2;
  emit_obs_rows<ENV>(E.obs, obs2, obs2_ld, &s_new[(threadIdx.x & ~31) >> 2][0], &s_hist[(threadIdx.x & ~31) >> 2][0], warp_env0, N, 0, 0xFFu, E.boxes);
}

// ---------------------------------------------------------------------------------------------------------------
// Reset kernel (PLE:150-171, ML:48-63): also owns the prioritized-sampling table update (PLE:235-240).
// mode 0: reset envs with done[i] != 0 (auto-reset after a step), sampling clip/phase
// mode 1: reset envs with mask[i] != 0 (mask == null: all), sampling
// mode 2: like mode 1 but clip/time given
// mode 3: no env is reset (table update only; auto_reset off)
struct ResetParams {
  int mode; const unsigned char* mask; const int* clip_in; const double* time_in;
  unsigned long long seed; long long gid0;
  int* winner_cur; int* winner_next;          // [n_clips]
  const double* avg_old; double* avg_new;     // [n_clips]
  double* prob;                               // [n_clips]  (written by block 0)
  const double* max_steps;                    // [n_clips]
  double factor;
  int update_table;                           // 1 after a step
};

template <int BLOCK, int ENV>
__global__ void __launch_bounds__(BLOCK) pmc_reset_kernel(EnvArrays E, MocapDev mc, StepParams P, const ModelConst* __restrict__ gmodel,
                                                          ResetParams RP, float* obs2, long long obs2_ld) {
  extern __shared__ double s_cdf[];            // [n_clips]
  __shared__ __align__(16) ModelConst M;
  __shared__ __align__(16) float s_new[BLOCK / 4][kNewObs];
  prefetch_model(gmodel, &M, BLOCK);
  __pipeline_commit();
  const int C = mc.n_clips;
  // ---- prioritized sampling table: every block recomputes it identically; block 0 publishes it
  for (int c = threadIdx.x; c < C; c += BLOCK) {
    double avg = RP.avg_old[c];
    if (RP.update_table) {
      int w = RP.winner_cur[c];
      if (w >= 0) avg = (double)E.done_reward[w] / RP.max_steps[c];
    }
    s_cdf[c] = avg;
  }
  __syncthreads();
  if (RP.update_table && blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += BLOCK) { RP.avg_new[c] = s_cdf[c]; RP.winner_next[c] = -1; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += BLOCK) s_cdf[c] = pow(1.0 - s_cdf[c], RP.factor);
  __syncthreads();
  // p = w / sum(w), cdf = cumsum(p) / cumsum(p)[-1] exactly as np.random.choice builds them: the two sums run sequentially on one
  // thread (their order fixes the last bits), the 2 C fp64 divisions -- 3/4 of this section's latency when thread 0 did them one
  // after the other -- run one per thread
  __shared__ double s_tot;
  if (threadIdx.x == 0) {
    double tot = 0;
    for (int c = 0; c < C; c++) tot += s_cdf[c];
    s_tot = tot;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += BLOCK) {
    const double pc = s_cdf[c] / s_tot;
    s_cdf[c] = pc;
    if (blockIdx.x == 0) RP.prob[c] = pc;
  }
  __syncthreads();
  if (threadIdx.x == 0 && C > 0) {
    double acc = 0;
    for (int c = 0; c < C; c++) { acc += s_cdf[c]; s_cdf[c] = acc; }
    s_tot = s_cdf[C - 1];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += BLOCK) s_cdf[c] = s_cdf[c] / s_tot;
  __syncthreads();

  __pipeline_wait_prior(0);
  __syncthreads();
  const int N = P.n_envs;
  const int gtid = blockIdx.x * BLOCK + threadIdx.x;
  const int env_raw = gtid >> 2;
  const int env = env_raw < N ? env_raw : N - 1;
  const bool valid = env_raw < N;
  const int k = threadIdx.x & 3;
  bool doit = valid;
  if (RP.mode == 3) doit = false;
  else if (RP.mode == 0) doit = doit && E.done[env] != 0;
  else if (RP.mask) doit = doit && (RP.mask[env] != 0 || (ENV == 2 && RP.mask[env ^ 1] != 0));   // SEPMC: a pair resets as a whole
  const unsigned wm = __ballot_sync(FULL, doit);
  if (wm == 0) return;                                   // warp-uniform: nothing to reset in these 8 envs
  const LegConst& L = M.leg[k];

  if (ENV == 2) {
    // ---------------- SEPMC reset (CTG:261-304, 204-230); draws keyed by the pair: stream 1 = [fix_spd, with_flag, friction, x0 |
    // y0, x1, y1, yaw0 | yaw1, flag x, flag y]
    const int robot = env & 1;
    const long long ep = E.episode[env];
    const long long gid = RP.gid0 + (env & ~1);
    double u0[4], u1[4], u2[4];
    stream_uniforms(RP.seed, gid, ep, 1, 0, u0);
    stream_uniforms(RP.seed, gid, ep, 1, 1, u1);
    stream_uniforms(RP.seed, gid, ep, 1, 2, u2);
    const float fix_spd = (float)(0.5 + 2.5 * u0[0]);
    const int wflag = (int)floor(2.0 * u0[1]);
    const double foot_mu = (double)P.fr_lo + u0[2] * ((double)P.fr_hi - (double)P.fr_lo);
    const double px = robot == 0 ? -2.0 + 4.0 * u0[3] : -2.0 + 4.0 * u1[1], py = robot == 0 ? -2.0 + 4.0 * u1[0] : -2.0 + 4.0 * u1[2];
    // both robots are handed the same mutable init dict => one running yaw for the pair (CTG:209-215)
    const double acc0 = E.aux[16 * N + (env & ~1)];
    const double yaw_a = fmod(acc0 + 360.0 * u1[3], 360.0), yaw_b = fmod(yaw_a + 360.0 * u2[0], 360.0);
    const double yaw_deg = robot == 0 ? yaw_a : yaw_b;
    double sn, cs;
    sincos(0.5 * yaw_deg * (3.14159265358979323846 / 180.0), &sn, &cs);
    const float* I0 = M.init_state;
    const Q4 qn = qmul(qnormalize(Q4{I0[3], I0[4], I0[5], I0[6]}), Q4{0.f, 0.f, (float)sn, (float)cs});
    float q[3], qd[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { q[i] = I0[13 + 3 * k + i]; qd[i] = I0[25 + 3 * k + i]; }
    const V3 lin = V3{I0[7], I0[8], I0[9]}, ang = V3{I0[10], I0[11], I0[12]};
    const Q4 qI = Q4{M.base.qI[0], M.base.qI[1], M.base.qI[2], M.base.qI[3]};
    const Q4 qp = qmul(qnormalize(qn), qconj(qI));
    PairState PS = {robot == 0 ? wflag : 1 - wflag, 0, 1, 0, -2.0 + 4.0 * u2[1], -2.0 + 4.0 * u2[2]};
    // reset() runs _prepare_drill too (CTG:302): its flag-switch test reads the stale manifolds of the previous episode's last step
    const bool touch_own = E.aux[17 * N + env] != 0.0;
    float* snew = &s_new[threadIdx.x >> 2][0];
    const float* spart = &s_new[(threadIdx.x >> 2) ^ 1][0];
    sepmc_pair_tail(M, L, k, robot, snew, spart, px, py, 0.5, qp, qn, lin, ang, q, touch_own, fix_spd, RP.seed, gid, ep, PS);
#pragma unroll
    for (int i = 0; i < 3; i++) { snew[3 * k + i] = q[i]; snew[12 + 3 * k + i] = qd[i]; }
    int push_draws = 0;
    float pf[3] = {0.f, 0.f, 0.f};
    if (P.push_enabled) push_draws = 1;                              // PR:52-54: draw #0 becomes the current _randomized_force
    if (doit) {
      float* sw = E.st;
      V3 f = mul(qmat(qp), foot_in_base(L, q[0], q[1], q[2]));
#pragma unroll
      for (int i = 0; i < 3; i++) { sw[(10 + 3 * k + i) * N + env] = q[i]; sw[(22 + 3 * k + i) * N + env] = qd[i]; }
      E.warm[k * N + env] = 0.f;
      E.foot_pos[(3 * k) * N + env] = (float)px + f.x; E.foot_pos[(3 * k + 1) * N + env] = (float)py + f.y; E.foot_pos[(3 * k + 2) * N + env] = 0.5f + f.z;
      if (k == 0) {
        E.pos[env] = px; E.pos[N + env] = py; E.pos[2 * N + env] = 0.5;
        float b[10] = {qn.x, qn.y, qn.z, qn.w, lin.x, lin.y, lin.z, ang.x, ang.y, ang.z};
#pragma unroll
        for (int i = 0; i < 10; i++) sw[i * N + env] = b[i];
        E.time[env] = 0.0; E.reward_sum[env] = 0.f; E.episode_steps[env] = 0; E.episode[env] = ep + 1;
        double* A = E.aux;
        A[env] = 0; A[N + env] = PS.with_flag; A[2 * N + env] = PS.flag_x; A[3 * N + env] = PS.flag_y; A[4 * N + env] = fix_spd;
        A[5 * N + env] = PS.visible; A[6 * N + env] = PS.sw; A[7 * N + env] = 0.0; A[8 * N + env] = 0.0; A[9 * N + env] = P.push_start_count;
        A[10 * N + env] = pf[0]; A[11 * N + env] = pf[1]; A[12 * N + env] = pf[2]; A[13 * N + env] = foot_mu; A[14 * N + env] = push_draws;
        A[15 * N + env] = PS.flag_draws; A[16 * N + env] = yaw_b; A[17 * N + env] = touch_own ? 1.0 : 0.0;
      }
    }
  } else if (ENV == 1 || ENV == 3) {
    // ---------------- EPMC reset (PGE:196-249)
    long long ep = E.episode[env];
    const long long gid = RP.gid0 + env;
    double u[4];
    stream_uniforms(RP.seed, gid, ep, 1, 0, u);
    const double foot_mu = (double)P.fr_lo + u[0] * ((double)P.fr_hi - (double)P.fr_lo);                   // PGE:209-210
    int push_draws = 0;
    float pf[3] = {0.f, 0.f, 0.f};
    if (P.push_enabled) epmc_randomize_push(P, RP.seed, gid, ep, push_draws, pf);                          // PR:52-54
    const int cmd_freq = P.cmd_freq_lo + (int)floor(u[2] * (double)(P.cmd_freq_hi - P.cmd_freq_lo));       // PGE:223
    const double yaw_deg = fmod(E.aux[16 * N + env] + 360.0 * u[1], 360.0);                                // PGE:181-189 (accumulates)
    double sn, cs;
    sincos(0.5 * yaw_deg * (3.14159265358979323846 / 180.0), &sn, &cs);
    const float* I0 = M.init_state;
    const Q4 qn = qmul(qnormalize(Q4{I0[3], I0[4], I0[5], I0[6]}), Q4{0.f, 0.f, (float)sn, (float)cs});
    float q[3], qd[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { q[i] = I0[13 + 3 * k + i]; qd[i] = I0[25 + 3 * k + i]; }
    const V3 lin = V3{I0[7], I0[8], I0[9]}, ang = V3{I0[10], I0[11], I0[12]};
    float* snew = &s_new[threadIdx.x >> 2][0];
    const M3 Rq = qmat(qnormalize(qn));
    const float target_spd = (float)E.aux[4 * N + env];            // persists across episodes (PGE:170-172)
    double tgx0 = 8.0;
    int nb0 = 0;
    if (ENV == 3) nb0 = generate_corridor(P, RP.seed, gid, ep, E.boxes + (size_t)env * (6 * kMaxBoxes), doit && k == 0, tgx0);   // PGE:216-219
#pragma unroll
    for (int i = 0; i < 3; i++) { snew[3 * k + i] = q[i]; snew[12 + 3 * k + i] = qd[i]; }
    if (k == 0) {
      V3 wl = tmul(Rq, ang), vl = tmul(Rq, lin);
      snew[24] = wl.x; snew[25] = wl.y; snew[26] = wl.z; snew[27] = vl.x; snew[28] = vl.y; snew[29] = vl.z;
      snew[30] = Rq.a20; snew[31] = Rq.a21; snew[32] = Rq.a22;
      snew[45] = Rq.a00; snew[46] = Rq.a01; snew[47] = Rq.a02; snew[48] = Rq.a10; snew[49] = Rq.a11; snew[50] = Rq.a12;
      snew[51] = Rq.a20; snew[52] = Rq.a21; snew[53] = Rq.a22;
      snew[54] = 0.f; snew[55] = 0.f; snew[56] = 0.5f;
      V3 d = tmul(Rq, V3{(float)tgx0, 0.f, -0.5f});                 // target - pos (0,0,0.5); element 0: (8,0,0) (BSE:247-248)
      float n2 = sqrtf(d.x * d.x + d.y * d.y);
      snew[57] = d.x / n2; snew[58] = d.y / n2; snew[59] = target_spd;
      snew[60] = 0.5f;
    }
    if (ENV == 3) {
      __syncwarp();                                                 // lane 0's boxes are visible to the env's other lanes
      stage_corridor_masks(snew, E.boxes + (size_t)env * (6 * kMaxBoxes), doit ? nb0 : E.nbox[env], k, 0.f, 0.f, 0.5f, atan2f(Rq.a10, Rq.a00));
    }
    if (doit) {
      float* sw = E.st;
      const Q4 qI = Q4{M.base.qI[0], M.base.qI[1], M.base.qI[2], M.base.qI[3]};
      V3 f = mul(qmat(qmul(qnormalize(qn), qconj(qI))), foot_in_base(L, q[0], q[1], q[2]));
#pragma unroll
      for (int i = 0; i < 3; i++) { sw[(10 + 3 * k + i) * N + env] = q[i]; sw[(22 + 3 * k + i) * N + env] = qd[i]; }
      E.warm[k * N + env] = 0.f;
      E.foot_pos[(3 * k) * N + env] = f.x; E.foot_pos[(3 * k + 1) * N + env] = f.y; E.foot_pos[(3 * k + 2) * N + env] = 0.5f + f.z;
      if (k == 0) {
        E.pos[env] = 0.0; E.pos[N + env] = 0.0; E.pos[2 * N + env] = 0.5;
        float b[10] = {qn.x, qn.y, qn.z, qn.w, lin.x, lin.y, lin.z, ang.x, ang.y, ang.z};
#pragma unroll
        for (int i = 0; i < 10; i++) sw[i * N + env] = b[i];
        E.time[env] = 0.0; E.reward_sum[env] = 0.f; E.episode_steps[env] = 0; E.episode[env] = ep + 1;
        double* A = E.aux;
        A[env] = 0; A[N + env] = cmd_freq; A[2 * N + env] = tgx0; A[3 * N + env] = 0.0; A[6 * N + env] = fabs(tgx0); A[7 * N + env] = 0.0;
        A[17 * N + env] = fabs(tgx0);                                  // init_pos_diff_len (PGE:192-195)
        if (ENV == 3) E.nbox[env] = nb0;
        A[8 * N + env] = 0.0; A[9 * N + env] = P.push_start_count; A[10 * N + env] = pf[0]; A[11 * N + env] = pf[1]; A[12 * N + env] = pf[2];
        A[13 * N + env] = foot_mu; A[14 * N + env] = push_draws; A[15 * N + env] = 0; A[16 * N + env] = yaw_deg;
      }
    }
  } else {
  int clip; double t0;
  long long ep = E.episode[env];
  if (RP.mode == 2) {
    clip = RP.clip_in[env]; t0 = RP.time_in[env];
    if (!doit) { clip = 0; t0 = 0.0; }     // entries of masked-out envs are not validated by the host: never index the table with them
  } else {
    long long gid = RP.gid0 + env;
    unsigned c4[4] = {(unsigned)gid, (unsigned)((unsigned long long)gid >> 32), (unsigned)ep, (unsigned)((unsigned long long)ep >> 32)};
    philox4x32_10(c4, (unsigned)RP.seed, (unsigned)(RP.seed >> 32));
    double u1 = ((double)c4[0] + 0.5) * (1.0 / 4294967296.0), u2 = ((double)c4[1] + 0.5) * (1.0 / 4294967296.0);
    clip = C - 1;
    for (int c = 0; c < C; c++) if (s_cdf[c] > u1) { clip = c; break; }
    int nf = mc.clip_off[clip + 1] - mc.clip_off[clip];
    t0 = u2 * (P.frame_dt * (double)(nf - P.margin - 1));
    ep += 1;
  }
  int frame_id = (int)floor(t0 / P.frame_dt);
  double frac = (t0 - frame_id * P.frame_dt) / P.frame_dt;
  {
    const int last = mc.clip_off[clip + 1] - mc.clip_off[clip] - P.margin + 2;
    if (frame_id > last) { frame_id = last; frac = 0.0; }
    if (frame_id < 0) { frame_id = 0; frac = 0.0; }
  }
  const MocapFrame* f0 = mc.frames + mc.clip_off[clip] + frame_id;
  KinBase kb = mocap_base(f0, f0 + 1, frac, P.frame_dt);
  float inv = (float)(1.0 / P.frame_dt), fr = (float)frac;
  float q[3], qd[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float c = f0->q[3 * k + i], n = f0[1].q[3 * k + i];
    q[i] = fmaf(fr, n - c, c);
    qd[i] = (n - c) * inv;
  }
  float* snew = &s_new[threadIdx.x >> 2][0];
  build_obs_new(mc, P, M, k, clip, frame_id, frac, kb.px, kb.py, kb.pz, kb.q, kb.lin, kb.ang, q, qd, snew);
  if (doit) {
    float* sw = E.st;
    const Q4 qI = Q4{M.base.qI[0], M.base.qI[1], M.base.qI[2], M.base.qI[3]};
    V3 f = mul(qmat(qmul(qnormalize(kb.q), qconj(qI))), foot_in_base(L, q[0], q[1], q[2]));
#pragma unroll
    for (int i = 0; i < 3; i++) {
      sw[(10 + 3 * k + i) * N + env] = q[i]; sw[(22 + 3 * k + i) * N + env] = qd[i];
      E.kin[(13 + 3 * k + i) * N + env] = q[i]; E.kin[(25 + 3 * k + i) * N + env] = qd[i];
    }
    E.warm[k * N + env] = 0.f;
    E.foot_pos[(3 * k) * N + env] = (float)kb.px + f.x; E.foot_pos[(3 * k + 1) * N + env] = (float)kb.py + f.y;
    E.foot_pos[(3 * k + 2) * N + env] = (float)kb.pz + f.z;
    if (k == 0) {
      E.pos[env] = kb.px; E.pos[N + env] = kb.py; E.pos[2 * N + env] = kb.pz;
      float b[13] = {kb.q.x, kb.q.y, kb.q.z, kb.q.w, kb.lin.x, kb.lin.y, kb.lin.z, kb.ang.x, kb.ang.y, kb.ang.z, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 10; i++) sw[i * N + env] = b[i];
      E.kin[env] = (float)kb.px; E.kin[N + env] = (float)kb.py; E.kin[2 * N + env] = (float)kb.pz;
#pragma unroll
      for (int i = 0; i < 10; i++) E.kin[(3 + i) * N + env] = b[i];
      E.time[env] = t0; E.clip[env] = clip; E.reward_sum[env] = 0.f; E.episode_steps[env] = 0; E.episode[env] = ep;
      E.ob_id[env] = 0;                                            // PLE:179
    }
  }
  }
  __syncwarp();
  unsigned rows = 0;
#pragma unroll
  for (int e = 0; e < 8; e++) if ((wm >> (4 * e)) & 1u) rows |= 1u << e;
  const int warp_env0 = (blockIdx.x * BLOCK + (threadIdx.x & ~31)) >> 2;
  emit_obs_rows<ENV>(E.obs, obs2, obs2_ld, &s_new[(threadIdx.x & ~31) >> 2][0], &s_new[0][0], warp_env0, N, 1, rows, E.boxes);
}

}  // namespace llq
