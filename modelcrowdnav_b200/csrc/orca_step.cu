// orca_step.cu -- ORCA human motion (K1), CrowdSim.step (K2), device reset, AoS<->SoA packing.
//
// Compiled with -fmad=false: the reference arithmetic (x86 rvo2 in float32, CPython in float64)
// never contracts mul+add, and collision/done codes must be bit-exact.  The only fused operation
// on the reference path is numpy's 2-element dot inside np.linalg.norm (norm2d, explicit fma()).
//
// Reference (file:line relative to the reference root):
//   ORCA.predict                     crowd_sim/envs/policy/orca.py:82-132
//   rvo2 Agent::computeNeighbors / computeNewVelocity / linearProgram1-3 (third-party RVO2 2.0.x)
//   CrowdSim.step                    crowd_sim/envs/crowd_sim.py:331-434
//   point_to_segment_dist            crowd_sim/envs/utils/utils.py:4-26
//   CrowdSim.reset + generators      crowd_sim/envs/crowd_sim.py:165-217,261-323
//   Explorer counters                crowd_nav/utils/explorer.py:41-51,92-141
#include "env_math.cuh"

#include <math.h>

#define RVO_EPSILON 0.00001f

namespace {

struct v2 { float x, y; };
struct Line { v2 point, direction; };

__device__ __forceinline__ v2 V(float x, float y) { v2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ v2 vneg(v2 a) { return V(-a.x, -a.y); }
__device__ __forceinline__ v2 vscale(float s, v2 a) { return V(s * a.x, s * a.y); }
__device__ __forceinline__ float vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float vdet(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float vabssq(v2 a) { return vdot(a, a); }
__device__ __forceinline__ float vabs(v2 a) { return sqrtf(vdot(a, a)); }
// RVO2 Vector2::operator/(float) multiplies by the reciprocal
__device__ __forceinline__ v2 vdiv(v2 a, float s) { const float inv = 1.0f / s; return V(a.x * inv, a.y * inv); }
__device__ __forceinline__ v2 vnormalize(v2 a) { return vdiv(a, vabs(a)); }
__device__ __forceinline__ float sqrf(float a) { return a * a; }

// linearProgram1: optimise along line `lineNo` subject to lines [0, lineNo) and the speed disc.
__device__ bool lp1(const Line *lines, int lineNo, float radius, v2 opt, bool directionOpt, v2 &result)
{
    const Line L = lines[lineNo];
    const float dotProduct = vdot(L.point, L.direction);
    const float discriminant = sqrf(dotProduct) + sqrf(radius) - vabssq(L.point);
    if (discriminant < 0.0f) return false;

    const float sqrtDiscriminant = sqrtf(discriminant);
    float tLeft = -dotProduct - sqrtDiscriminant;
    float tRight = -dotProduct + sqrtDiscriminant;

    for (int i = 0; i < lineNo; ++i) {
        const Line Li = lines[i];
        const float denominator = vdet(L.direction, Li.direction);
        const float numerator = vdet(Li.direction, vsub(L.point, Li.point));
        if (fabsf(denominator) <= RVO_EPSILON) {
            if (numerator < 0.0f) return false;
            continue;
        }
        const float t = numerator / denominator;
        if (denominator >= 0.0f) tRight = fminf(tRight, t);
        else tLeft = fmaxf(tLeft, t);
        if (tLeft > tRight) return false;
    }

    if (directionOpt) {
        if (vdot(opt, L.direction) > 0.0f) result = vadd(L.point, vscale(tRight, L.direction));
        else result = vadd(L.point, vscale(tLeft, L.direction));
    } else {
        const float t = vdot(L.direction, vsub(opt, L.point));
        if (t < tLeft) result = vadd(L.point, vscale(tLeft, L.direction));
        else if (t > tRight) result = vadd(L.point, vscale(tRight, L.direction));
        else result = vadd(L.point, vscale(t, L.direction));
    }
    return true;
}

__device__ int lp2(const Line *lines, int n, float radius, v2 opt, bool directionOpt, v2 &result)
{
    if (directionOpt) result = vscale(radius, opt);
    else if (vabssq(opt) > sqrf(radius)) result = vscale(radius, vnormalize(opt));
    else result = opt;

    for (int i = 0; i < n; ++i) {
        if (vdet(lines[i].direction, vsub(lines[i].point, result)) > 0.0f) {
            const v2 tmp = result;
            if (!lp1(lines, i, radius, opt, directionOpt, result)) {
                result = tmp;
                return i;
            }
        }
    }
    return n;
}

__device__ void lp3(const Line *lines, int n, int beginLine, float radius, v2 &result)
{
    float distance = 0.0f;
    Line proj[CN_MAX_NEIGHBORS];
    for (int i = beginLine; i < n; ++i) {
        if (vdet(lines[i].direction, vsub(lines[i].point, result)) > distance) {
            int np = 0;
            const Line Li = lines[i];
            for (int j = 0; j < i; ++j) {
                const Line Lj = lines[j];
                Line line;
                const float determinant = vdet(Li.direction, Lj.direction);
                if (fabsf(determinant) <= RVO_EPSILON) {
                    if (vdot(Li.direction, Lj.direction) > 0.0f) continue;
                    line.point = vscale(0.5f, vadd(Li.point, Lj.point));
                } else {
                    const float t = vdet(Lj.direction, vsub(Li.point, Lj.point)) / determinant;
                    line.point = vadd(Li.point, vscale(t, Li.direction));
                }
                line.direction = vnormalize(vsub(Lj.direction, Li.direction));
                proj[np++] = line;
            }
            const v2 tmp = result;
            if (lp2(proj, np, radius, V(-Li.direction.y, Li.direction.x), true, result) < np) result = tmp;
            distance = vdet(Li.direction, vsub(Li.point, result));
        }
    }
}

// One ORCA solve: `self` against the candidate list cand[0..n_cand) of agent indices, visited in
// list order (rvo2 kd-tree single-leaf order = addAgent order, orca.py:100-104).
__device__ void orca_solve(const EnvParams &p, const double *__restrict__ st, int e, int self,
                           const int *cand_first, int n_first, int extra_agent /* -1 or robot */,
                           double safety_space, float &out_vx, float &out_vy)
{
    const EnvDims d = p.d;
    auto ldf = [&](int field, int agent) { return (float)st[st_idx(d, field, agent, e)]; };
    const v2 pos = V(ldf(F_PX, self), ldf(F_PY, self));
    const v2 vel = V(ldf(F_VX, self), ldf(F_VY, self));
    const float radius_self = (float)(st[st_idx(d, F_R, self, e)] + 0.01 + safety_space);
    const float max_speed = (float)st[st_idx(d, F_VPREF, self, e)];
    const v2 pref = V((float)(st[st_idx(d, F_GX, self, e)] - st[st_idx(d, F_PX, self, e)]),
                      (float)(st[st_idx(d, F_GY, self, e)] - st[st_idx(d, F_PY, self, e)]));

    // Agent::computeNeighbors / insertAgentNeighbor
    float nb_d2[CN_MAX_NEIGHBORS];
    int nb_id[CN_MAX_NEIGHBORS];
    int nn = 0;
    const int maxN = p.max_neighbors;
    float rangeSq = sqrf(p.neighbor_dist);
    const int n_cand = n_first + (extra_agent >= 0 ? 1 : 0);
    // positions: one base pointer, the agent stride is the env count (st[(field * A1 + agent) * E + env])
    const double *__restrict__ pxs = st + st_idx(d, F_PX, 0, e), *__restrict__ pys = st + st_idx(d, F_PY, 0, e);
    if (maxN == 10) {
        // The reference's max_neighbors (orca.py:55-67): the list lives in REGISTERS, kept sorted by a branch-free insertion.
        // Same result as the loop below: empty slots hold +inf, so "list not full or closer than its last entry" is one strict
        // comparison with slot 9 (rangeSq only ever shrinks from neighbor_dist^2 to that entry), and the new entry lands
        // behind every entry <= it (the strict `<` of insertAgentNeighbor: ties keep visiting order), pushing the rest down.
        float d0 = INFINITY, d1 = INFINITY, d2 = INFINITY, d3 = INFINITY, d4 = INFINITY, d5 = INFINITY, d6 = INFINITY,
              d7 = INFINITY, d8 = INFINITY, d9 = INFINITY;
        int i0 = 0, i1 = 0, i2 = 0, i3 = 0, i4 = 0, i5 = 0, i6 = 0, i7 = 0, i8 = 0, i9 = 0;
        const float nd2 = rangeSq;
        for (int c = 0; c < n_cand; ++c) {
            int o;
            if (c < n_first) { o = cand_first[0] + c; if (o >= self && cand_first[1]) ++o; }
            else o = extra_agent;
            const size_t oo = (size_t)o * d.E;
            const float distSq = vabssq(vsub(pos, V((float)pxs[oo], (float)pys[oo])));
            if (distSq < nd2 && distSq < d9) {
                if (nn < 10) ++nn;
#define CN_INS(hi, lo) { const bool sh = distSq < d##lo; const bool here = distSq < d##hi; \
                         i##hi = sh ? i##lo : (here ? o : i##hi); d##hi = sh ? d##lo : (here ? distSq : d##hi); }
                CN_INS(9, 8) CN_INS(8, 7) CN_INS(7, 6) CN_INS(6, 5) CN_INS(5, 4) CN_INS(4, 3) CN_INS(3, 2) CN_INS(2, 1) CN_INS(1, 0)
#undef CN_INS
                if (distSq < d0) { d0 = distSq; i0 = o; }
            }
        }
        nb_id[0] = i0; nb_id[1] = i1; nb_id[2] = i2; nb_id[3] = i3; nb_id[4] = i4;
        nb_id[5] = i5; nb_id[6] = i6; nb_id[7] = i7; nb_id[8] = i8; nb_id[9] = i9;
    } else if (maxN > 0) {
        for (int c = 0; c < n_cand; ++c) {
            int o;
            if (c < n_first) { o = cand_first[0] + c; if (o >= self && cand_first[1]) ++o; }
            else o = extra_agent;
            const float distSq = vabssq(vsub(pos, V(ldf(F_PX, o), ldf(F_PY, o))));
            if (distSq < rangeSq) {
                if (nn < maxN) { nb_d2[nn] = distSq; nb_id[nn] = o; ++nn; }
                int i = nn - 1;
                while (i != 0 && distSq < nb_d2[i - 1]) {
                    nb_d2[i] = nb_d2[i - 1]; nb_id[i] = nb_id[i - 1];
                    --i;
                }
                nb_d2[i] = distSq; nb_id[i] = o;
                if (nn == maxN) rangeSq = nb_d2[nn - 1];
            }
        }
    }

    // Agent::computeNewVelocity
    Line lines[CN_MAX_NEIGHBORS];
    const float invTimeHorizon = 1.0f / p.time_horizon;
    for (int k = 0; k < nn; ++k) {
        const int o = nb_id[k];
        const v2 relativePosition = vsub(V(ldf(F_PX, o), ldf(F_PY, o)), pos);
        const v2 relativeVelocity = vsub(vel, V(ldf(F_VX, o), ldf(F_VY, o)));
        const float distSq = vabssq(relativePosition);
        const float radius_o = (float)(st[st_idx(d, F_R, o, e)] + 0.01 + safety_space);
        const float combinedRadius = radius_self + radius_o;
        const float combinedRadiusSq = sqrf(combinedRadius);
        Line line;
        v2 u;
        if (distSq > combinedRadiusSq) {
            const v2 w = vsub(relativeVelocity, vscale(invTimeHorizon, relativePosition));
            const float wLengthSq = vabssq(w);
            const float dotProduct1 = vdot(w, relativePosition);
            if (dotProduct1 < 0.0f && sqrf(dotProduct1) > combinedRadiusSq * wLengthSq) {
                const float wLength = sqrtf(wLengthSq);
                const v2 unitW = vdiv(w, wLength);
                line.direction = V(unitW.y, -unitW.x);
                u = vscale(combinedRadius * invTimeHorizon - wLength, unitW);
            } else {
                const float leg = sqrtf(distSq - combinedRadiusSq);
                if (vdet(relativePosition, w) > 0.0f) {
                    line.direction = vdiv(V(relativePosition.x * leg - relativePosition.y * combinedRadius,
                                            relativePosition.x * combinedRadius + relativePosition.y * leg), distSq);
                } else {
                    line.direction = vneg(vdiv(V(relativePosition.x * leg + relativePosition.y * combinedRadius,
                                                 -relativePosition.x * combinedRadius + relativePosition.y * leg), distSq));
                }
                const float dotProduct2 = vdot(relativeVelocity, line.direction);
                u = vsub(vscale(dotProduct2, line.direction), relativeVelocity);
            }
        } else {
            const float invTimeStep = 1.0f / p.time_step_f;
            const v2 w = vsub(relativeVelocity, vscale(invTimeStep, relativePosition));
            const float wLength = vabs(w);
            const v2 unitW = vdiv(w, wLength);
            line.direction = V(unitW.y, -unitW.x);
            u = vscale(combinedRadius * invTimeStep - wLength, unitW);
        }
        line.point = vadd(vel, vscale(0.5f, u));
        lines[k] = line;
    }

    v2 newV;
    const int lineFail = lp2(lines, nn, max_speed, pref, false, newV);
    if (lineFail < nn) lp3(lines, nn, lineFail, max_speed, newV);
    out_vx = newV.x;
    out_vy = newV.y;
}

// K1: one thread per (human, env); consecutive threads = consecutive envs (coalesced SoA reads).
__global__ void __launch_bounds__(128) orca_humans_kernel(EnvParams p, const double *__restrict__ st,
                                                          const uint8_t *__restrict__ frozen,
                                                          double *__restrict__ human_v)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int E = p.d.E, H = p.d.H;
    if (tid >= E * H) return;
    const int h = tid / E, e = tid - h * E;
    if (frozen[e]) return;
    // others = humans 1..H except self, in list order (crowd_sim.py:339), robot last if visible (:340-341)
    const int cand[2] = {1, 1};
    float vx, vy;
    orca_solve(p, st, e, h + 1, cand, H - 1, p.robot_visible ? 0 : -1, p.human_safety_space, vx, vy);
    human_v[(size_t)(0 * H + h) * E + e] = (double)vx;
    human_v[(size_t)(1 * H + h) * E + e] = (double)vy;
}

// Robot with an ORCA policy (imitation learning, train.py:157-166): self = robot, others = all humans.
__global__ void __launch_bounds__(128) orca_robot_kernel(EnvParams p, const double *__restrict__ st,
                                                         const uint8_t *__restrict__ frozen, double safety_space,
                                                         double *__restrict__ action_xy,
                                                         int32_t *__restrict__ action_idx)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.d.E || frozen[e]) return;
    const int cand[2] = {1, 0};
    float vx, vy;
    orca_solve(p, st, e, 0, cand, p.d.H, -1, safety_space, vx, vy);
    action_xy[e] = (double)vx;
    action_xy[p.d.E + e] = (double)vy;
    action_idx[e] = -1;
}

constexpr int kWarpResetHumans = 16;     // from this human count on, device resets run one warp per env (reset_env_warp)

// K2: CrowdSim.step for one env per thread (crowd_sim.py:344-434).
// CrowdSim.reset on the device for ONE env: crowd_sim.py:165-217 distributions and rejection rule; Philox stream
// keyed by (seed, global env id, episode counter).
__device__ void reset_env(const EnvParams &p, int e, double *__restrict__ st, double *__restrict__ time,
                          uint8_t *__restrict__ frozen, const EnvAccum &acc, double *__restrict__ theta)
{
    const EnvDims d = p.d;
    PhiloxStream rng;
    rng.init(p.seed, (uint64_t)(p.env_id_offset + e), acc.episode_ctr[e], 0u);
    acc.episode_ctr[e] += 1;
    const double PI = 3.141592653589793;
    // robot (crowd_sim.py:284)
    st[st_idx(d, F_PX, 0, e)] = 0.0; st[st_idx(d, F_PY, 0, e)] = -p.circle_radius;
    st[st_idx(d, F_VX, 0, e)] = 0.0; st[st_idx(d, F_VY, 0, e)] = 0.0;
    st[st_idx(d, F_GX, 0, e)] = 0.0; st[st_idx(d, F_GY, 0, e)] = p.circle_radius;
    st[st_idx(d, F_R, 0, e)] = p.robot_radius; st[st_idx(d, F_VPREF, 0, e)] = p.robot_v_pref;
    // the agents placed so far, in thread-local arrays (L1) instead of being read back from global memory on every check of the
    // rejection sampling (this path serves crowds below kWarpResetHumans; larger ones use reset_env_warp)
    constexpr int kLocal = 16;
    double cpx[kLocal], cpy[kLocal], cgx[kLocal], cgy[kLocal], cr[kLocal];
    cpx[0] = 0.0; cpy[0] = -p.circle_radius; cgx[0] = 0.0; cgy[0] = p.circle_radius; cr[0] = p.robot_radius;
    const bool cached = d.H < kLocal;
    auto PX = [&](int a) { return cached ? cpx[a] : st[st_idx(d, F_PX, a, e)]; };
    auto PY = [&](int a) { return cached ? cpy[a] : st[st_idx(d, F_PY, a, e)]; };
    auto GX = [&](int a) { return cached ? cgx[a] : st[st_idx(d, F_GX, a, e)]; };
    auto GY = [&](int a) { return cached ? cgy[a] : st[st_idx(d, F_GY, a, e)]; };
    auto RR = [&](int a) { return cached ? cr[a] : st[st_idx(d, F_R, a, e)]; };
    const int MAX_TRIES = 4096;
    for (int i = 1; i <= d.H; ++i) {
        double px = 0, py = 0, gx = 0, gy = 0;
        double h_radius = p.human_radius, h_v_pref = p.human_v_pref;
        if (p.randomize_attributes) {                       // agent.py:39-45, drawn before the position (crowd_sim.py:167-168)
            h_v_pref = 0.5 + (1.5 - 0.5) * rng.next();
            h_radius = 0.3 + (0.5 - 0.3) * rng.next();
        }
        if (p.sim_rule == CN_CIRCLE_CROSSING) {
            for (int tries = 0; tries < MAX_TRIES; ++tries) {
                const double angle = rng.next() * PI * 2;
                const double px_noise = (rng.next() - 0.5) * h_v_pref;
                const double py_noise = (rng.next() - 0.5) * h_v_pref;
                px = p.circle_radius * cos(angle) + px_noise;
                py = p.circle_radius * sin(angle) + py_noise;
                bool collide = false;
                for (int a = 0; a < i; ++a) {
                    const double min_dist = h_radius + RR(a) + p.discomfort_dist;
                    if (norm2d(px - PX(a), py - PY(a)) < min_dist || norm2d(px - GX(a), py - GY(a)) < min_dist) {
                        collide = true;
                        break;
                    }
                }
                if (!collide) break;
            }
            gx = -px; gy = -py;
        } else {
            const double sign = (rng.next() > 0.5) ? -1.0 : 1.0;
            for (int tries = 0; tries < MAX_TRIES; ++tries) {
                px = rng.next() * p.square_width * 0.5 * sign;
                py = (rng.next() - 0.5) * p.square_width;
                bool collide = false;
                for (int a = 0; a < i; ++a) {
                    if (norm2d(px - PX(a), py - PY(a)) < h_radius + RR(a) + p.discomfort_dist) { collide = true; break; }
                }
                if (!collide) break;
            }
            for (int tries = 0; tries < MAX_TRIES; ++tries) {
                gx = rng.next() * p.square_width * 0.5 * -sign;
                gy = (rng.next() - 0.5) * p.square_width;
                bool collide = false;
                for (int a = 0; a < i; ++a) {
                    if (norm2d(gx - GX(a), gy - GY(a)) < h_radius + RR(a) + p.discomfort_dist) { collide = true; break; }
                }
                if (!collide) break;
            }
        }
        st[st_idx(d, F_PX, i, e)] = px; st[st_idx(d, F_PY, i, e)] = py;
        st[st_idx(d, F_VX, i, e)] = 0.0; st[st_idx(d, F_VY, i, e)] = 0.0;
        st[st_idx(d, F_GX, i, e)] = gx; st[st_idx(d, F_GY, i, e)] = gy;
        st[st_idx(d, F_R, i, e)] = h_radius; st[st_idx(d, F_VPREF, i, e)] = h_v_pref;
        if (cached) { cpx[i] = px; cpy[i] = py; cgx[i] = gx; cgy[i] = gy; cr[i] = h_radius; }
    }
    time[e] = 0.0;
    theta[e] = 1.5707963267948966;                       // crowd_sim.py:284: robot.set(..., np.pi / 2)
    frozen[e] = 0;
    acc.ep_steps[e] = 0; acc.ep_return[e] = 0.0;
}

__global__ void __launch_bounds__(128) step_kernel(EnvParams p, double *__restrict__ st, double *__restrict__ time,
                                                   const double *__restrict__ human_v,
                                                   const double *__restrict__ act, int act_aos, int update,
                                                   double *__restrict__ reward_o, uint8_t *__restrict__ done_o,
                                                   uint8_t *__restrict__ info_o, double *__restrict__ dmin_o,
                                                   double *__restrict__ next_obs, uint8_t *__restrict__ frozen,
                                                   EnvAccum acc, double *__restrict__ theta, int fuse_reset)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const EnvDims d = p.d;
    if (e >= d.E) return;
    if (frozen[e]) { reward_o[e] = 0.0; done_o[e] = 1; return; }
    const int E = d.E, H = d.H;
    // the action: (vx, vy) holonomic, (v, r) otherwise; (ax, ay) = the velocity the step moves and collision-checks with
    const double a0 = act_aos ? act[2 * (size_t)e] : act[e];
    const double a1 = act_aos ? act[2 * (size_t)e + 1] : act[E + e];
    const double th = p.kinematics != CN_KIN_HOLONOMIC ? theta[e] : 0.0;
    double ax, ay;
    cn_effective_velocity(p.kinematics, th, a0, a1, ax, ay);
    const double dt = p.time_step;
    const double t = time[e];
    auto ag = [&](int f, int a) { return st[st_idx(d, f, a, e)]; };
    const StepOutcome oc = cn_step_outcome(p, ag, H, t, ax, ay);
    const double reward = oc.reward, dmin = oc.dmin;
    const int done = oc.done, info = oc.info;
    const double endx = ag(F_PX, 0) + ax * dt, endy = ag(F_PY, 0) + ay * dt;
    reward_o[e] = reward; done_o[e] = (uint8_t)done; info_o[e] = (uint8_t)info; dmin_o[e] = dmin;

    if (update) {
        // crowd_sim.py:414-417, agent.py:122-135
        st[st_idx(d, F_PX, 0, e)] = endx;
        st[st_idx(d, F_PY, 0, e)] = endy;
        if (p.kinematics == CN_KIN_HOLONOMIC) {
            st[st_idx(d, F_VX, 0, e)] = ax;
            st[st_idx(d, F_VY, 0, e)] = ay;
        } else {
            // agent.py:131-133: the heading wraps to [0, 2 pi) (Python float %), the stored velocity uses the wrapped heading
            const double TWO_PI = 2 * 3.141592653589793;
            double t2 = fmod(th + a1, TWO_PI);
            if (t2 < 0) t2 += TWO_PI;
            theta[e] = t2;
            st[st_idx(d, F_VX, 0, e)] = a0 * cos(t2);
            st[st_idx(d, F_VY, 0, e)] = a0 * sin(t2);
        }
        // one human ahead: the loads of human h + 1 are issued before the stores of human h (the compiler cannot move a load
        // of st above a store to st, so the plain loop was one DRAM round trip per human: 100 us at 50 humans)
        double hvx = human_v[(size_t)(0 * H) * E + e], hvy = human_v[(size_t)(1 * H) * E + e];
        double opx = st[st_idx(d, F_PX, 1, e)], opy = st[st_idx(d, F_PY, 1, e)];
        for (int h = 1; h <= H; ++h) {
            double nvx = 0, nvy = 0, npx = 0, npy = 0;
            if (h < H) {
                nvx = human_v[(size_t)(0 * H + h) * E + e]; nvy = human_v[(size_t)(1 * H + h) * E + e];
                npx = st[st_idx(d, F_PX, h + 1, e)]; npy = st[st_idx(d, F_PY, h + 1, e)];
            }
            st[st_idx(d, F_PX, h, e)] = opx + hvx * dt;
            st[st_idx(d, F_PY, h, e)] = opy + hvy * dt;
            st[st_idx(d, F_VX, h, e)] = hvx;
            st[st_idx(d, F_VY, h, e)] = hvy;
            hvx = nvx; hvy = nvy; opx = npx; opy = npy;
        }
        const double tn = t + dt;
        time[e] = tn;
        // explorer.py counters
        const int k = acc.ep_steps[e];
        const double disc = pow(p.gamma, (double)k * dt * st[st_idx(d, F_VPREF, 0, e)]);
        const double ret = acc.ep_return[e] + disc * reward;
        acc.steps[e] += 1;
        if (info == CN_DANGER) { acc.too_close[e] += 1; acc.sum_min_dist[e] += dmin; }
        if (done) {
            acc.episodes[e] += 1;
            if (info == CN_REACHGOAL) { acc.success[e] += 1; acc.sum_success_time[e] += tn; }
            else if (info == CN_COLLISION) { acc.collision[e] += 1; acc.sum_collision_time[e] += tn; }
            else { acc.timeout[e] += 1; acc.sum_timeout_time[e] += p.time_limit; }
            acc.sum_return[e] += ret;
            acc.ep_return[e] = 0.0; acc.ep_steps[e] = 0;
            if (!p.auto_reset) frozen[e] = 1;
            // rollout step of an auto-reset env: the finished episode is re-generated here instead of by a reset_kernel launch
            // behind this one (a 13 us latency-bound kernel per step for the ~2 % of envs that finish)
            else if (fuse_reset) reset_env(p, e, st, time, frozen, acc, theta);
        } else {
            acc.ep_return[e] = ret; acc.ep_steps[e] = k + 1;
        }
    } else {
        // onestep_lookahead observation (crowd_sim.py:428-430, agent.py:63-74)
        for (int h = 1; h <= H; ++h) {
            const double hvx = human_v[(size_t)(0 * H + h - 1) * E + e];
            const double hvy = human_v[(size_t)(1 * H + h - 1) * E + e];
            next_obs[(size_t)(0 * H + h - 1) * E + e] = st[st_idx(d, F_PX, h, e)] + hvx * dt;
            next_obs[(size_t)(1 * H + h - 1) * E + e] = st[st_idx(d, F_PY, h, e)] + hvy * dt;
            next_obs[(size_t)(2 * H + h - 1) * E + e] = hvx;
            next_obs[(size_t)(3 * H + h - 1) * E + e] = hvy;
            next_obs[(size_t)(4 * H + h - 1) * E + e] = st[st_idx(d, F_R, h, e)];
        }
    }
}

// CrowdSim.reset of ONE env by a whole WARP, for crowds where the rejection sampling dominates (H = 50 in a 10 m square: every
// human is re-drawn several times and checked against up to 50 earlier agents -- the thread-per-env form took 1.3 ms of a
// 6.7 ms step, profiles/r02a_bench.json).  Every lane runs the SAME Philox stream, so the draws and the accept / reject
// decisions are those of reset_env; only the distance checks against the earlier agents are split over the lanes
// (agent a = lane, lane + 32, ...) and joined by a warp vote.  Bit-identical scenes (tests: shard invariance, properties).
__device__ void reset_env_warp(const EnvParams &p, int e, int lane, double *__restrict__ st, double *__restrict__ time,
                               uint8_t *__restrict__ frozen, const EnvAccum &acc, double *__restrict__ theta)
{
    const EnvDims d = p.d;
    PhiloxStream rng;
    rng.init(p.seed, (uint64_t)(p.env_id_offset + e), acc.episode_ctr[e], 0u);
    __syncwarp();
    if (lane == 0) acc.episode_ctr[e] += 1;
    const double PI = 3.141592653589793;
    if (lane == 0) {
        st[st_idx(d, F_PX, 0, e)] = 0.0; st[st_idx(d, F_PY, 0, e)] = -p.circle_radius;
        st[st_idx(d, F_VX, 0, e)] = 0.0; st[st_idx(d, F_VY, 0, e)] = 0.0;
        st[st_idx(d, F_GX, 0, e)] = 0.0; st[st_idx(d, F_GY, 0, e)] = p.circle_radius;
        st[st_idx(d, F_R, 0, e)] = p.robot_radius; st[st_idx(d, F_VPREF, 0, e)] = p.robot_v_pref;
    }
    __syncwarp();
    // The agents placed so far, cached in REGISTERS of the lane that checks them (agent a = lane + 32 k): a retry of the
    // rejection sampling then costs no memory round trip (they were re-read from global memory on every try: ~200 us at
    // 50 humans).  Same doubles, same arithmetic, same decisions.
    constexpr int kSlots = (CN_MAX_HUMANS + 32) / 32;
    double cpx[kSlots], cpy[kSlots], cgx[kSlots], cgy[kSlots], cr[kSlots];
#pragma unroll
    for (int k = 0; k < kSlots; ++k) { cpx[k] = cpy[k] = cgx[k] = cgy[k] = cr[k] = 0.0; }
    if (lane == 0) { cpx[0] = 0.0; cpy[0] = -p.circle_radius; cgx[0] = 0.0; cgy[0] = p.circle_radius; cr[0] = p.robot_radius; }
    const int MAX_TRIES = 4096;
    for (int i = 1; i <= d.H; ++i) {
        double px = 0, py = 0, gx = 0, gy = 0;
        double h_radius = p.human_radius, h_v_pref = p.human_v_pref;
        if (p.randomize_attributes) {
            h_v_pref = 0.5 + (1.5 - 0.5) * rng.next();
            h_radius = 0.3 + (0.5 - 0.3) * rng.next();
        }
        if (p.sim_rule == CN_CIRCLE_CROSSING) {
            for (int tries = 0; tries < MAX_TRIES; ++tries) {
                const double angle = rng.next() * PI * 2;
                const double px_noise = (rng.next() - 0.5) * h_v_pref;
                const double py_noise = (rng.next() - 0.5) * h_v_pref;
                px = p.circle_radius * cos(angle) + px_noise;
                py = p.circle_radius * sin(angle) + py_noise;
                bool collide = false;
#pragma unroll
                for (int k = 0; k < kSlots; ++k) {
                    if (lane + 32 * k >= i) continue;
                    const double min_dist = h_radius + cr[k] + p.discomfort_dist;
                    collide = collide || norm2d(px - cpx[k], py - cpy[k]) < min_dist || norm2d(px - cgx[k], py - cgy[k]) < min_dist;
                }
                if (!__any_sync(0xffffffffu, collide)) break;
            }
            gx = -px; gy = -py;
        } else {
            const double sign = (rng.next() > 0.5) ? -1.0 : 1.0;
            for (int tries = 0; tries < MAX_TRIES; ++tries) {
                px = rng.next() * p.square_width * 0.5 * sign;
                py = (rng.next() - 0.5) * p.square_width;
                bool collide = false;
#pragma unroll
                for (int k = 0; k < kSlots; ++k)
                    if (lane + 32 * k < i)
                        collide = collide || norm2d(px - cpx[k], py - cpy[k]) < h_radius + cr[k] + p.discomfort_dist;
                if (!__any_sync(0xffffffffu, collide)) break;
            }
            for (int tries = 0; tries < MAX_TRIES; ++tries) {
                gx = rng.next() * p.square_width * 0.5 * -sign;
                gy = (rng.next() - 0.5) * p.square_width;
                bool collide = false;
#pragma unroll
                for (int k = 0; k < kSlots; ++k)
                    if (lane + 32 * k < i)
                        collide = collide || norm2d(gx - cgx[k], gy - cgy[k]) < h_radius + cr[k] + p.discomfort_dist;
                if (!__any_sync(0xffffffffu, collide)) break;
            }
        }
        if (lane == 0) {
            st[st_idx(d, F_PX, i, e)] = px; st[st_idx(d, F_PY, i, e)] = py;
            st[st_idx(d, F_VX, i, e)] = 0.0; st[st_idx(d, F_VY, i, e)] = 0.0;
            st[st_idx(d, F_GX, i, e)] = gx; st[st_idx(d, F_GY, i, e)] = gy;
            st[st_idx(d, F_R, i, e)] = h_radius; st[st_idx(d, F_VPREF, i, e)] = h_v_pref;
        }
#pragma unroll
        for (int k = 0; k < kSlots; ++k)
            if (lane + 32 * k == i) { cpx[k] = px; cpy[k] = py; cgx[k] = gx; cgy[k] = gy; cr[k] = h_radius; }
    }
    if (lane == 0) {
        time[e] = 0.0;
        theta[e] = 1.5707963267948966;
        frozen[e] = 0;
        acc.ep_steps[e] = 0; acc.ep_return[e] = 0.0;
    }
}

// One WARP per env: the form of reset_kernel for large crowds (see reset_env_warp).
__global__ void __launch_bounds__(128) reset_warp_kernel(EnvParams p, double *__restrict__ st, double *__restrict__ time,
                                                         const uint8_t *__restrict__ done, int only_done,
                                                         uint8_t *__restrict__ frozen, EnvAccum acc, double *__restrict__ theta)
{
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= p.d.E) return;                      // warp-uniform
    if (only_done && !done[e]) return;
    reset_env_warp(p, e, lane, st, time, frozen, acc, theta);
}

// One thread per env (explicit resets, and the auto-reset of finished episodes when it is not fused into step_kernel).
__global__ void __launch_bounds__(128) reset_kernel(EnvParams p, double *__restrict__ st, double *__restrict__ time,
                                                    const uint8_t *__restrict__ done, int only_done,
                                                    uint8_t *__restrict__ frozen, EnvAccum acc, double *__restrict__ theta)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.d.E) return;
    if (only_done && !done[e]) return;
    reset_env(p, e, st, time, frozen, acc, theta);
}

// stage (AoS, E x A1 x 8) <-> state (SoA, 8 x A1 x E)
__global__ void pack_kernel(EnvDims d, double *__restrict__ st, double *__restrict__ stage, int to_soa)
{
    const size_t n = (size_t)d.E * d.A1 * F_COUNT;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        // i indexes the SoA side so that its accesses are coalesced
        const int e = (int)(i % d.E);
        const size_t r = i / d.E;
        const int a = (int)(r % d.A1), f = (int)(r / d.A1);
        const size_t j = ((size_t)e * d.A1 + a) * F_COUNT + f;
        if (to_soa) st[i] = stage[j];
        else stage[j] = st[i];
    }
}

// Packed host exchange block (cn_rollout_step_host_packed):  [agents E x A1 x 8 f64 | times E f64 | reward E f64 |
// action_idx E i32 | done E u8 | info E u8]; the input block is its first two segments.
__global__ void io_unpack_kernel(EnvDims d, double *__restrict__ st, double *__restrict__ time, const double *__restrict__ blk)
{
    const size_t n = (size_t)d.E * d.A1 * F_COUNT;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n + d.E; i += (size_t)gridDim.x * blockDim.x) {
        if (i >= n) { time[i - n] = blk[i]; continue; }
        const int e = (int)(i % d.E);
        const size_t r = i / d.E;
        const int a = (int)(r % d.A1), f = (int)(r / d.A1);
        st[i] = blk[((size_t)e * d.A1 + a) * F_COUNT + f];
    }
}

__global__ void io_pack_kernel(EnvDims d, const double *__restrict__ st, const double *__restrict__ time,
                               const double *__restrict__ reward, const int32_t *__restrict__ action_idx,
                               const uint8_t *__restrict__ done, const uint8_t *__restrict__ info, double *__restrict__ blk)
{
    const size_t n = (size_t)d.E * d.A1 * F_COUNT;
    int32_t *b_idx = reinterpret_cast<int32_t *>(blk + n + 2 * (size_t)d.E);
    uint8_t *b_done = reinterpret_cast<uint8_t *>(b_idx + d.E), *b_info = b_done + d.E;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n + d.E; i += (size_t)gridDim.x * blockDim.x) {
        if (i >= n) {
            const size_t e = i - n;
            blk[n + e] = time[e];
            blk[n + d.E + e] = reward[e];
            b_idx[e] = action_idx[e]; b_done[e] = done[e]; b_info[e] = info[e];
            continue;
        }
        const int e = (int)(i % d.E);
        const size_t r = i / d.E;
        const int a = (int)(r % d.A1), f = (int)(r / d.A1);
        blk[((size_t)e * d.A1 + a) * F_COUNT + f] = st[i];
    }
}

__global__ void clear_episode_kernel(int E, uint8_t *frozen, EnvAccum acc, uint8_t *done, double *theta)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    frozen[e] = 0; done[e] = 0;
    theta[e] = 1.5707963267948966;                       // a host-generated scene is a reset: robot heading pi / 2
    acc.ep_steps[e] = 0; acc.ep_return[e] = 0.0;
}

// human_v (2 x H x E) -> E x H x 2 ; next_obs (5 x H x E) -> E x H x 5
__global__ void transpose_out_kernel(int E, int H, int C, const double *__restrict__ src, double *__restrict__ dst)
{
    const size_t n = (size_t)E * H * C;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(i % E);
        const size_t r = i / E;
        const int h = (int)(r % H), c = (int)(r / H);
        dst[((size_t)e * H + h) * C + c] = src[i];
    }
}

__global__ void set_actions_kernel(int E, const double *__restrict__ aos, double *__restrict__ action_xy,
                                   int32_t *__restrict__ action_idx)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    action_xy[e] = aos[2 * (size_t)e];
    action_xy[E + e] = aos[2 * (size_t)e + 1];
    action_idx[e] = -1;
}

// E x H x 2 (env-major, what a host world model produces) -> human_v[2][H][E] (the layout ORCA writes)
__global__ void set_human_v_kernel(int E, int H, const double *__restrict__ aos, double *__restrict__ human_v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * H) return;
    const int h = i / E, e = i - h * E;
    human_v[(size_t)(0 * H + h) * E + e] = aos[((size_t)e * H + h) * 2];
    human_v[(size_t)(1 * H + h) * E + e] = aos[((size_t)e * H + h) * 2 + 1];
}

}  // namespace

static inline int grid_for(int n, int block) { return (n + block - 1) / block; }

int cn_launch_set_human_v(cn_env *env, const double *aos_dev, cudaStream_t s)
{
    const int n = env->p.d.E * env->p.d.H;
    set_human_v_kernel<<<grid_for(n, 128), 128, 0, s>>>(env->p.d.E, env->p.d.H, aos_dev, env->human_v);
    CN_LAUNCH_CHECK();
    env->orca_valid = 1;          // the step (and a query_env lookahead) use these velocities instead of an ORCA solve
    return CN_OK;
}

int cn_launch_orca(cn_env *env, cudaStream_t s)
{
    const int n = env->p.d.E * env->p.d.H;
    const int bs = cn_small_block();
    orca_humans_kernel<<<grid_for(n, bs), bs, 0, s>>>(env->p, env->state, env->frozen, env->human_v);
    CN_LAUNCH_CHECK();
    env->orca_valid = 1;
    return CN_OK;
}

int cn_launch_robot_orca(cn_env *env, double safety_space, cudaStream_t s)
{
    orca_robot_kernel<<<grid_for(env->p.d.E, 128), 128, 0, s>>>(env->p, env->state, env->frozen, safety_space,
                                                                  env->action_xy, env->action_idx);
    CN_LAUNCH_CHECK();
    return CN_OK;
}

int cn_launch_step(cn_env *env, const double *action_xy_dev, int update, cudaStream_t s, int fuse_reset)
{
    const double *act = action_xy_dev ? action_xy_dev : env->action_xy;
    const int bs = cn_small_block();
    // large crowds: the re-generation of finished episodes is a warp-per-env kernel of its own instead of one thread of the
    // step kernel doing the whole rejection sampling while its warp waits
    const bool split_reset = fuse_reset && update && env->p.d.H >= kWarpResetHumans;
    if (split_reset) fuse_reset = 0;
    step_kernel<<<grid_for(env->p.d.E, bs), bs, 0, s>>>(env->p, env->state, env->time, env->human_v, act,
                                                            action_xy_dev ? 1 : 0, update, env->reward, env->done,
                                                            env->info, env->dmin, env->next_obs, env->frozen,
                                                            env->acc, env->theta, fuse_reset);
    CN_LAUNCH_CHECK();
    if (update) env->orca_valid = 0;
    if (split_reset) return cn_launch_reset(env, 1, s);
    return CN_OK;
}

int cn_launch_reset(cn_env *env, int only_done, cudaStream_t s)
{
    const int bs = cn_small_block();
    if (env->p.d.H >= kWarpResetHumans)
        reset_warp_kernel<<<grid_for(env->p.d.E * 32, 128), 128, 0, s>>>(env->p, env->state, env->time, env->done, only_done,
                                                                          env->frozen, env->acc, env->theta);
    else
        reset_kernel<<<grid_for(env->p.d.E, bs), bs, 0, s>>>(env->p, env->state, env->time, env->done, only_done,
                                                             env->frozen, env->acc, env->theta);
    CN_LAUNCH_CHECK();
    env->orca_valid = 0;
    return CN_OK;
}

int cn_launch_pack(cn_env *env, int to_soa, cudaStream_t s)
{
    const size_t n = (size_t)env->p.d.E * env->p.d.A1 * F_COUNT;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    pack_kernel<<<grid, 256, 0, s>>>(env->p.d, env->state, env->stage, to_soa);
    CN_LAUNCH_CHECK();
    if (to_soa) {
        clear_episode_kernel<<<grid_for(env->p.d.E, 128), 128, 0, s>>>(env->p.d.E, env->frozen, env->acc, env->done, env->theta);
        CN_LAUNCH_CHECK();
        env->orca_valid = 0;
    }
    return CN_OK;
}

// host-stepped mode: refresh the SoA from the staging buffer but keep episode accumulators / frozen flags
int cn_launch_pack_keep(cn_env *env, cudaStream_t s)
{
    const size_t n = (size_t)env->p.d.E * env->p.d.A1 * F_COUNT;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    pack_kernel<<<grid, 256, 0, s>>>(env->p.d, env->state, env->stage, 1);
    CN_LAUNCH_CHECK();
    env->orca_valid = 0;
    return CN_OK;
}

int cn_launch_io(cn_env *env, double *blk, int unpack, cudaStream_t s)
{
    const size_t n = (size_t)env->p.d.E * env->p.d.A1 * F_COUNT + env->p.d.E;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    // <= 32 registers / thread: these blocks fit beside a resident tc_rows_pair CTA (see cn_small_block)
    grid *= 2;
    if (unpack) {
        io_unpack_kernel<<<grid, 128, 0, s>>>(env->p.d, env->state, env->time, blk);
        env->orca_valid = 0;
    } else {
        io_pack_kernel<<<grid, 128, 0, s>>>(env->p.d, env->state, env->time, env->reward, env->action_idx, env->done, env->info, blk);
    }
    CN_LAUNCH_CHECK();
    return CN_OK;
}

int cn_launch_transpose_out(int E, int H, int C, const double *src, double *dst, cudaStream_t s)
{
    const size_t n = (size_t)E * H * C;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    transpose_out_kernel<<<grid, 256, 0, s>>>(E, H, C, src, dst);
    CN_LAUNCH_CHECK();
    return CN_OK;
}

int cn_launch_set_actions(cn_env *env, const double *aos_dev, cudaStream_t s)
{
    set_actions_kernel<<<grid_for(env->p.d.E, 128), 128, 0, s>>>(env->p.d.E, aos_dev, env->action_xy,
                                                                   env->action_idx);
    CN_LAUNCH_CHECK();
    return CN_OK;
}
