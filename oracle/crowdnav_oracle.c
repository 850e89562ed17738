/*
 * crowdnav_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See crowdnav_oracle.h for scope and parity status.  Compile with
 *   gcc -O2 -fPIC -shared -ffp-contract=off -fopenmp   (see oracle/Makefile)
 * -ffp-contract=off matters: x86 builds of rvo2 and CPython do not contract
 * mul+add into FMA; the only fused operation on the reference path is numpy's
 * 2-element dot inside np.linalg.norm, restated explicitly in norm2() below.
 *
 * All file:line citations are relative to /root/reference.
 */
#include "crowdnav_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ===================================================================== */
/* rvo2 restatement (third-party RVO2 Library 2.0.x, float32)            */
/* Published algorithm: Agent.cpp (computeNeighbors, insertAgentNeighbor, */
/* computeNewVelocity, linearProgram1/2/3), Vector2.h, Definitions.h.     */
/* Reference call sites: crowd_sim/envs/policy/orca.py:94-129.            */
/* ===================================================================== */

#define RVO_EPSILON 0.00001f

typedef struct { float x, y; } v2;
typedef struct { v2 point, direction; } rvo_line;

static inline v2 V(float x, float y) { v2 r = {x, y}; return r; }
static inline v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 vneg(v2 a) { return V(-a.x, -a.y); }
static inline v2 vscale(float s, v2 a) { return V(s * a.x, s * a.y); }
static inline float vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline float vdet(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
static inline float vabssq(v2 a) { return vdot(a, a); }
static inline float vabs(v2 a) { return sqrtf(vdot(a, a)); }
/* Vector2::operator/(float): multiplies by the reciprocal */
static inline v2 vdiv(v2 a, float s) { const float inv = 1.0f / s; return V(a.x * inv, a.y * inv); }
static inline v2 vnormalize(v2 a) { return vdiv(a, vabs(a)); }
static inline float sqrf(float a) { return a * a; }

/* linearProgram1 (RVO2 Agent.cpp) */
static int rvo_lp1(const rvo_line *lines, int lineNo, float radius, v2 opt, int directionOpt, v2 *result)
{
    const float dotProduct = vdot(lines[lineNo].point, lines[lineNo].direction);
    const float discriminant = sqrf(dotProduct) + sqrf(radius) - vabssq(lines[lineNo].point);
    if (discriminant < 0.0f) return 0;

    const float sqrtDiscriminant = sqrtf(discriminant);
    float tLeft = -dotProduct - sqrtDiscriminant;
    float tRight = -dotProduct + sqrtDiscriminant;

    for (int i = 0; i < lineNo; ++i) {
        const float denominator = vdet(lines[lineNo].direction, lines[i].direction);
        const float numerator = vdet(lines[i].direction, vsub(lines[lineNo].point, lines[i].point));
        if (fabsf(denominator) <= RVO_EPSILON) {
            if (numerator < 0.0f) return 0;
            continue;
        }
        const float t = numerator / denominator;
        if (denominator >= 0.0f) tRight = fminf(tRight, t); /* std::min */
        else tLeft = fmaxf(tLeft, t);                       /* std::max */
        if (tLeft > tRight) return 0;
    }

    if (directionOpt) {
        if (vdot(opt, lines[lineNo].direction) > 0.0f)
            *result = vadd(lines[lineNo].point, vscale(tRight, lines[lineNo].direction));
        else
            *result = vadd(lines[lineNo].point, vscale(tLeft, lines[lineNo].direction));
    } else {
        const float t = vdot(lines[lineNo].direction, vsub(opt, lines[lineNo].point));
        if (t < tLeft) *result = vadd(lines[lineNo].point, vscale(tLeft, lines[lineNo].direction));
        else if (t > tRight) *result = vadd(lines[lineNo].point, vscale(tRight, lines[lineNo].direction));
        else *result = vadd(lines[lineNo].point, vscale(t, lines[lineNo].direction));
    }
    return 1;
}

/* linearProgram2 (RVO2 Agent.cpp) */
static int rvo_lp2(const rvo_line *lines, int n, float radius, v2 opt, int directionOpt, v2 *result)
{
    if (directionOpt) *result = vscale(radius, opt);
    else if (vabssq(opt) > sqrf(radius)) *result = vscale(radius, vnormalize(opt));
    else *result = opt;

    for (int i = 0; i < n; ++i) {
        if (vdet(lines[i].direction, vsub(lines[i].point, *result)) > 0.0f) {
            const v2 tmp = *result;
            if (!rvo_lp1(lines, i, radius, opt, directionOpt, result)) {
                *result = tmp;
                return i;
            }
        }
    }
    return n;
}

/* linearProgram3 (RVO2 Agent.cpp); numObstLines == 0 on this path */
static void rvo_lp3(const rvo_line *lines, int n, int beginLine, float radius, v2 *result)
{
    float distance = 0.0f;
    rvo_line proj[ORC_MAX_NEIGHBORS];
    for (int i = beginLine; i < n; ++i) {
        if (vdet(lines[i].direction, vsub(lines[i].point, *result)) > distance) {
            int np = 0;
            for (int j = 0; j < i; ++j) {
                rvo_line line;
                const float determinant = vdet(lines[i].direction, lines[j].direction);
                if (fabsf(determinant) <= RVO_EPSILON) {
                    if (vdot(lines[i].direction, lines[j].direction) > 0.0f) continue;
                    line.point = vscale(0.5f, vadd(lines[i].point, lines[j].point));
                } else {
                    const float t = vdet(lines[j].direction, vsub(lines[i].point, lines[j].point)) / determinant;
                    line.point = vadd(lines[i].point, vscale(t, lines[i].direction));
                }
                line.direction = vnormalize(vsub(lines[j].direction, lines[i].direction));
                proj[np++] = line;
            }
            const v2 tmp = *result;
            if (rvo_lp2(proj, np, radius, V(-lines[i].direction.y, lines[i].direction.x), 1, result) < np)
                *result = tmp;
            distance = vdet(lines[i].direction, vsub(lines[i].point, *result));
        }
    }
}

void orc_rvo_new_velocity(int n, const float *px, const float *py, const float *vx, const float *vy,
                          const float *radius, int self, float max_speed, float pref_x, float pref_y,
                          float neighbor_dist, int max_neighbors, float time_horizon, float time_step,
                          float *out_vx, float *out_vy, int *out_info)
{
    /* Agent::computeNeighbors + insertAgentNeighbor */
    float nb_d2[ORC_MAX_NEIGHBORS];
    int nb_id[ORC_MAX_NEIGHBORS];
    int nn = 0;
    if (max_neighbors > ORC_MAX_NEIGHBORS) max_neighbors = ORC_MAX_NEIGHBORS;
    const v2 pos = V(px[self], py[self]);
    const v2 vel = V(vx[self], vy[self]);
    if (max_neighbors > 0) {
        float rangeSq = sqrf(neighbor_dist);
        for (int o = 0; o < n; ++o) {
            if (o == self) continue;
            const float distSq = vabssq(vsub(pos, V(px[o], py[o])));
            if (distSq < rangeSq) {
                if (nn < max_neighbors) { nb_d2[nn] = distSq; nb_id[nn] = o; ++nn; }
                int i = nn - 1;
                while (i != 0 && distSq < nb_d2[i - 1]) {
                    nb_d2[i] = nb_d2[i - 1]; nb_id[i] = nb_id[i - 1];
                    --i;
                }
                nb_d2[i] = distSq; nb_id[i] = o;
                if (nn == max_neighbors) rangeSq = nb_d2[nn - 1];
            }
        }
    }

    /* Agent::computeNewVelocity */
    rvo_line lines[ORC_MAX_NEIGHBORS];
    const float invTimeHorizon = 1.0f / time_horizon;
    for (int k = 0; k < nn; ++k) {
        const int o = nb_id[k];
        const v2 relativePosition = vsub(V(px[o], py[o]), pos);
        const v2 relativeVelocity = vsub(vel, V(vx[o], vy[o]));
        const float distSq = vabssq(relativePosition);
        const float combinedRadius = radius[self] + radius[o];
        const float combinedRadiusSq = sqrf(combinedRadius);
        rvo_line line;
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
            const float invTimeStep = 1.0f / time_step;
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
    const int lineFail = rvo_lp2(lines, nn, max_speed, V(pref_x, pref_y), 0, &newV);
    if (lineFail < nn) rvo_lp3(lines, nn, lineFail, max_speed, &newV);
    *out_vx = newV.x;
    *out_vy = newV.y;
    if (out_info) { out_info[0] = nn; out_info[1] = lineFail; }
}

void orc_rvo_do_step(int n, float *px, float *py, float *vx, float *vy, const float *radius,
                     const float *max_speed, const float *pref_x, const float *pref_y,
                     const float *neighbor_dist, const int *max_neighbors, const float *time_horizon,
                     float time_step)
{
    float *nvx = (float *)malloc(sizeof(float) * (size_t)n * 2);
    float *nvy = nvx + n;
    for (int i = 0; i < n; ++i)
        orc_rvo_new_velocity(n, px, py, vx, vy, radius, i, max_speed[i], pref_x[i], pref_y[i],
                             neighbor_dist[i], max_neighbors[i], time_horizon[i], time_step,
                             &nvx[i], &nvy[i], NULL);
    for (int i = 0; i < n; ++i) { /* Agent::update */
        vx[i] = nvx[i]; vy[i] = nvy[i];
        px[i] += vx[i] * time_step;
        py[i] += vy[i] * time_step;
    }
    free(nvx);
}

/* ===================================================================== */
/* CrowdSim (float64)                                                     */
/* ===================================================================== */

/* np.linalg.norm of a 2-vector = sqrt(dot(x,x)); numpy's dot evaluates the
 * 2-element product as fma(b, b, a*a) (verified against numpy in this
 * container on 2e5 random pairs, scripts/gen_golden.py --check-norm). */
static inline double norm2(double a, double b) { return sqrt(fma(b, b, a * a)); }

/* crowd_sim/envs/utils/utils.py:4-26 */
double orc_point_to_segment_dist(double x1, double y1, double x2, double y2, double x3, double y3)
{
    const double px = x2 - x1, py = y2 - y1;
    if (px == 0 && py == 0) return norm2(x3 - x1, y3 - y1);
    double u = ((x3 - x1) * px + (y3 - y1) * py) / (px * px + py * py);
    if (u > 1) u = 1;
    else if (u < 0) u = 0;
    const double x = x1 + u * px, y = y1 + u * py;
    return norm2(x - x3, y - y3);
}

#define AG(a, f) agents[(a) * ORC_AGENT_STRIDE + (f)]
enum { F_PX = 0, F_PY, F_VX, F_VY, F_GX, F_GY, F_R, F_VPREF };

/* ORCA.predict for `self` with the listed others (orca.py:82-132).
 * doubles are rounded to float32 at the rvo2 FFI (Cython float arguments). */
static void orca_predict(const orc_env_cfg *cfg, const double *agents, int self, const int *others,
                         int n_others, double safety_space, double *out_vxy)
{
    float px[ORC_MAX_NEIGHBORS * 8], py[ORC_MAX_NEIGHBORS * 8], vx[ORC_MAX_NEIGHBORS * 8],
        vy[ORC_MAX_NEIGHBORS * 8], rad[ORC_MAX_NEIGHBORS * 8];
    const int n = n_others + 1;
    px[0] = (float)AG(self, F_PX); py[0] = (float)AG(self, F_PY);
    vx[0] = (float)AG(self, F_VX); vy[0] = (float)AG(self, F_VY);
    rad[0] = (float)(AG(self, F_R) + 0.01 + safety_space);        /* orca.py:100 */
    for (int k = 0; k < n_others; ++k) {
        const int o = others[k];
        px[k + 1] = (float)AG(o, F_PX); py[k + 1] = (float)AG(o, F_PY);
        vx[k + 1] = (float)AG(o, F_VX); vy[k + 1] = (float)AG(o, F_VY);
        rad[k + 1] = (float)(AG(o, F_R) + 0.01 + safety_space);   /* orca.py:103 */
    }
    /* pref_vel = goal - pos, un-normalised (orca.py:113,123) */
    const float prefx = (float)(AG(self, F_GX) - AG(self, F_PX));
    const float prefy = (float)(AG(self, F_GY) - AG(self, F_PY));
    float ovx, ovy;
    orc_rvo_new_velocity(n, px, py, vx, vy, rad, 0, (float)AG(self, F_VPREF), prefx, prefy,
                         (float)cfg->neighbor_dist, cfg->max_neighbors, (float)cfg->time_horizon,
                         (float)cfg->time_step, &ovx, &ovy, NULL);
    out_vxy[0] = (double)ovx;  /* getAgentVelocity -> Python float */
    out_vxy[1] = (double)ovy;
}

/* crowd_sim.py:337-342 */
void orc_human_actions(const orc_env_cfg *cfg, int H, const double *agents, double *out_vxy)
{
    int others[ORC_MAX_NEIGHBORS * 8];
    for (int h = 1; h <= H; ++h) {
        int n = 0;
        for (int o = 1; o <= H; ++o) if (o != h) others[n++] = o;
        if (cfg->robot_visible) others[n++] = 0;                  /* crowd_sim.py:340-341 */
        orca_predict(cfg, agents, h, others, n, cfg->human_safety_space, &out_vxy[2 * (h - 1)]);
    }
}

/* robot.act(ob) with an ORCA policy (train.py:157-166, robot.py:9-14) */
void orc_robot_orca_action(const orc_env_cfg *cfg, int H, const double *agents, double safety_space,
                           double *out_vxy)
{
    int others[ORC_MAX_NEIGHBORS * 8];
    for (int o = 1; o <= H; ++o) others[o - 1] = o;
    orca_predict(cfg, agents, 0, others, H, safety_space, out_vxy);
}

/* crowd_sim.py:344-403 */
/* Robot action -> the velocity the reference's non-holonomic branches use for collision checking and for
 * compute_position: v * cos(theta + r), v * sin(theta + r) (crowd_sim.py:353-354, agent.py:115-117).  Holonomic
 * actions are already velocities. */
static void effective_velocity(int kinematics, double theta, double a0, double a1, double *ax, double *ay)
{
    if (kinematics == ORC_KIN_HOLONOMIC) { *ax = a0; *ay = a1; }
    else { *ax = a0 * cos(a1 + theta); *ay = a0 * sin(a1 + theta); }
}

void orc_step_outcome_k(const orc_env_cfg *cfg, int H, const double *agents, double global_time, int kinematics,
                        double theta, double a0, double a1, double *reward, int *done, int *info, double *dmin_out)
{
    double ax, ay;
    effective_velocity(kinematics, theta, a0, a1, &ax, &ay);
    orc_step_outcome(cfg, H, agents, global_time, ax, ay, reward, done, info, dmin_out);
}

/* agent.py:122-135 for the robot + crowd_sim.py:414-417 for the humans.  Non-holonomic: the position moves with the
 * velocity of the UNWRAPPED heading theta + r, the stored velocity uses the heading wrapped to [0, 2 pi). */
void orc_apply_step_k(const orc_env_cfg *cfg, int H, double *agents, double *global_time, int kinematics,
                      double *theta, double a0, double a1, const double *human_vxy)
{
    double ax, ay;
    effective_velocity(kinematics, *theta, a0, a1, &ax, &ay);
    orc_apply_step(cfg, H, agents, global_time, ax, ay, human_vxy);
    if (kinematics != ORC_KIN_HOLONOMIC) {
        double t = fmod(*theta + a1, 2 * 3.141592653589793);      /* Python float %: result takes the divisor's sign */
        if (t < 0) t += 2 * 3.141592653589793;
        *theta = t;
        AG(0, F_VX) = a0 * cos(t);
        AG(0, F_VY) = a0 * sin(t);
    }
}

void orc_step_outcome(const orc_env_cfg *cfg, int H, const double *agents, double global_time,
                      double ax, double ay, double *reward, int *done, int *info, double *dmin_out)
{
    double dmin = INFINITY;
    int collision = 0;
    for (int h = 1; h <= H; ++h) {
        const double px = AG(h, F_PX) - AG(0, F_PX);
        const double py = AG(h, F_PY) - AG(0, F_PY);
        const double vx = AG(h, F_VX) - ax;
        const double vy = AG(h, F_VY) - ay;
        const double ex = px + vx * cfg->time_step;
        const double ey = py + vy * cfg->time_step;
        const double closest = orc_point_to_segment_dist(px, py, ex, ey, 0, 0) - AG(h, F_R) - AG(0, F_R);
        if (closest < 0) { collision = 1; break; }
        else if (closest < dmin) dmin = closest;
    }
    /* crowd_sim.py:379-380, agent.py:110-120 */
    const double endx = AG(0, F_PX) + ax * cfg->time_step;
    const double endy = AG(0, F_PY) + ay * cfg->time_step;
    const int reaching_goal = norm2(endx - AG(0, F_GX), endy - AG(0, F_GY)) < AG(0, F_R);

    if (global_time >= cfg->time_limit - 1) { *reward = 0; *done = 1; *info = ORC_TIMEOUT; }
    else if (collision) { *reward = cfg->collision_penalty; *done = 1; *info = ORC_COLLISION; }
    else if (reaching_goal) { *reward = cfg->success_reward; *done = 1; *info = ORC_REACHGOAL; }
    else if (dmin < cfg->discomfort_dist) {
        *reward = (dmin - cfg->discomfort_dist) * cfg->discomfort_penalty_factor * cfg->time_step;
        *done = 0; *info = ORC_DANGER;
    } else { *reward = 0; *done = 0; *info = ORC_NOTHING; }
    if (dmin_out) *dmin_out = dmin;
}

/* crowd_sim.py:414-417 */
void orc_apply_step(const orc_env_cfg *cfg, int H, double *agents, double *global_time, double ax,
                    double ay, const double *human_vxy)
{
    AG(0, F_PX) = AG(0, F_PX) + ax * cfg->time_step;
    AG(0, F_PY) = AG(0, F_PY) + ay * cfg->time_step;
    AG(0, F_VX) = ax; AG(0, F_VY) = ay;
    for (int h = 1; h <= H; ++h) {
        const double hvx = human_vxy[2 * (h - 1)], hvy = human_vxy[2 * (h - 1) + 1];
        AG(h, F_PX) = AG(h, F_PX) + hvx * cfg->time_step;
        AG(h, F_PY) = AG(h, F_PY) + hvy * cfg->time_step;
        AG(h, F_VX) = hvx; AG(h, F_VY) = hvy;
    }
    *global_time += cfg->time_step;
}

/* ===================================================================== */
/* Policy                                                                 */
/* ===================================================================== */

/* cadrl.py:82-102 (holonomic). numpy: np.exp, np.e, np.linspace(0, 2pi, R, endpoint=False),
 * np.cos/np.sin -- glibc double functions, as used here. */
int orc_action_space(double v_pref, int speed_samples, int rotation_samples, double *out_xy)
{
    const double E_ = 2.718281828459045;
    int n = 0;
    out_xy[0] = 0; out_xy[1] = 0; n = 1;
    const double step = (2 * 3.141592653589793 - 0) / rotation_samples; /* linspace step */
    for (int r = 0; r < rotation_samples; ++r) {
        const double rotation = r * step;
        for (int s = 0; s < speed_samples; ++s) {
            const double speed = (exp((double)(s + 1) / speed_samples) - 1) / (E_ - 1) * v_pref;
            out_xy[2 * n] = speed * cos(rotation);
            out_xy[2 * n + 1] = speed * sin(rotation);
            ++n;
        }
    }
    return n;
}

/* cadrl.py:82-102 for any kinematics: holonomic -> (vx, vy) as above; otherwise ActionRot (v, r) with
 * r in np.linspace(-pi/4, pi/4, R) (endpoint included: step = (stop - start) / (R - 1), last sample = stop). */
int orc_action_space_k(double v_pref, int speed_samples, int rotation_samples, int kinematics, double *out)
{
    if (kinematics == ORC_KIN_HOLONOMIC) return orc_action_space(v_pref, speed_samples, rotation_samples, out);
    const double E_ = 2.718281828459045, PI = 3.141592653589793;
    int n = 1;
    out[0] = 0; out[1] = 0;
    const double start = -PI / 4, stop = PI / 4;
    const int div = rotation_samples - 1;
    const double step = div > 0 ? (stop - start) / div : 0.0;
    for (int r = 0; r < rotation_samples; ++r) {
        double rotation = start + r * step;                  /* numpy: arange(0, num) * step + start */
        if (div > 0 && r == rotation_samples - 1) rotation = stop;   /* numpy sets y[-1] = stop */
        for (int sidx = 0; sidx < speed_samples; ++sidx) {
            out[2 * n] = (exp((double)(sidx + 1) / speed_samples) - 1) / (E_ - 1) * v_pref;
            out[2 * n + 1] = rotation;
            ++n;
        }
    }
    return n;
}

/* cadrl.py:217-252 with the theta slot of the 'unicycle' branch (cadrl.py:236-237): theta - rot in float32 */
void orc_rotate_k(const float *s, int kinematics, float *o)
{
    orc_rotate(s, o);
    if (kinematics == ORC_KIN_UNICYCLE) {
        const float dx = s[5] - s[0], dy = s[6] - s[1];
        o[2] = s[8] - atan2f(dy, dx);
    }
}

/* cadrl.py:217-252, torch float32 ops; theta slot is zero (cadrl.py:238-240) */
void orc_rotate(const float *s, float *o)
{
    const float dx = s[5] - s[0], dy = s[6] - s[1];
    const float rot = atan2f(dy, dx);
    const float c = cosf(rot), sn = sinf(rot);
    o[0] = sqrtf(dx * dx + dy * dy);                     /* dg */
    o[1] = s[7];                                         /* v_pref */
    o[2] = 0.0f;                                         /* theta */
    o[3] = s[4];                                         /* radius */
    o[4] = s[2] * c + s[3] * sn;                         /* vx */
    o[5] = s[3] * c - s[2] * sn;                         /* vy */
    o[6] = (s[9] - s[0]) * c + (s[10] - s[1]) * sn;      /* px1 */
    o[7] = (s[10] - s[1]) * c - (s[9] - s[0]) * sn;      /* py1 */
    o[8] = s[11] * c + s[12] * sn;                       /* vx1 */
    o[9] = s[12] * c - s[11] * sn;                       /* vy1 */
    o[10] = s[13];                                       /* radius1 */
    const float ax = s[0] - s[9], ay = s[1] - s[10];
    o[11] = sqrtf(ax * ax + ay * ay);                    /* da */
    o[12] = s[4] + s[13];                                /* radius_sum */
}

/* multi_human_rl.py:65-88 */
double orc_compute_reward(double nav_px, double nav_py, double nav_radius, double nav_gx, double nav_gy,
                          int H, const double *hpx, const double *hpy, const double *hr, double time_step)
{
    double dmin = INFINITY;
    int collision = 0;
    for (int i = 0; i < H; ++i) {
        const double dist = norm2(nav_px - hpx[i], nav_py - hpy[i]) - nav_radius - hr[i];
        if (dist < 0) { collision = 1; break; }
        if (dist < dmin) dmin = dist;
    }
    const int reaching_goal = norm2(nav_px - nav_gx, nav_py - nav_gy) < nav_radius;
    if (collision) return -0.25;
    if (reaching_goal) return 1;
    if (dmin < 0.2) return (dmin - 0.2) * 0.5 * time_step;
    return 0;
}

/* ---- SARL ValueNetwork (sarl.py:9-65, cadrl.py:11-19) ------------------- */

/* nn.Linear with the weight stored transposed ([in][out]) so that the loop over
 * outputs vectorises while every output still accumulates its inputs in index
 * order (acc = 0; acc += w[o][k]*x[k] for k = 0..in-1; acc += b[o]). */
typedef struct { float *wt; const float *b; int in, out; } lin_t;
typedef struct { lin_t m1[2], m2[2], at[3], m3[4]; float *store; } sarl_prep;

static const float *take_linear(const float *p, int in, int out, lin_t *l, float **store)
{
    l->wt = *store; *store += (size_t)in * out;
    for (int o = 0; o < out; ++o)
        for (int k = 0; k < in; ++k) l->wt[(size_t)k * out + o] = p[(size_t)o * in + k];
    l->b = p + (size_t)in * out; l->in = in; l->out = out;
    return p + (size_t)in * out + out;
}

int64_t orc_sarl_param_count(const orc_sarl_cfg *c)
{
    int64_t n = 0;
    int in = c->input_dim;
    for (int i = 0; i < 2; ++i) { n += (int64_t)in * c->mlp1_dims[i] + c->mlp1_dims[i]; in = c->mlp1_dims[i]; }
    in = c->mlp1_dims[1];
    for (int i = 0; i < 2; ++i) { n += (int64_t)in * c->mlp2_dims[i] + c->mlp2_dims[i]; in = c->mlp2_dims[i]; }
    in = c->mlp1_dims[1] * 2;
    for (int i = 0; i < 3; ++i) { n += (int64_t)in * c->attn_dims[i] + c->attn_dims[i]; in = c->attn_dims[i]; }
    in = c->mlp2_dims[1] + c->self_state_dim;
    for (int i = 0; i < 4; ++i) { n += (int64_t)in * c->mlp3_dims[i] + c->mlp3_dims[i]; in = c->mlp3_dims[i]; }
    return n;
}

static void sarl_prepare(const orc_sarl_cfg *c, const float *weights, sarl_prep *P)
{
    P->store = (float *)malloc(sizeof(float) * (size_t)orc_sarl_param_count(c));
    float *st = P->store;
    const float *p = weights;
    int in = c->input_dim;
    for (int i = 0; i < 2; ++i) { p = take_linear(p, in, c->mlp1_dims[i], &P->m1[i], &st); in = c->mlp1_dims[i]; }
    in = c->mlp1_dims[1];
    for (int i = 0; i < 2; ++i) { p = take_linear(p, in, c->mlp2_dims[i], &P->m2[i], &st); in = c->mlp2_dims[i]; }
    in = c->mlp1_dims[1] * 2;
    for (int i = 0; i < 3; ++i) { p = take_linear(p, in, c->attn_dims[i], &P->at[i], &st); in = c->attn_dims[i]; }
    in = c->mlp2_dims[1] + c->self_state_dim;
    for (int i = 0; i < 4; ++i) { p = take_linear(p, in, c->mlp3_dims[i], &P->m3[i], &st); in = c->mlp3_dims[i]; }
}

#define ORC_MAXDIM 512
#define ORC_MAXH 64

__attribute__((target_clones("avx2", "default")))
static void linear_fwd(const lin_t *l, const float *x, float *y, int relu)
{
    float acc[ORC_MAXDIM];
    const int out = l->out;
    for (int o = 0; o < out; ++o) acc[o] = 0.0f;
    for (int k = 0; k < l->in; ++k) {
        const float xk = x[k];
        const float *wr = l->wt + (size_t)k * out;
        for (int o = 0; o < out; ++o) acc[o] += wr[o] * xk;
    }
    for (int o = 0; o < out; ++o) {
        const float a = acc[o] + l->b[o];
        y[o] = (relu && a < 0.0f) ? 0.0f : a;
    }
}

static float sarl_forward_prepared(const orc_sarl_cfg *c, const sarl_prep *P, int H, const float *x, float *attn_out)
{
    const lin_t *m1 = P->m1, *m2 = P->m2, *at = P->at, *m3 = P->m3;
    const int G = c->mlp1_dims[1], F = c->mlp2_dims[1];
    static __thread float m1out[ORC_MAXH][ORC_MAXDIM], feat[ORC_MAXH][ORC_MAXDIM];
    float t0[ORC_MAXDIM], t1[ORC_MAXDIM], glob[ORC_MAXDIM], scores[ORC_MAXH];

    for (int h = 0; h < H; ++h) {
        linear_fwd(&m1[0], x + (size_t)h * c->input_dim, t0, 1);       /* sarl.py:37, last_relu=True */
        linear_fwd(&m1[1], t0, m1out[h], 1);
        linear_fwd(&m2[0], m1out[h], t0, 1);                           /* sarl.py:38 */
        linear_fwd(&m2[1], t0, feat[h], 0);
    }
    for (int k = 0; k < G; ++k) {                                      /* sarl.py:42 mean over humans */
        float s = 0.0f;
        for (int h = 0; h < H; ++h) s += m1out[h][k];
        glob[k] = s / (float)H;
    }
    for (int h = 0; h < H; ++h) {                                      /* sarl.py:45-48 */
        float ain[2 * ORC_MAXDIM];
        memcpy(ain, m1out[h], sizeof(float) * G);
        memcpy(ain + G, glob, sizeof(float) * G);
        linear_fwd(&at[0], ain, t0, 1);
        linear_fwd(&at[1], t0, t1, 1);
        linear_fwd(&at[2], t1, &scores[h], 0);
    }
    /* masked, un-stabilised softmax (sarl.py:52-53) */
    float se[ORC_MAXH], ssum = 0.0f;
    for (int h = 0; h < H; ++h) { se[h] = expf(scores[h]) * (scores[h] != 0.0f ? 1.0f : 0.0f); ssum += se[h]; }
    float joint[ORC_MAXDIM];
    for (int k = 0; k < c->self_state_dim; ++k) joint[k] = x[k];       /* sarl.py:36 */
    for (int k = 0; k < F; ++k) joint[c->self_state_dim + k] = 0.0f;
    for (int h = 0; h < H; ++h) {
        const float w = se[h] / ssum;
        if (attn_out) attn_out[h] = w;
        for (int k = 0; k < F; ++k) joint[c->self_state_dim + k] += w * feat[h][k];  /* sarl.py:60 */
    }
    linear_fwd(&m3[0], joint, t0, 1);                                  /* sarl.py:64 */
    linear_fwd(&m3[1], t0, t1, 1);
    linear_fwd(&m3[2], t1, t0, 1);
    float v;
    linear_fwd(&m3[3], t0, &v, 0);
    return v;
}

float orc_sarl_forward(const orc_sarl_cfg *c, const float *weights, int H, const float *x, float *attn_out)
{
    sarl_prep P;
    sarl_prepare(c, weights, &P);
    const float v = sarl_forward_prepared(c, &P, H, x, attn_out);
    free(P.store);
    return v;
}

void orc_transform_k(int H, const double *agents, int kinematics, double theta, float *out)
{
    for (int h = 1; h <= H; ++h) {
        float row[14];
        row[0] = (float)AG(0, F_PX); row[1] = (float)AG(0, F_PY);
        row[2] = (float)AG(0, F_VX); row[3] = (float)AG(0, F_VY);
        row[4] = (float)AG(0, F_R);  row[5] = (float)AG(0, F_GX);
        row[6] = (float)AG(0, F_GY); row[7] = (float)AG(0, F_VPREF);
        row[8] = (float)theta;
        row[9] = (float)AG(h, F_PX); row[10] = (float)AG(h, F_PY);
        row[11] = (float)AG(h, F_VX); row[12] = (float)AG(h, F_VY);
        row[13] = (float)AG(h, F_R);
        orc_rotate_k(row, kinematics, out + (size_t)(h - 1) * 13);
    }
}

/* multi_human_rl.py:90-104 */
void orc_transform(int H, const double *agents, float *out)
{
    for (int h = 1; h <= H; ++h) {
        float row[14];
        row[0] = (float)AG(0, F_PX); row[1] = (float)AG(0, F_PY);
        row[2] = (float)AG(0, F_VX); row[3] = (float)AG(0, F_VY);
        row[4] = (float)AG(0, F_R);  row[5] = (float)AG(0, F_GX);
        row[6] = (float)AG(0, F_GY); row[7] = (float)AG(0, F_VPREF);
        row[8] = (float)1.5707963267948966; /* robot theta = pi/2 (crowd_sim.py:284); unused */
        row[9] = (float)AG(h, F_PX); row[10] = (float)AG(h, F_PY);
        row[11] = (float)AG(h, F_VX); row[12] = (float)AG(h, F_VY);
        row[13] = (float)AG(h, F_R);
        orc_rotate(row, out + (size_t)(h - 1) * 13);
    }
}

/* multi_human_rl.py:11-63 (greedy branch) */
typedef double (*value_fn)(const void *ctx, int H, const float *x);
static int lookahead_generic_om(const orc_env_cfg *ecfg, value_fn vf, const void *vctx, const int *order,
                  const orc_om_cfg *om, int H,
                  const double *agents, double global_time, int kinematics, double theta, int A, const double *actions,
                  int query_env, const double *human_vxy, double gamma, double *values_out, int *reached);
static int lookahead_generic(const orc_env_cfg *ecfg, value_fn vf, const void *vctx, const int *order, int H,
                  const double *agents, double global_time, int kinematics, double theta, int A, const double *actions,
                  int query_env, const double *human_vxy, double gamma, double *values_out, int *reached)
{
    return lookahead_generic_om(ecfg, vf, vctx, order, NULL, H, agents, global_time, kinematics, theta, A, actions, query_env,
                                human_vxy, gamma, values_out, reached);
}

typedef struct { const orc_sarl_cfg *c; const sarl_prep *P; } sarl_ctx;
static double sarl_value(const void *ctx, int H, const float *x)
{
    const sarl_ctx *s = (const sarl_ctx *)ctx;
    return (double)sarl_forward_prepared(s->c, s->P, H, x, NULL);
}

static int lookahead_prepared(const orc_env_cfg *ecfg, const orc_sarl_cfg *scfg, const sarl_prep *P, int H,
                  const double *agents, double global_time, int kinematics, double theta, int A, const double *actions,
                  int query_env, const double *human_vxy, double gamma, double *values_out, int *reached)
{
    const sarl_ctx ctx = {scfg, P};
    return lookahead_generic(ecfg, sarl_value, &ctx, NULL, H, agents, global_time, kinematics, theta, A, actions, query_env,
                             human_vxy, gamma, values_out, reached);
}

/* multi_human_rl.py:109-163 */
void orc_occupancy_maps(const orc_om_cfg *om, int H, const double *humans, float *out)
{
    const int cn = om->cell_num, cells = cn * cn, ch = om->om_channel_size;
    for (int i = 0; i < H; ++i) {
        const double *hi = humans + 4 * i;
        double cnt[64], sx[64], sy[64];
        for (int c = 0; c < cells; ++c) cnt[c] = sx[c] = sy[c] = 0.0;
        const double angle = atan2(hi[3], hi[2]);                       /* human_velocity_angle */
        for (int j = 0; j < H; ++j) {
            if (j == i) continue;
            const double *hj = humans + 4 * j;
            const double opx = hj[0] - hi[0], opy = hj[1] - hi[1];
            const double rotation = atan2(opy, opx) - angle;
            const double distance = sqrt(opx * opx + opy * opy);       /* np.linalg.norm([px, py], axis=0) */
            const double rx = cos(rotation) * distance, ry = sin(rotation) * distance;
            const double xi = floor(rx / om->cell_size + cn / 2.0), yi = floor(ry / om->cell_size + cn / 2.0);
            if (xi < 0 || xi >= cn || yi < 0 || yi >= cn) continue;    /* -inf index: in no cell */
            const int idx = (int)(cn * yi + xi);
            const double vrot = atan2(hj[3], hj[2]) - angle;
            const double speed = sqrt(hj[2] * hj[2] + hj[3] * hj[3]);  /* np.linalg.norm(v, axis=1) */
            cnt[idx] += 1.0;
            sx[idx] += cos(vrot) * speed;                               /* python sum(): left to right, start 0 */
            sy[idx] += sin(vrot) * speed;
        }
        float *o = out + (size_t)i * cells * ch;
        for (int c = 0; c < cells; ++c) {
            if (ch == 1) o[c] = cnt[c] > 0 ? 1.0f : 0.0f;
            else if (ch == 2) { o[2 * c] = cnt[c] > 0 ? (float)(sx[c] / cnt[c]) : 0.0f; o[2 * c + 1] = cnt[c] > 0 ? (float)(sy[c] / cnt[c]) : 0.0f; }
            else {
                o[3 * c] = cnt[c] > 0 ? 1.0f : 0.0f;                    /* sum([1, 1, ..]) / len */
                o[3 * c + 1] = cnt[c] > 0 ? (float)(sx[c] / cnt[c]) : 0.0f;
                o[3 * c + 2] = cnt[c] > 0 ? (float)(sy[c] / cnt[c]) : 0.0f;
            }
        }
    }
}

/* order (may be NULL = env order): position j of the network input is human order[j] (0-based) */
static int lookahead_generic_om(const orc_env_cfg *ecfg, value_fn vf, const void *vctx, const int *order,
                  const orc_om_cfg *om, int H,
                  const double *agents, double global_time, int kinematics, double theta, int A, const double *actions,
                  int query_env, const double *human_vxy, double gamma, double *values_out, int *reached)
{
    const int om_dim = om ? om->cell_num * om->cell_num * om->om_channel_size : 0, in_dim = 13 + om_dim;
    static __thread float omap[ORC_MAXH * 64 * 3];
    int om_ready = 0;
    if (reached) *reached = 0;
    /* policy.py:43-49 reach_destination: norm((py-gy, px-gx)) < radius */
    if (norm2(AG(0, F_PY) - AG(0, F_GY), AG(0, F_PX) - AG(0, F_GX)) < AG(0, F_R)) {
        if (reached) *reached = 1;
        return 0;
    }
    const double dt = ecfg->time_step;
    const double gamma_bar = pow(gamma, dt * AG(0, F_VPREF));           /* multi_human_rl.py:52 */
    double max_value = -INFINITY;
    int max_action = -1;
    double nhx[ORC_MAXH], nhy[ORC_MAXH], nhvx[ORC_MAXH], nhvy[ORC_MAXH], hr[ORC_MAXH];
    static __thread float x[ORC_MAXH * (13 + 64 * 3)];
    for (int a = 0; a < A; ++a) {
        /* propagate robot (cadrl.py:113-125): non-holonomic actions are (v, r), next_theta = theta + r */
        double ax, ay;
        effective_velocity(kinematics, theta, actions[2 * a], actions[2 * a + 1], &ax, &ay);
        const double next_theta = kinematics == ORC_KIN_HOLONOMIC ? theta : theta + actions[2 * a + 1];
        const double npx = AG(0, F_PX) + ax * dt, npy = AG(0, F_PY) + ay * dt;
        double reward;
        for (int j = 1; j <= H; ++j) {
            const int h = order ? order[j - 1] + 1 : j;
            double hvx, hvy;
            if (query_env) { hvx = human_vxy[2 * (h - 1)]; hvy = human_vxy[2 * (h - 1) + 1]; } /* agent.py:63-74 */
            else { hvx = AG(h, F_VX); hvy = AG(h, F_VY); }                                    /* cadrl.py:107-109 */
            nhx[j - 1] = AG(h, F_PX) + hvx * dt;
            nhy[j - 1] = AG(h, F_PY) + hvy * dt;
            nhvx[j - 1] = hvx; nhvy[j - 1] = hvy; hr[j - 1] = AG(h, F_R);
        }
        if (query_env) {
            int done, info;
            orc_step_outcome(ecfg, H, agents, global_time, ax, ay, &reward, &done, &info, NULL);
        } else {
            reward = orc_compute_reward(npx, npy, AG(0, F_R), AG(0, F_GX), AG(0, F_GY), H, nhx, nhy, hr, dt);
        }
        for (int h = 0; h < H; ++h) {
            float row[14];
            row[0] = (float)npx; row[1] = (float)npy; row[2] = (float)ax; row[3] = (float)ay;
            row[4] = (float)AG(0, F_R); row[5] = (float)AG(0, F_GX); row[6] = (float)AG(0, F_GY);
            row[7] = (float)AG(0, F_VPREF); row[8] = (float)next_theta;
            row[9] = (float)nhx[h]; row[10] = (float)nhy[h]; row[11] = (float)nhvx[h];
            row[12] = (float)nhvy[h]; row[13] = (float)hr[h];
            orc_rotate_k(row, kinematics, x + (size_t)h * in_dim);
        }
        if (om) {
            if (!om_ready) {        /* built once, from the first action's next human states (multi_human_rl.py:47-49) */
                double hs[ORC_MAXH * 4];
                for (int h = 0; h < H; ++h) { hs[4 * h] = nhx[h]; hs[4 * h + 1] = nhy[h]; hs[4 * h + 2] = nhvx[h]; hs[4 * h + 3] = nhvy[h]; }
                orc_occupancy_maps(om, H, hs, omap);
                om_ready = 1;
            }
            for (int h = 0; h < H; ++h) memcpy(x + (size_t)h * in_dim + 13, omap + (size_t)h * om_dim, sizeof(float) * om_dim);
        }
        const double v = vf(vctx, H, x);
        const double value = reward + gamma_bar * v;
        values_out[a] = value;
        if (value > max_value) { max_value = value; max_action = a; }
    }
    return max_action;
}

int orc_lookahead(const orc_env_cfg *ecfg, const orc_sarl_cfg *scfg, const float *weights, int H,
                  const double *agents, double global_time, int A, const double *actions, int query_env,
                  const double *human_vxy, double gamma, double *values_out, int *reached)
{
    sarl_prep P;
    sarl_prepare(scfg, weights, &P);
    const int r = lookahead_prepared(ecfg, scfg, &P, H, agents, global_time, ORC_KIN_HOLONOMIC, 0.0, A, actions,
                                     query_env, human_vxy, gamma, values_out, reached);
    free(P.store);
    return r;
}

int orc_lookahead_k(const orc_env_cfg *ecfg, const orc_sarl_cfg *scfg, const float *weights, int H,
                    const double *agents, double global_time, int kinematics, double theta, int A,
                    const double *actions, int query_env, const double *human_vxy, double gamma, double *values_out,
                    int *reached)
{
    sarl_prep P;
    sarl_prepare(scfg, weights, &P);
    const int r = lookahead_prepared(ecfg, scfg, &P, H, agents, global_time, kinematics, theta, A, actions, query_env,
                                     human_vxy, gamma, values_out, reached);
    free(P.store);
    return r;
}

/* ---- CADRL / LSTM-RL value networks --------------------------------------------------------- */
typedef struct { lin_t m1[4], m[4], ih, hh; int n1; float *store; } net_prep;

int64_t orc_net_param_count(const orc_net_cfg *c)
{
    int64_t n = 0;
    int in = c->input_dim;
    if (c->network == ORC_NET_CADRL) {
        for (int i = 0; i < 4; ++i) { n += (int64_t)in * c->mlp_dims[i] + c->mlp_dims[i]; in = c->mlp_dims[i]; }
        return n;
    }
    int lstm_in = c->input_dim;
    if (c->mlp1_dims[0] > 0) {
        for (int i = 0; i < 4; ++i) { n += (int64_t)in * c->mlp1_dims[i] + c->mlp1_dims[i]; in = c->mlp1_dims[i]; }
        lstm_in = c->mlp1_dims[3];
    }
    in = c->self_state_dim + c->lstm_hidden;
    for (int i = 0; i < 4; ++i) { n += (int64_t)in * c->mlp_dims[i] + c->mlp_dims[i]; in = c->mlp_dims[i]; }
    const int G = 4 * c->lstm_hidden;
    n += (int64_t)G * lstm_in + (int64_t)G * c->lstm_hidden + 2 * G;
    return n;
}

static void net_prepare(const orc_net_cfg *c, const float *weights, net_prep *P)
{
    P->store = (float *)malloc(sizeof(float) * (size_t)orc_net_param_count(c));
    float *st = P->store;
    const float *p = weights;
    int in = c->input_dim;
    P->n1 = 0;
    if (c->network == ORC_NET_CADRL) {
        for (int i = 0; i < 4; ++i) { p = take_linear(p, in, c->mlp_dims[i], &P->m[i], &st); in = c->mlp_dims[i]; }
        return;
    }
    int lstm_in = c->input_dim;
    if (c->mlp1_dims[0] > 0) {
        P->n1 = 4;
        for (int i = 0; i < 4; ++i) { p = take_linear(p, in, c->mlp1_dims[i], &P->m1[i], &st); in = c->mlp1_dims[i]; }
        lstm_in = c->mlp1_dims[3];
    }
    in = c->self_state_dim + c->lstm_hidden;
    for (int i = 0; i < 4; ++i) { p = take_linear(p, in, c->mlp_dims[i], &P->m[i], &st); in = c->mlp_dims[i]; }
    /* nn.LSTM parameters: weight_ih_l0 [4h][in], weight_hh_l0 [4h][h], bias_ih_l0 [4h], bias_hh_l0 [4h] */
    const int G = 4 * c->lstm_hidden, Hh = c->lstm_hidden;
    const float *w_ih = p, *w_hh = p + (size_t)G * lstm_in, *b_ih = w_hh + (size_t)G * Hh, *b_hh = b_ih + G;
    P->ih.wt = st; st += (size_t)G * lstm_in;
    for (int o = 0; o < G; ++o) for (int k = 0; k < lstm_in; ++k) P->ih.wt[(size_t)k * G + o] = w_ih[(size_t)o * lstm_in + k];
    P->ih.b = b_ih; P->ih.in = lstm_in; P->ih.out = G;
    P->hh.wt = st; st += (size_t)G * Hh;
    for (int o = 0; o < G; ++o) for (int k = 0; k < Hh; ++k) P->hh.wt[(size_t)k * G + o] = w_hh[(size_t)o * Hh + k];
    P->hh.b = b_hh; P->hh.in = Hh; P->hh.out = G;
}

static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

static void net_forward_prepared(const orc_net_cfg *c, const net_prep *P, int H, const float *x, float *out)
{
    float t0[ORC_MAXDIM], t1[ORC_MAXDIM];
    if (c->network == ORC_NET_CADRL) {                                  /* cadrl.py:22-30 */
        for (int h = 0; h < H; ++h) {
            linear_fwd(&P->m[0], x + (size_t)h * c->input_dim, t0, 1);
            linear_fwd(&P->m[1], t0, t1, 1);
            linear_fwd(&P->m[2], t1, t0, 1);
            linear_fwd(&P->m[3], t0, &out[h], 0);
        }
        return;
    }
    const int Hh = c->lstm_hidden;
    float hs[ORC_MAXDIM], cs[ORC_MAXDIM], gi[ORC_MAXDIM], gh[ORC_MAXDIM];
    for (int k = 0; k < Hh; ++k) hs[k] = cs[k] = 0.0f;                  /* lstm_rl.py:27-28 */
    for (int t = 0; t < H; ++t) {
        const float *xt = x + (size_t)t * c->input_dim;
        if (P->n1) {                                                    /* ValueNetwork2: mlp1, no last ReLU (lstm_rl.py:56) */
            linear_fwd(&P->m1[0], xt, t0, 1);
            linear_fwd(&P->m1[1], t0, t1, 1);
            linear_fwd(&P->m1[2], t1, t0, 1);
            linear_fwd(&P->m1[3], t0, t1, 0);
            xt = t1;
        }
        linear_fwd(&P->ih, xt, gi, 0);
        linear_fwd(&P->hh, hs, gh, 0);
        for (int k = 0; k < Hh; ++k) {                                  /* gate order i, f, g, o */
            const float ig = sigmoidf_(gi[k] + gh[k]), fg = sigmoidf_(gi[Hh + k] + gh[Hh + k]);
            const float gg = tanhf(gi[2 * Hh + k] + gh[2 * Hh + k]), og = sigmoidf_(gi[3 * Hh + k] + gh[3 * Hh + k]);
            cs[k] = fg * cs[k] + ig * gg;
            hs[k] = og * tanhf(cs[k]);
        }
    }
    float joint[ORC_MAXDIM];
    for (int k = 0; k < c->self_state_dim; ++k) joint[k] = x[k];        /* lstm_rl.py:25,31 */
    for (int k = 0; k < Hh; ++k) joint[c->self_state_dim + k] = hs[k];
    linear_fwd(&P->m[0], joint, t0, 1);
    linear_fwd(&P->m[1], t0, t1, 1);
    linear_fwd(&P->m[2], t1, t0, 1);
    linear_fwd(&P->m[3], t0, &out[0], 0);
}

void orc_net_forward(const orc_net_cfg *c, const float *weights, int H, const float *x, float *out)
{
    net_prep P;
    net_prepare(c, weights, &P);
    net_forward_prepared(c, &P, H, x, out);
    free(P.store);
}

typedef struct { const orc_net_cfg *c; const net_prep *P; } net_ctx;
static double net_value(const void *ctx, int H, const float *x)
{
    const net_ctx *n = (const net_ctx *)ctx;
    float out[ORC_MAXH];
    net_forward_prepared(n->c, n->P, H, x, out);
    if (n->c->network == ORC_NET_CADRL) {                               /* cadrl.py:165: torch.min(outputs, 0) */
        float m = out[0];
        for (int h = 1; h < H; ++h) if (out[h] < m) m = out[h];
        return (double)m;
    }
    return (double)out[0];
}

/* lstm_rl.py:99-104: sorted(human_states, key=dist, reverse=True) -- stable, ties keep the env order */
void orc_lstm_human_order(int H, const double *agents, int *order)
{
    double d[ORC_MAXH];
    for (int h = 0; h < H; ++h) {
        d[h] = norm2(AG(h + 1, F_PX) - AG(0, F_PX), AG(h + 1, F_PY) - AG(0, F_PY));
        order[h] = h;
    }
    for (int i = 1; i < H; ++i) {                                       /* stable insertion sort, descending */
        const int oi = order[i];
        int j = i - 1;
        while (j >= 0 && d[order[j]] < d[oi]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = oi;
    }
}

int orc_lookahead_net(const orc_env_cfg *ecfg, const orc_net_cfg *ncfg, const float *weights, int H,
                      const double *agents, double global_time, int kinematics, double theta, int A,
                      const double *actions, int query_env, const double *human_vxy, double gamma, double *values_out,
                      int *reached)
{
    net_prep P;
    net_prepare(ncfg, weights, &P);
    const net_ctx ctx = {ncfg, &P};
    int order[ORC_MAXH];
    const int sorted = ncfg->network == ORC_NET_LSTM_RL && !query_env;
    if (sorted) orc_lstm_human_order(H, agents, order);
    const int r = lookahead_generic(ecfg, net_value, &ctx, sorted ? order : NULL, H, agents, global_time, kinematics, theta,
                                    A, actions, query_env, human_vxy, gamma, values_out, reached);
    free(P.store);
    return r;
}

int orc_lookahead_om(const orc_env_cfg *ecfg, int net, const orc_sarl_cfg *scfg, const orc_net_cfg *ncfg,
                     const orc_om_cfg *om, const float *weights, int H, const double *agents, double global_time,
                     int kinematics, double theta, int A, const double *actions, int query_env, const double *human_vxy,
                     double gamma, double *values_out, int *reached)
{
    int r;
    if (net == ORC_NET_SARL) {
        sarl_prep P;
        sarl_prepare(scfg, weights, &P);
        const sarl_ctx ctx = {scfg, &P};
        r = lookahead_generic_om(ecfg, sarl_value, &ctx, NULL, om, H, agents, global_time, kinematics, theta, A, actions,
                                 query_env, human_vxy, gamma, values_out, reached);
        free(P.store);
    } else {
        net_prep P;
        net_prepare(ncfg, weights, &P);
        const net_ctx ctx = {ncfg, &P};
        int order[ORC_MAXH];
        const int sorted = net == ORC_NET_LSTM_RL && !query_env;
        if (sorted) orc_lstm_human_order(H, agents, order);
        r = lookahead_generic_om(ecfg, net_value, &ctx, sorted ? order : NULL, om, H, agents, global_time, kinematics, theta,
                                 A, actions, query_env, human_vxy, gamma, values_out, reached);
        free(P.store);
    }
    return r;
}

void orc_transform_om(const orc_om_cfg *om, int H, const double *agents, const int *order, int kinematics, double theta,
                      float *out)
{
    const int om_dim = om->cell_num * om->cell_num * om->om_channel_size, in_dim = 13 + om_dim;
    float rows[ORC_MAXH * 13];
    static __thread float omap[ORC_MAXH * 64 * 3];
    double hs[ORC_MAXH * 4];
    orc_transform_k(H, agents, kinematics, theta, rows);
    for (int j = 0; j < H; ++j) {
        const int h = (order ? order[j] : j) + 1;
        hs[4 * j] = AG(h, F_PX); hs[4 * j + 1] = AG(h, F_PY); hs[4 * j + 2] = AG(h, F_VX); hs[4 * j + 3] = AG(h, F_VY);
    }
    orc_occupancy_maps(om, H, hs, omap);
    for (int j = 0; j < H; ++j) {
        const int h = order ? order[j] : j;
        memcpy(out + (size_t)j * in_dim, rows + (size_t)h * 13, sizeof(float) * 13);
        memcpy(out + (size_t)j * in_dim + 13, omap + (size_t)j * om_dim, sizeof(float) * om_dim);
    }
}

void orc_batch_lookahead_step(const orc_env_cfg *ecfg, const orc_sarl_cfg *scfg, const float *weights,
                              int E, int H, double *agents_all, double *global_time, int A,
                              const double *actions, int query_env, double gamma, int32_t *action_idx,
                              double *reward, uint8_t *done, uint8_t *info, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    sarl_prep P;
    sarl_prepare(scfg, weights, &P);
#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 1)
    for (int e = 0; e < E; ++e) {
        if (done[e]) continue;
        double *agents = agents_all + (size_t)e * (H + 1) * ORC_AGENT_STRIDE;
        double hv[ORC_MAXH * 2], values[ORC_NUM_ACTIONS_MAX];
        orc_human_actions(ecfg, H, agents, hv);
        int reached;
        int best = lookahead_prepared(ecfg, scfg, &P, H, agents, global_time[e], ORC_KIN_HOLONOMIC, 0.0, A, actions, query_env,
                                      hv, gamma, values, &reached);
        if (best < 0) best = 0;
        action_idx[e] = best;
        double r; int d, inf;
        orc_step_outcome(ecfg, H, agents, global_time[e], actions[2 * best], actions[2 * best + 1], &r, &d, &inf, NULL);
        orc_apply_step(ecfg, H, agents, &global_time[e], actions[2 * best], actions[2 * best + 1], hv);
        reward[e] = r; done[e] = (uint8_t)d; info[e] = (uint8_t)inf;
    }
    free(P.store);
}
