// env_math.cuh -- float64 reward/collision arithmetic shared by the step and lookahead kernels.
// Every translation unit that includes this file is compiled with -fmad=false (bit-exact with CPython).
#pragma once
#include "cn_common.cuh"

// crowd_sim/envs/utils/utils.py:4-26 with (x3, y3) = (0, 0)
__device__ __forceinline__ double cn_point_to_segment_dist0(double x1, double y1, double x2, double y2)
{
    const double px = x2 - x1, py = y2 - y1;
    if (px == 0 && py == 0) return norm2d(0.0 - x1, 0.0 - y1);
    double u = ((0.0 - x1) * px + (0.0 - y1) * py) / (px * px + py * py);
    if (u > 1) u = 1;
    else if (u < 0) u = 0;
    const double x = x1 + u * px, y = y1 + u * py;
    return norm2d(x - 0.0, y - 0.0);
}

struct StepOutcome {
    double reward, dmin;
    int done, info;
};

// Collision / goal / reward ladder of CrowdSim.step (crowd_sim.py:344-403).  `ag(field, agent)` returns
// the f64 state of the env being evaluated.
template <typename Acc>
__device__ __forceinline__ StepOutcome cn_step_outcome(const EnvParams &p, Acc ag, int H, double t, double ax, double ay)
{
    const double rpx = ag(F_PX, 0), rpy = ag(F_PY, 0), rr = ag(F_R, 0);
    const double dt = p.time_step;
    double dmin = INFINITY;
    bool collision = false;
    for (int h = 1; h <= H; ++h) {
        const double px = ag(F_PX, h) - rpx;
        const double py = ag(F_PY, h) - rpy;
        const double vx = ag(F_VX, h) - ax;
        const double vy = ag(F_VY, h) - ay;
        const double ex = px + vx * dt;
        const double ey = py + vy * dt;
        const double closest = cn_point_to_segment_dist0(px, py, ex, ey) - ag(F_R, h) - rr;
        // crowd_sim.py:353-358 breaks at the first collision; here the loop runs on (no early exit: the loads of the next
        // humans need not wait for this human's verdict) and simply stops updating -- same collision flag, same dmin
        if (!collision) {
            if (closest < 0) collision = true;
            else if (closest < dmin) dmin = closest;
        }
    }
    const double endx = rpx + ax * dt, endy = rpy + ay * dt;
    const bool reaching_goal = norm2d(endx - ag(F_GX, 0), endy - ag(F_GY, 0)) < rr;
    StepOutcome o;
    o.dmin = dmin;
    if (t >= p.time_limit - 1) { o.reward = 0; o.done = 1; o.info = CN_TIMEOUT; }
    else if (collision) { o.reward = p.collision_penalty; o.done = 1; o.info = CN_COLLISION; }
    else if (reaching_goal) { o.reward = p.success_reward; o.done = 1; o.info = CN_REACHGOAL; }
    else if (dmin < p.discomfort_dist) {
        o.reward = (dmin - p.discomfort_dist) * p.discomfort_penalty_factor * dt;
        o.done = 0; o.info = CN_DANGER;
    } else { o.reward = 0; o.done = 0; o.info = CN_NOTHING; }
    return o;
}

// CADRL.rotate on one joint-state row (cadrl.py:217-252), float32; theta slot = theta - rot only for
// kinematics == 'unicycle' (cadrl.py:236-240), else 0.
// s: px py vx vy radius gx gy v_pref theta px1 py1 vx1 vy1 radius1
__device__ __forceinline__ void cn_rotate(const float *s, float *o, int kinematics = CN_KIN_HOLONOMIC)
{
    const float dx = s[5] - s[0], dy = s[6] - s[1];
    const float rot = atan2f(dy, dx);
    float sn, c;
    sincosf(rot, &sn, &c);
    o[0] = sqrtf(dx * dx + dy * dy);
    o[1] = s[7];
    o[2] = kinematics == CN_KIN_UNICYCLE ? s[8] - rot : 0.0f;
    o[3] = s[4];
    o[4] = s[2] * c + s[3] * sn;
    o[5] = s[3] * c - s[2] * sn;
    o[6] = (s[9] - s[0]) * c + (s[10] - s[1]) * sn;
    o[7] = (s[10] - s[1]) * c - (s[9] - s[0]) * sn;
    o[8] = s[11] * c + s[12] * sn;
    o[9] = s[12] * c - s[11] * sn;
    o[10] = s[13];
    const float ax = s[0] - s[9], ay = s[1] - s[10];
    o[11] = sqrtf(ax * ax + ay * ay);
    o[12] = s[4] + s[13];
}
