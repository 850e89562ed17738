// scenes_host.cu -- CrowdSim.reset's scene generator on the HOST, natively (crowd_sim/envs/crowd_sim.py:165-217,261-323):
// numpy's legacy MT19937 stream seeded per case (np.random.seed(counter_offset + case), crowd_sim.py:286), the same draws in
// the same order, the same double arithmetic (np.linalg.norm of a pair = sqrt(fma(b, b, a * a)), glibc cos / sin), so the
// scenes are bit-identical to the reference's -- what modelcrowdnav_b200/scenes.py does in Python at ~0.2 ms per scene.
// A batched roll-out of 512 training episodes needs its 512 scenes in well under a millisecond; this is that path.
#include "cn_common.cuh"

#include <math.h>

#include <thread>
#include <vector>

namespace {

struct MT19937 {
    uint32_t key[624];
    int pos;
    explicit MT19937(uint32_t seed)
    {
        for (int i = 0; i < 624; ++i) {
            key[i] = seed;
            seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
        }
        pos = 624;
    }
    uint32_t next32()
    {
        if (pos == 624) {
            int k;
            uint32_t y;
            for (k = 0; k < 624 - 397; ++k) {
                y = (key[k] & 0x80000000u) | (key[k + 1] & 0x7fffffffu);
                key[k] = key[k + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            for (; k < 623; ++k) {
                y = (key[k] & 0x80000000u) | (key[k + 1] & 0x7fffffffu);
                key[k] = key[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            y = (key[623] & 0x80000000u) | (key[0] & 0x7fffffffu);
            key[623] = key[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            pos = 0;
        }
        uint32_t y = key[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    double random_sample()      // numpy legacy_double: 53-bit resolution
    {
        const int32_t a = (int32_t)(next32() >> 5), b = (int32_t)(next32() >> 6);
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
};

inline double norm2(double a, double b) { return sqrt(fma(b, b, a * a)); }

void one_scene(uint32_t seed, int H, int rule, double circle_radius, double square_width, double base_radius, double base_v_pref,
               double discomfort, double robot_radius, double robot_v_pref, int randomize, double *A /*(H+1) x 8*/)
{
    MT19937 rs(seed);
    const double PI = 3.141592653589793;
    double *r = A;
    r[0] = 0; r[1] = -circle_radius; r[2] = 0; r[3] = 0; r[4] = 0; r[5] = circle_radius; r[6] = robot_radius; r[7] = robot_v_pref;
    for (int i = 1; i <= H; ++i) {
        double radius = base_radius, v_pref = base_v_pref;
        if (randomize) {                                           // agent.py:39-45: uniform(lo, hi) = lo + (hi - lo) * sample
            v_pref = 0.5 + (1.5 - 0.5) * rs.random_sample();
            radius = 0.3 + (0.5 - 0.3) * rs.random_sample();
        }
        double *a = A + (size_t)i * 8;
        if (rule == CN_CIRCLE_CROSSING) {                          // crowd_sim.py:165-186
            double px, py;
            while (true) {
                const double angle = rs.random_sample() * PI * 2;
                const double px_noise = (rs.random_sample() - 0.5) * v_pref;
                const double py_noise = (rs.random_sample() - 0.5) * v_pref;
                px = circle_radius * cos(angle) + px_noise;
                py = circle_radius * sin(angle) + py_noise;
                bool collide = false;
                for (int j = 0; j < i; ++j) {
                    const double *q = A + (size_t)j * 8;
                    const double min_dist = radius + q[6] + discomfort;
                    if (norm2(px - q[0], py - q[1]) < min_dist || norm2(px - q[4], py - q[5]) < min_dist) { collide = true; break; }
                }
                if (!collide) break;
            }
            a[0] = px; a[1] = py; a[2] = 0; a[3] = 0; a[4] = -px; a[5] = -py; a[6] = radius; a[7] = v_pref;
        } else {                                                   // crowd_sim.py:188-217
            const double sign = rs.random_sample() > 0.5 ? -1.0 : 1.0;
            double px, py, gx, gy;
            while (true) {
                px = rs.random_sample() * square_width * 0.5 * sign;
                py = (rs.random_sample() - 0.5) * square_width;
                bool collide = false;
                for (int j = 0; j < i && !collide; ++j) {
                    const double *q = A + (size_t)j * 8;
                    collide = norm2(px - q[0], py - q[1]) < radius + q[6] + discomfort;
                }
                if (!collide) break;
            }
            while (true) {
                gx = rs.random_sample() * square_width * 0.5 * -sign;
                gy = (rs.random_sample() - 0.5) * square_width;
                bool collide = false;
                for (int j = 0; j < i && !collide; ++j) {
                    const double *q = A + (size_t)j * 8;
                    collide = norm2(gx - q[4], gy - q[5]) < radius + q[6] + discomfort;
                }
                if (!collide) break;
            }
            a[0] = px; a[1] = py; a[2] = 0; a[3] = 0; a[4] = gx; a[5] = gy; a[6] = radius; a[7] = v_pref;
        }
    }
}

}  // namespace

extern "C" int cn_scenes_generate(int32_t n, const int64_t *seeds, int32_t human_num, int32_t rule, double circle_radius,
                                  double square_width, double human_radius, double human_v_pref, double discomfort_dist,
                                  double robot_radius, double robot_v_pref, int32_t randomize_attributes, double *agents_out)
{
    if (n < 0 || !seeds || !agents_out) { cn_set_error("null argument"); return CN_EINVAL; }
    if (human_num < 1 || human_num > CN_MAX_HUMANS) { cn_set_error("1 <= human_num <= %d required", CN_MAX_HUMANS); return CN_EINVAL; }
    if (rule != CN_CIRCLE_CROSSING && rule != CN_SQUARE_CROSSING) { cn_set_error("rule must be circle or square crossing"); return CN_EINVAL; }
    for (int32_t i = 0; i < n; ++i)
        if (seeds[i] < 0 || seeds[i] > 4294967295LL) { cn_set_error("Seed must be between 0 and 2**32 - 1"); return CN_EINVAL; }
    const size_t stride = (size_t)(human_num + 1) * 8;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt > 16) nt = 16;
    if (nt < 1 || n < 64) nt = 1;
    auto work = [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i)
            one_scene((uint32_t)seeds[i], human_num, rule, circle_radius, square_width, human_radius, human_v_pref, discomfort_dist,
                      robot_radius, robot_v_pref, randomize_attributes, agents_out + (size_t)i * stride);
    };
    if (nt == 1) { work(0, n); return CN_OK; }
    std::vector<std::thread> th;
    const int per = (n + (int)nt - 1) / (int)nt;
    for (unsigned t = 0; t < nt; ++t) {
        const int lo = (int)t * per, hi = lo + per < n ? lo + per : n;
        if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto &t : th) t.join();
    return CN_OK;
}
