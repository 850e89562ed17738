// lookahead_tc.cu -- SARL one-step lookahead on tcgen05 tensor cores (CN_PREC_F16_TC).
//
// fp16 operands, fp32 accumulation in TMEM, M = 128 rows per UMMA.  Two persistent kernels (1 CTA / SM):
//
//  tc_rows_kernel   rows = (env, action, human).  Per 128-row tile, all on chip:
//       propagate + reward + rotate  -> X (13 features, split hi+lo fp16 so layer 1 sees ~22-bit inputs)
//       mlp1.0 -> mlp1.2 -> {mlp2.0 -> mlp2.2 ; attention.0 on [mlp1_out | group mean] -> attention.2}
//       attention.4 (fp32 dot) -> masked softmax over the humans of a group -> weighted feature
//       -> joint state J (fp16, 160 B per (env, action)) to HBM, already in UMMA operand order
//     All six GEMM stages chain smem -> tcgen05.mma -> TMEM -> tcgen05.ld -> smem; the full fp16 weight set
//     of the row network (157 KB) stays resident in shared memory for the life of the CTA.
//  tc_mlp3_kernel   rows = (env, action).  mlp3.0 -> .2 -> .4 on tensor cores, .6 as an fp32 dot,
//       value = reward + gamma_bar * V  -> values[E][A]; the shared argmax kernel follows.
//
// Biases ride inside the GEMMs: every activation tile carries two constant-one columns and the weight
// images hold bias_hi / bias_lo (fp16 split) in the matching K rows, so epilogues are pure
// tcgen05.ld -> cvt.rn.relu.f16x2 -> st.shared.
//
// Reference (file:line relative to the reference root): see lookahead_f32.cu; the network is
// crowd_nav/policy/sarl.py:28-65 with the default [sarl] dims of crowd_nav/configs/policy.config:41-49.
#include "env_math.cuh"
#include "umma.cuh"

#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

int cn_lookahead_prepare(cn_policy *p, cn_env *env, cudaStream_t s);
int cn_lookahead_argmax(cn_policy *p, cn_env *env, double epsilon, cudaStream_t s);

namespace {

using namespace umma;

// ---- padded GEMM shapes (default SARL dims) ---------------------------------------------------------
constexpr int K_X = 32;     // 13 hi | 1 1 0 | 13 lo | 0 0 0
constexpr int N_H1 = 160;   // 150 + ones(150,151)
constexpr int N_M1 = 112;   // 100 + ones(100,101)
constexpr int N_F = 64;     // 50
constexpr int K_A1 = 224;   // [mlp1_out 112 | group mean 112]
constexpr int K_J = 80;     // 50 weighted | 6 pad | 6 self_hi 1 1 | 6 self_lo 0 0 | 8 pad
constexpr int ROWS = 128;

__host__ __device__ constexpr uint32_t bytes_of(int rows, int K) { return (uint32_t)rows * K * 2; }
// weight image of tc_rows_kernel
constexpr uint32_t OFF_W1 = 0;
constexpr uint32_t OFF_W2 = OFF_W1 + bytes_of(N_H1, K_X);
constexpr uint32_t OFF_W3 = OFF_W2 + bytes_of(N_M1, N_H1);
constexpr uint32_t OFF_W4 = OFF_W3 + bytes_of(N_M1, N_M1);
constexpr uint32_t OFF_WA1 = OFF_W4 + bytes_of(N_F, N_M1);
constexpr uint32_t OFF_WA2 = OFF_WA1 + bytes_of(N_M1, K_A1);
constexpr uint32_t OFF_TAILA = OFF_WA2 + bytes_of(N_M1, N_M1);   // fp32: attention.4 weight[100], bias
constexpr uint32_t TAIL_BYTES = 416;
constexpr uint32_t IMG_A_BYTES = OFF_TAILA + TAIL_BYTES;
// weight image of tc_mlp3_kernel
constexpr uint32_t OFF_M1 = 0;
constexpr uint32_t OFF_M2 = OFF_M1 + bytes_of(N_H1, K_J);
constexpr uint32_t OFF_M3 = OFF_M2 + bytes_of(N_M1, N_H1);
constexpr uint32_t OFF_TAILB = OFF_M3 + bytes_of(N_M1, N_M1);    // fp32: mlp3.6 weight[100], bias
constexpr uint32_t IMG_B_BYTES = OFF_TAILB + TAIL_BYTES;

// shared memory maps
constexpr uint32_t A_BUFA = (IMG_A_BYTES + 127) & ~127u;                 // 128 x 160 fp16
constexpr uint32_t A_BUFB = A_BUFA + bytes_of(ROWS, N_H1);               // 128 x 112 fp16
constexpr uint32_t A_MISC = A_BUFB + bytes_of(ROWS, N_M1);               // S[128] f32, mbar, tmem ptr
constexpr uint32_t A_SMEM = A_MISC + 1024 + 16;
constexpr uint32_t B_BUFJ = (IMG_B_BYTES + 127) & ~127u;                 // 128 x 80 fp16
constexpr uint32_t B_BUFU0 = B_BUFJ + bytes_of(ROWS, K_J);
constexpr uint32_t B_BUFU1 = B_BUFU0 + bytes_of(ROWS, N_H1);
constexpr uint32_t B_MISC = B_BUFU1 + bytes_of(ROWS, N_M1);
constexpr uint32_t B_SMEM = B_MISC + 512 + 16;
static_assert(A_SMEM <= 232448, "tc_rows_kernel exceeds 227 KB of shared memory");
constexpr uint32_t J_TILE_BYTES = bytes_of(ROWS, K_J);

constexpr int kTmemCols = 256;

struct TailW;
struct TcState {
    uint8_t *img_a, *img_b;   // device weight images
    uint8_t *img_pair;        // two half images of the row network for tc_rows_pair_kernel (rank 0 | rank 1)
    long long *dbg;           // optional phase timestamps of CTA 0 (cn_debug_tc_timing)
    uint8_t *X;               // layer-1 operand tiles of the CTA-pair path (tc_features_kernel -> tc_rows_pair_kernel)
    size_t cap_xtiles;
    uint8_t *J;               // joint-state tiles
    double *rew;              // NG rewards
    size_t cap_groups;
    int num_sms;
    int variant;              // 0 = one tile in flight per SM (tc_rows_kernel), 2 = CTA pairs, two tiles in flight per SM
    float tail_a[104];        // attention.4 weight[100] + bias (kernel parameter of tc_rows_pair_kernel)
    float tail_b[104];        // mlp3.6 weight[100] + bias (kernel parameter of tc_mlp3_pair_kernel)
    uint8_t *img_pair_b;      // two half images of mlp3 for tc_mlp3_pair_kernel
};

__device__ __forceinline__ void copy_image_to_smem(uint8_t *dst, const uint8_t *__restrict__ src, uint32_t bytes)
{
    for (uint32_t i = threadIdx.x * 16; i < bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4 *>(dst + i) = __ldg(reinterpret_cast<const uint4 *>(src + i));
}

__device__ __forceinline__ uint32_t h2(float lo, float hi) { return pack_f16x2(lo, hi); }

// split x into fp16 hi + fp16 lo (x ~ hi + lo to ~22 bits)
__device__ __forceinline__ void split_hl(float x, float &hi, float &lo)
{
    hi = __half2float(__float2half_rn(x));
    lo = x - hi;
}

// =====================================================================================================
// kernel A: the per-(env, action, human) row network.  256 threads: warps w and w+4 own TMEM lanes
// 32*(w%4)..+31 (the hardware ties a warp to the lane quarter warpid%4) and split the columns of every
// epilogue between them.
// =====================================================================================================
constexpr int kThreadsTC = 256;      // worker threads (8 warps)
// A dedicated issuer warp was measured (tcgen05.mma issue blocks for about the MMA duration, ~62 cycles per
// M=128,K=16 instruction) but with one tile in flight it does not pay: 1.424 ms vs 1.372 ms per lookahead.
constexpr bool kIssuerWarp = false;
constexpr int kThreadsRows = kIssuerWarp ? 288 : 256;

// Inputs of one (env, action, human) row, fetched straight from the SoA (L2-resident) one tile ahead.
struct RowIn {
    double rpx, rpy, rgx, rgy, rr, rvp, hpx, hpy, hvx, hvy, hr, ax, ay;
    int valid;
};

__device__ __forceinline__ void load_row_inputs(RowIn &in, const EnvDims &ed, const double *__restrict__ st,
                                                const double *__restrict__ human_v, const double *__restrict__ actions,
                                                int A, int query_env, int NG, int G, int tile, int gl, int h)
{
    const int H = ed.H;
    const int g = tile * G + gl;
    in.valid = (gl < G && g < NG) ? 1 : 0;
    if (!in.valid) return;
    const int e = g / A, a = g - e * A;
    in.rpx = st[st_idx(ed, F_PX, 0, e)]; in.rpy = st[st_idx(ed, F_PY, 0, e)];
    in.rgx = st[st_idx(ed, F_GX, 0, e)]; in.rgy = st[st_idx(ed, F_GY, 0, e)];
    in.rr = st[st_idx(ed, F_R, 0, e)];   in.rvp = st[st_idx(ed, F_VPREF, 0, e)];
    in.hpx = st[st_idx(ed, F_PX, h + 1, e)]; in.hpy = st[st_idx(ed, F_PY, h + 1, e)];
    in.hr = st[st_idx(ed, F_R, h + 1, e)];
    if (query_env) {                                                                    // agent.py:63-74
        in.hvx = human_v[(size_t)(0 * H + h) * ed.E + e]; in.hvy = human_v[(size_t)(1 * H + h) * ed.E + e];
    } else {                                                                            // cadrl.py:107-109
        in.hvx = st[st_idx(ed, F_VX, h + 1, e)]; in.hvy = st[st_idx(ed, F_VY, h + 1, e)];
    }
    in.ax = actions[2 * a]; in.ay = actions[2 * a + 1];
}

// propagate + rotate (cadrl.py:104-129,217-252) -> the four 16-byte K-chunks of the X operand row:
// [x_hi(13) 1 1 0 | x_lo(13) 0 0 0]
__device__ __forceinline__ void row_features(const RowIn &in, double dt, uint4 &c0, uint4 &c1, uint4 &c2, uint4 &c3)
{
    c0 = make_uint4(0, 0, 0, 0); c1 = c0; c2 = c0; c3 = c0;
    if (!in.valid) return;
    float s[14], o[13];
    s[0] = (float)(in.rpx + in.ax * dt); s[1] = (float)(in.rpy + in.ay * dt);
    s[2] = (float)in.ax; s[3] = (float)in.ay; s[4] = (float)in.rr;
    s[5] = (float)in.rgx; s[6] = (float)in.rgy; s[7] = (float)in.rvp; s[8] = 0.0f;
    s[9] = (float)(in.hpx + in.hvx * dt); s[10] = (float)(in.hpy + in.hvy * dt);
    s[11] = (float)in.hvx; s[12] = (float)in.hvy; s[13] = (float)in.hr;
    cn_rotate(s, o);
    float hi[13], lo[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) split_hl(o[k], hi[k], lo[k]);
    c0 = make_uint4(h2(hi[0], hi[1]), h2(hi[2], hi[3]), h2(hi[4], hi[5]), h2(hi[6], hi[7]));
    c1 = make_uint4(h2(hi[8], hi[9]), h2(hi[10], hi[11]), h2(hi[12], 1.0f), h2(1.0f, 0.0f));
    c2 = make_uint4(h2(lo[0], lo[1]), h2(lo[2], lo[3]), h2(lo[4], lo[5]), h2(lo[6], lo[7]));
    c3 = make_uint4(h2(lo[8], lo[9]), h2(lo[10], lo[11]), h2(lo[12], 0.0f), 0u);
}

// Lookahead reward per (env, action) group, on the upper 128 threads (warps 4..7) of tc_rows_kernel: thread (gl, h)
// evaluates human h's clearance, a named barrier joins the 128 threads, then thread gl < G folds the H clearances
// through the reward ladder.  "Break on the first collision" (crowd_sim.py:360-363, multi_human_rl.py:71-73) only
// matters for dmin, which is unused once any clearance is negative, so min/any over all humans is the same result.
// The SoA loads (group_load) are issued one MMA wait earlier than the arithmetic (group_compute); the self-state
// chunks are written by the h == 0 ROW thread (row_features_j).
struct GrpIn {
    double rpx, rpy, rr, hpx, hpy, cvx, cvy, hr, ax, ay;   // role (gl, h): clearance of human h
    double gax, gay, gpx, gpy, grr, ggx, ggy, gt;          // role group t2 < G: reward ladder
    int valid, gvalid;
};

__device__ __forceinline__ void group_load(GrpIn &in, const EnvDims &ed, const double *__restrict__ st,
                                           const double *__restrict__ time, const double *__restrict__ actions, int A,
                                           int query_env, int NG, int G, int tile, int t2, int gl, int h)
{
    {
        const int g = tile * G + gl;
        in.valid = (gl < G && g < NG) ? 1 : 0;
        if (in.valid) {
            const int e = g / A, a = g - e * A;
            in.rpx = st[st_idx(ed, F_PX, 0, e)]; in.rpy = st[st_idx(ed, F_PY, 0, e)]; in.rr = st[st_idx(ed, F_R, 0, e)];
            in.hpx = st[st_idx(ed, F_PX, h + 1, e)]; in.hpy = st[st_idx(ed, F_PY, h + 1, e)];
            in.cvx = st[st_idx(ed, F_VX, h + 1, e)]; in.cvy = st[st_idx(ed, F_VY, h + 1, e)];
            in.hr = st[st_idx(ed, F_R, h + 1, e)];
            in.ax = actions[2 * a]; in.ay = actions[2 * a + 1];
        }
    }
    {
        const int g = tile * G + t2;
        in.gvalid = (t2 < G && g < NG) ? 1 : 0;
        if (in.gvalid) {
            const int e = g / A, a = g - e * A;
            in.gax = actions[2 * a]; in.gay = actions[2 * a + 1];
            in.gpx = st[st_idx(ed, F_PX, 0, e)]; in.gpy = st[st_idx(ed, F_PY, 0, e)]; in.grr = st[st_idx(ed, F_R, 0, e)];
            in.ggx = st[st_idx(ed, F_GX, 0, e)]; in.ggy = st[st_idx(ed, F_GY, 0, e)];
            in.gt = query_env ? time[e] : 0.0;
        }
    }
}

__device__ __forceinline__ void group_compute(const EnvParams &p, const GrpIn &in, int H, int query_env, int G, int tile,
                                              int t2, double *__restrict__ D, double *__restrict__ rew)
{
    const double dt = p.time_step;
    double clear = INFINITY;
    if (in.valid) {
        if (query_env) {   // crowd_sim.py:347-359
            const double px = in.hpx - in.rpx, py = in.hpy - in.rpy;
            const double vx = in.cvx - in.ax, vy = in.cvy - in.ay;
            const double ex = px + vx * dt, ey = py + vy * dt;
            clear = cn_point_to_segment_dist0(px, py, ex, ey) - in.hr - in.rr;
        } else {           // multi_human_rl.py:69-70
            const double npx = in.rpx + in.ax * dt, npy = in.rpy + in.ay * dt;
            const double nhx = in.hpx + in.cvx * dt, nhy = in.hpy + in.cvy * dt;
            clear = norm2d(npx - nhx, npy - nhy) - in.rr - in.hr;
        }
    }
    D[t2] = clear;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (in.gvalid) {
        double dmin = INFINITY;
        bool collision = false;
        for (int k = 0; k < H; ++k) {
            const double c = D[t2 * H + k];
            if (c < 0) collision = true;
            else if (c < dmin) dmin = c;
        }
        const double npx = in.gpx + in.gax * dt, npy = in.gpy + in.gay * dt;
        const bool reaching_goal = norm2d(npx - in.ggx, npy - in.ggy) < in.grr;
        double reward;
        if (query_env) {                                                                 // crowd_sim.py:382-403
            if (in.gt >= p.time_limit - 1) reward = 0;
            else if (collision) reward = p.collision_penalty;
            else if (reaching_goal) reward = p.success_reward;
            else if (dmin < p.discomfort_dist) reward = (dmin - p.discomfort_dist) * p.discomfort_penalty_factor * dt;
            else reward = 0;
        } else {                                                                         // multi_human_rl.py:77-86
            if (collision) reward = -0.25;
            else if (reaching_goal) reward = 1;
            else if (dmin < 0.2) reward = (dmin - 0.2) * 0.5 * dt;
            else reward = 0;
        }
        rew[tile * G + t2] = reward;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");   // D aliases the score scratch
}

// row_features plus, on the h == 0 row of a group, the self-state chunks 7..9 of the joint state (sarl.py:36)
__device__ __forceinline__ void row_features_j(const RowIn &in, double dt, int h, int g, uint8_t *__restrict__ J,
                                               uint4 &c0, uint4 &c1, uint4 &c2, uint4 &c3)
{
    row_features(in, dt, c0, c1, c2, c3);
    if (in.valid && h == 0) {
        // c0 = hi[0..7], c2 = lo[0..7]: the self state is columns 0..5 of the rotated row
        uint8_t *jt = J + (size_t)(g >> 7) * J_TILE_BYTES;
        const int rb = g & 127;
        *reinterpret_cast<uint4 *>(jt + chunk_off(ROWS, rb, 7)) = make_uint4(c0.x, c0.y, c0.z, h2(1.0f, 1.0f));
        *reinterpret_cast<uint4 *>(jt + chunk_off(ROWS, rb, 8)) = make_uint4(c2.x, c2.y, c2.z, 0u);
        *reinterpret_cast<uint4 *>(jt + chunk_off(ROWS, rb, 9)) = make_uint4(0, 0, 0, 0);
    }
}

__device__ __forceinline__ void pin(const uint4 &a, const uint4 &b, const uint4 &c, const uint4 &d)
{
    // keeps the prefetched feature words computed where they are written in the source (under the MMA wait)
    asm volatile("" :: "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w),
                       "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w), "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w));
}

__global__ void __launch_bounds__(kThreadsRows, 1)
tc_rows_kernel(EnvParams p, const double *__restrict__ st, const double *__restrict__ time,
               const double *__restrict__ human_v, const double *__restrict__ actions, int A, int query_env,
               const uint8_t *__restrict__ wimg, uint8_t *__restrict__ J, double *__restrict__ rew, int NG, int G,
               int ntiles, long long *__restrict__ dbg)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const EnvDims ed = p.d;
    const int H = ed.H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, hf = warp >> 2;   // hf: 0/1 = column half of the 8 worker warps, 2 = MMA issuer warp
    const bool worker = warp < 8;
    const bool issuer = kIssuerWarp ? ((warp == 8) && (lane == 0)) : (tid == 0);
    const int row = q * 32 + lane;            // TMEM lane == tile row owned by this thread
    uint8_t *bufA = smem + A_BUFA, *bufB = smem + A_BUFB;
    float *S0 = reinterpret_cast<float *>(smem + A_MISC);          // [128] partial scores, columns [0,64)
    float *S1 = S0 + 128;                                          // [128] partial scores, columns [64,100)
    const uint32_t mbar = smem_u32(smem + A_MISC + 1024);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + A_MISC + 1024 + 8);
    const float *tail = reinterpret_cast<const float *>(smem + OFF_TAILA);
    const double dt = p.time_step;

    // features / rewards of this CTA's first tile overlap the weight-image copy
    uint4 c0, c1, c2, c3;
    RowIn in;
    GrpIn gin;
    const int t2 = tid & 127;                 // row index (lower half) / (group, human) slot (upper half)
    const int my_gl = t2 / H, my_h = t2 - my_gl * H;
    double *Dscr = reinterpret_cast<double *>(smem + A_MISC);      // 128 clearances; aliases S0/S1
    if ((int)blockIdx.x < ntiles) {
        if (tid < 128) {
            load_row_inputs(in, ed, st, human_v, actions, A, query_env, NG, G, blockIdx.x, my_gl, my_h);
            row_features_j(in, dt, my_h, blockIdx.x * G + my_gl, J, c0, c1, c2, c3);
        } else if (worker) {
            group_load(gin, ed, st, time, actions, A, query_env, NG, G, blockIdx.x, t2, my_gl, my_h);
            group_compute(p, gin, H, query_env, G, blockIdx.x, t2, Dscr, rew);
        }
    }
    copy_image_to_smem(smem, wimg, IMG_A_BYTES);
    if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    constexpr uint32_t T_P = 448;
    // constant same-group matrix P[r][k] = (k / H == r / H), r, k < G*H, kept in TMEM for the whole kernel as the
    // A operand of the two group reductions (TS mode: column c of lane r holds k = 2c and 2c + 1)
    if (tid < 128) {
        const int lo = (tid / H) * H, hi = (tid < G * H) ? lo + H : lo;
        for (int c8 = 0; c8 < 8; ++c8) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int ka = (c8 * 8 + j) * 2, kb = ka + 1;
                w[j] = ((ka >= lo && ka < hi) ? 0x3C00u : 0u) | ((kb >= lo && kb < hi) ? 0x3C000000u : 0u);
            }
            st8(tlane + T_P + c8 * 8, w);
        }
        wait_st();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t sW1 = smem_u32(smem + OFF_W1), sW2 = smem_u32(smem + OFF_W2), sW3 = smem_u32(smem + OFF_W3);
    const uint32_t sW4 = smem_u32(smem + OFF_W4), sWA1 = smem_u32(smem + OFF_WA1), sWA2 = smem_u32(smem + OFF_WA2);
    const uint32_t sA = smem_u32(bufA), sB = smem_u32(bufB);
    uint32_t phase = 0;
    const int rows = G * H;
    constexpr int T_F = 0, T_A2 = N_F, T_D2 = N_F + N_M1;   // TMEM columns of the last two stages

#define TPROBE(i) do { if (dbg && blockIdx.x == 0 && tile == (int)(3 * gridDim.x)) { if (tid == 0) dbg[i] = clock64(); else if (tid == 255) dbg[32 + (i)] = clock64(); else if (tid == 64) dbg[64 + (i)] = clock64(); } } while (0)
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int g0 = tile * G;
        const int next = tile + gridDim.x;
        TPROBE(0);
        // ---- X operand of this tile (computed one tile ahead, lives in registers until here) ----
        if (tid < 128) {
            *reinterpret_cast<uint4 *>(bufB + chunk_off(ROWS, tid, 0)) = c0;
            *reinterpret_cast<uint4 *>(bufB + chunk_off(ROWS, tid, 1)) = c1;
            *reinterpret_cast<uint4 *>(bufB + chunk_off(ROWS, tid, 2)) = c2;
            *reinterpret_cast<uint4 *>(bufB + chunk_off(ROWS, tid, 3)) = c3;
        }
        fence_async_smem();
        __syncthreads();
        TPROBE(1);
        // ---- mlp1.0: X (K=32) -> TMEM[0,160) ----
        if (issuer) {
            fence_after_sync();
            mma_layer(tmem + 0, sB, ROWS, sW1, N_H1, K_X, N_H1, false);
            commit(mbar);
        }
        if (worker) mbar_wait(mbar, phase);
        phase ^= 1;
        TPROBE(2);
        fence_after_sync();
        if (hf == 0) epilogue_to_smem<true>(tlane, 0, 96, bufA, row, 0);       // H1
        else if (hf == 1) epilogue_to_smem<true>(tlane, 96, 64, bufA, row, 12);
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        TPROBE(3);
        // ---- mlp1.2: H1 (K=160) -> TMEM[0,112) ----
        if (issuer) {
            fence_after_sync();
            mma_layer(tmem + 0, sA, ROWS, sW2, N_M1, N_H1, N_M1, false);
            commit(mbar);
        }
        // next tile's row inputs: loads fly while the tensor cores work
        if (next < ntiles) {
            if (tid < 128) load_row_inputs(in, ed, st, human_v, actions, A, query_env, NG, G, next, my_gl, my_h);
            else if (worker) group_load(gin, ed, st, time, actions, A, query_env, NG, G, next, t2, my_gl, my_h);
        }
        if (worker) mbar_wait(mbar, phase);
        phase ^= 1;
        TPROBE(4);
        fence_after_sync();
        if (hf == 0) epilogue_to_smem<true>(tlane, 0, 64, bufB, row, 0);       // mlp1 output (X is dead)
        else if (hf == 1) epilogue_to_smem<true>(tlane, 64, 48, bufB, row, 8);
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        TPROBE(5);
        // ---- group sum of the mlp1 output over the humans of a group (sarl.py:42) as a UMMA: P (TMEM) x mlp1_out
        //      (read MN-major from bufB) -> TMEM[112,224); mlp2.0 -> TMEM[0,112) is queued right behind it and runs
        //      while the mean is converted ----
        if (issuer) {
            fence_after_sync();
            mma_layer_ts_bmn(tmem + N_M1, tmem + T_P, sB, ROWS, N_M1, false);
            commit(mbar);
            mma_layer(tmem + 0, sB, ROWS, sW3, N_M1, N_M1, N_M1, false);
        }
        TPROBE(14);
        if (worker) mbar_wait(mbar, phase);
        phase ^= 1;
        fence_after_sync();
        TPROBE(15);
        {
            const float inv = 1.0f / (float)H;                                  // mean = sum / H -> fp16 -> bufA
            if (hf == 0) epilogue_scaled_to_smem(tlane, N_M1, 64, inv, bufA, row, 0);
            else if (hf == 1) epilogue_scaled_to_smem(tlane, N_M1 + 64, 48, inv, bufA, row, 8);
        }
        TPROBE(16);
        fence_async_smem();
        fence_before_sync();
        TPROBE(17);
        __syncthreads();
        TPROBE(6);
        // ---- attention.0 on [mlp1_out | mean] (K = 224) -> TMEM[112,224) ----
        if (issuer) {
            fence_after_sync();
            mma_layer(tmem + N_M1, sB, ROWS, sWA1, N_M1, N_M1, N_M1, false);
            mma_layer(tmem + N_M1, sA, ROWS, sWA1 + (N_M1 / 8) * (N_M1 * 16), N_M1, N_M1, N_M1, true);
            commit(mbar);
        }
        // next tile: rotate + pack (rows) | rewards + self-state chunks (groups), hidden under the longest MMA
        if (next < ntiles) {
            if (tid < 128) { row_features_j(in, dt, my_h, next * G + my_gl, J, c0, c1, c2, c3); pin(c0, c1, c2, c3); }
            else if (worker) group_compute(p, gin, H, query_env, G, next, t2, Dscr, rew);
        }
        TPROBE(18);
        if (worker) mbar_wait(mbar, phase);
        phase ^= 1;
        TPROBE(7);
        fence_after_sync();
        if (hf == 0) epilogue_to_smem<true>(tlane, 0, N_M1, bufA, row, 0);     // mlp2.0 out
        else if (hf == 1) epilogue_to_smem<true>(tlane, N_M1, N_M1, bufB, row, 0);   // attention.0 out
        TPROBE(19);
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        TPROBE(8);
        // ---- mlp2.2 -> TMEM[0,64) ; attention.2 -> TMEM[64,176) ----
        if (issuer) {
            fence_after_sync();
            mma_layer(tmem + T_F, sA, ROWS, sW4, N_F, N_M1, N_F, false);
            mma_layer(tmem + T_A2, sB, ROWS, sWA2, N_M1, N_M1, N_M1, false);
            commit(mbar);
        }
        if (worker) mbar_wait(mbar, phase);
        phase ^= 1;
        TPROBE(9);
        fence_after_sync();
        // ---- attention.4 (fp32 dot over ReLU(attention.2)), split between the two warps of a lane quarter ----
        if (worker) {
            float part = 0.0f;
            if (hf == 0) {
                uint32_t v[32], u[32];
                ld32(tlane + T_A2, v);
                ld32(tlane + T_A2 + 32, u);
                wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) part = fmaf(fmaxf(__uint_as_float(v[k]), 0.0f), tail[k], part);
#pragma unroll
                for (int k = 0; k < 32; ++k) part = fmaf(fmaxf(__uint_as_float(u[k]), 0.0f), tail[32 + k], part);
                S0[row] = part;
            } else {
                uint32_t v[32], u[16];
                ld32(tlane + T_A2 + 64, v);
                ld16(tlane + T_A2 + 96, u);
                wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) part = fmaf(fmaxf(__uint_as_float(v[k]), 0.0f), tail[64 + k], part);
#pragma unroll
                for (int k = 0; k < 4; ++k) part = fmaf(fmaxf(__uint_as_float(u[k]), 0.0f), tail[96 + k], part);
                S1[row] = part;
            }
        }
        __syncthreads();
        TPROBE(10);
        // ---- masked un-stabilised softmax over the group (sarl.py:52-53); F' = w .* F as fp16 -> bufB ----
        if (worker) {
            float w = 0.0f;
            if (row < rows) {
                const int gl = row / H;
                float ssum = 0.0f, mine = 0.0f;
                for (int h = 0; h < H; ++h) {
                    const int r2 = gl * H + h;
                    const float sc = S0[r2] + S1[r2] + tail[100];
                    const float se = expf(sc) * (sc != 0.0f ? 1.0f : 0.0f);
                    ssum += se;
                    if (r2 == row) mine = se;
                }
                w = mine / ssum;
            }
            uint32_t v[32];
            ld32(tlane + T_F + hf * 32, v);
            wait_ld();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float *f = reinterpret_cast<const float *>(v) + c * 8;
                uint4 o;
                o.x = h2(w * f[0], w * f[1]); o.y = h2(w * f[2], w * f[3]);
                o.z = h2(w * f[4], w * f[5]); o.w = h2(w * f[6], w * f[7]);
                *reinterpret_cast<uint4 *>(bufB + chunk_off(ROWS, row, hf * 4 + c)) = o;
            }
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        TPROBE(11);
        if (issuer) {
            fence_after_sync();
            mma_layer_ts_bmn(tmem + T_D2, tmem + T_P, sB, ROWS, N_F, false);   // every row of a group gets the group sum
            commit(mbar);
        }
        if (worker) mbar_wait(mbar, phase);
        phase ^= 1;
        TPROBE(12);
        fence_after_sync();
        // ---- weighted feature of the group (replicated on its rows; the h == 0 row stores it) -> J chunks 0..6 ----
        if (tid < 128) {
            uint32_t v[32], u[32];
            ld32(tlane + T_D2, v);
            ld32(tlane + T_D2 + 32, u);
            wait_ld();
            const int g = g0 + my_gl;
            if (my_h == 0 && my_gl < G && g < NG) {
                uint8_t *jt = J + (size_t)(g >> 7) * J_TILE_BYTES;
                const int rb = g & 127;
#pragma unroll
                for (int c = 0; c < 4; ++c) cvt_store8<false>(v + 8 * c, jt + chunk_off(ROWS, rb, c));
#pragma unroll
                for (int c = 0; c < 3; ++c) cvt_store8<false>(u + 8 * c, jt + chunk_off(ROWS, rb, 4 + c));
            }
        }
        fence_before_sync();
        __syncthreads();   // bufA / bufB / TMEM are reused by the next tile
        TPROBE(13);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

#include "tc_rows_pair.cuh"
#include "tc_mlp3_pair.cuh"

// =====================================================================================================
// kernel B: mlp3 on the joint states + scoring (256 threads, same lane-quarter / column-half split)
// =====================================================================================================
__global__ void __launch_bounds__(kThreadsTC, 1)
tc_mlp3_kernel(EnvParams p, const double *__restrict__ st, const uint8_t *__restrict__ wimg,
               const uint8_t *__restrict__ J, const double *__restrict__ rew, int A, int NG, double gamma,
               double gamma_bar_host, double v_pref_host, double *__restrict__ values, int ntiles)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, hf = warp >> 2;
    const int row = q * 32 + lane;
    uint8_t *bufJ = smem + B_BUFJ, *bufU0 = smem + B_BUFU0, *bufU1 = smem + B_BUFU1;
    float *S1 = reinterpret_cast<float *>(smem + B_MISC);          // [128] partial of the upper column half
    const uint32_t mbar = smem_u32(smem + B_MISC + 512);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + B_MISC + 512 + 8);
    const float *tail = reinterpret_cast<const float *>(smem + OFF_TAILB);

    copy_image_to_smem(smem, wimg, IMG_B_BYTES);
    if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t sM1 = smem_u32(smem + OFF_M1), sM2 = smem_u32(smem + OFF_M2), sM3 = smem_u32(smem + OFF_M3);
    const uint32_t sJ = smem_u32(bufJ), sU0 = smem_u32(bufU0), sU1 = smem_u32(bufU1);
    uint32_t phase = 0;
    constexpr int kJVec = J_TILE_BYTES / 16 / kThreadsTC;   // 5 x 16 B per thread
    static_assert(J_TILE_BYTES == kJVec * 16 * kThreadsTC, "J tile must split evenly");

    uint4 pre[kJVec];
    if ((int)blockIdx.x < ntiles) {
        const uint8_t *src = J + (size_t)blockIdx.x * J_TILE_BYTES;
#pragma unroll
        for (int i = 0; i < kJVec; ++i) pre[i] = *reinterpret_cast<const uint4 *>(src + (size_t)(i * kThreadsTC + tid) * 16);
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll
        for (int i = 0; i < kJVec; ++i) *reinterpret_cast<uint4 *>(bufJ + (size_t)(i * kThreadsTC + tid) * 16) = pre[i];
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {   // mlp3.0
            fence_after_sync();
            mma_layer(tmem + 0, sJ, ROWS, sM1, N_H1, K_J, N_H1, false);
            commit(mbar);
        }
        if (tile + (int)gridDim.x < ntiles) {   // prefetch the next joint-state tile while the tensor cores work
            const uint8_t *src = J + (size_t)(tile + gridDim.x) * J_TILE_BYTES;
#pragma unroll
            for (int i = 0; i < kJVec; ++i) pre[i] = *reinterpret_cast<const uint4 *>(src + (size_t)(i * kThreadsTC + tid) * 16);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        fence_after_sync();
        if (hf == 0) epilogue_to_smem<true>(tlane, 0, 96, bufU0, row, 0);
        else epilogue_to_smem<true>(tlane, 96, 64, bufU0, row, 12);
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        if (tid == 0) {   // mlp3.2
            fence_after_sync();
            mma_layer(tmem + 0, sU0, ROWS, sM2, N_M1, N_H1, N_M1, false);
            commit(mbar);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        fence_after_sync();
        if (hf == 0) epilogue_to_smem<true>(tlane, 0, 64, bufU1, row, 0);
        else epilogue_to_smem<true>(tlane, 64, 48, bufU1, row, 8);
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        if (tid == 0) {   // mlp3.4
            fence_after_sync();
            mma_layer(tmem + 0, sU1, ROWS, sM3, N_M1, N_M1, N_M1, false);
            commit(mbar);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        fence_after_sync();
        // mlp3.6 as an fp32 dot over ReLU(mlp3.4), then value = reward + gamma_bar * V (multi_human_rl.py:52)
        float part = 0.0f;
        if (hf == 0) {
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t x[32];
                ld32(tlane + c0, x);
                wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) part = fmaf(fmaxf(__uint_as_float(x[k]), 0.0f), tail[c0 + k], part);
            }
        } else {
            uint32_t x[32];
            ld32(tlane + 64, x);
            wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) part = fmaf(fmaxf(__uint_as_float(x[k]), 0.0f), tail[64 + k], part);
            uint32_t u[16];
            ld16(tlane + 96, u);
            wait_ld();
#pragma unroll
            for (int k = 0; k < 4; ++k) part = fmaf(fmaxf(__uint_as_float(u[k]), 0.0f), tail[96 + k], part);
            S1[row] = part;
        }
        fence_before_sync();
        __syncthreads();
        if (hf == 0) {
            const float v = part + S1[row] + tail[100];
            const int g = tile * ROWS + row;
            if (g < NG) {
                const int e = g / A;
                const double vp = st[st_idx(p.d, F_VPREF, 0, e)];
                const double gamma_bar = (vp == v_pref_host) ? gamma_bar_host : pow(gamma, p.time_step * vp);
                values[g] = rew[g] + gamma_bar * (double)v;
            }
        }
        // S1 is rewritten only after the next tile's three barriers; bufJ after the next loop-top barrier
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// =====================================================================================================
// self-test kernel: D[128 x N] = A[128 x K] * B[N x K]^T
// =====================================================================================================
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const uint8_t *__restrict__ a_img, const uint8_t *__restrict__ b_img, float *__restrict__ d,
                     int N, int K, int mode /*0 SS, 1 SS with MN-major B, 2 TS (A in TMEM)*/, int reps,
                     long long *__restrict__ cycles, int a_col, int d_col)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t a_bytes = bytes_of(ROWS, K), b_bytes = mode == 1 ? bytes_of(ROWS, N) : bytes_of(N, K);
    uint8_t *sa = smem, *sb = smem + a_bytes;
    uint8_t *misc = smem + ((a_bytes + b_bytes + 15) & ~15u);
    const uint32_t mbar = smem_u32(misc);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(misc + 8);
    copy_image_to_smem(sa, a_img, a_bytes);
    copy_image_to_smem(sb, b_img, b_bytes);
    if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    // TS variants: mode 2 -> A at column 256, 3 -> 128, 4 -> right after D (column N), 5 -> 448
    const uint32_t T_A = mode == 7 ? (uint32_t)a_col : (mode == 3 ? 128 : (mode == 4 ? (uint32_t)N : (mode == 5 ? 448 : 256)));
    const uint32_t T_D = mode == 7 ? (uint32_t)d_col : 0;
    if (mode >= 2 && mode != 6) {
        // row `tid` of A: K/2 packed words, 8 per tcgen05.st
        for (int c8 = 0; c8 < K / 16; ++c8) {
            uint32_t w[8];
            for (int j = 0; j < 8; ++j) {
                const int k = c8 * 16 + 2 * j;   // element k lives in chunk k/8 at position k%8
                const __half2 h = *reinterpret_cast<const __half2 *>(sa + chunk_off(ROWS, tid, k >> 3) + (k & 7) * 2);
                w[j] = *reinterpret_cast<const uint32_t *>(&h);
            }
            st8(tlane + T_A + c8 * 8, w);
        }
        wait_st();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
    }
    uint32_t ph = 0;
    long long t0 = 0, t1 = 0;
    if (mode == 6) {
        // reproduce the ping-pong pattern: accumulator [0,K) written by an MMA, compacted in place to fp16 [0,K/2)
        // by tcgen05.ld/st, then used as the TS A operand of a second product whose D sits right after it.
        if (tid == 0) {
            mma_layer(tmem, smem_u32(sa), ROWS, smem_u32(sb), N, K, N, false);   // needs N >= K: D[0,N) covers [0,K)
            commit(mbar);
        }
        mbar_wait(mbar, ph); ph ^= 1;
        fence_after_sync();
        compact_to_tmem<false>(tlane, 0, K, 0, 1.0f);
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        if (tid == 0) t0 = clock64();
        for (int rep = 0; rep < reps; ++rep) {
            if (tid == 0) {
                mma_layer_ts(tmem + d_col, tmem, smem_u32(sb), N, K, N, false);
                commit(mbar);
            }
            mbar_wait(mbar, ph); ph ^= 1;
        }
        if (tid == 0) { t1 = clock64(); if (cycles) *cycles = t1 - t0; }
        fence_after_sync();
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            ld16(tlane + d_col + c0, v);
            wait_ld();
            for (int k = 0; k < 16; ++k) d[(size_t)tid * N + c0 + k] = __uint_as_float(v[k]);
        }
        fence_before_sync();
        __syncthreads();
        if (warp == 0) tmem_dealloc(tmem, 512);
        return;
    }
    if (tid == 0) t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        if (tid == 0) {
            if (mode == 1) mma_layer_bmn(tmem, smem_u32(sa), ROWS, smem_u32(sb), K, N, false);
            else if (mode >= 2) mma_layer_ts(tmem + T_D, tmem + T_A, smem_u32(sb), N, K, N, false);
            else mma_layer(tmem, smem_u32(sa), ROWS, smem_u32(sb), N, K, N, false);
            commit(mbar);
        }
        mbar_wait(mbar, ph); ph ^= 1;
    }
    if (tid == 0) { t1 = clock64(); if (cycles) *cycles = t1 - t0; }
    fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        ld16(tlane + T_D + c0, v);
        wait_ld();
        for (int k = 0; k < 16; ++k) d[(size_t)tid * N + c0 + k] = __uint_as_float(v[k]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- host-side weight images ------------------------------------------------------------------------
struct HostLinear { const float *w, *b; int in, out; };

inline void put(std::vector<uint8_t> &img, uint32_t base, int R, int r, int k, float v)
{
    const __half h = __float2half_rn(v);
    memcpy(&img[base + chunk_off(R, r, k >> 3) + (k & 7) * 2], &h, 2);
}
inline float f16r(float v) { return __half2float(__float2half_rn(v)); }

// generic layer: rows n < out get W[n][k] at K index kmap(k), bias hi/lo at kb/kb+1; ones columns at
// output rows (ones_n, ones_n+1) fed from K index kb (which holds 1.0 in the activation tile)
void fill_layer(std::vector<uint8_t> &img, uint32_t base, int R, const HostLinear &L, int k_in0, int n_in, int k_dst0,
                int kb, int ones_n)
{
    for (int n = 0; n < L.out; ++n) {
        for (int k = 0; k < n_in; ++k) put(img, base, R, n, k_dst0 + k, L.w[(size_t)n * L.in + k_in0 + k]);
        if (kb >= 0) {
            const float bh = f16r(L.b[n]);
            put(img, base, R, n, kb, bh);
            put(img, base, R, n, kb + 1, L.b[n] - bh);
        }
    }
    if (ones_n >= 0 && kb >= 0) { put(img, base, R, ones_n, kb, 1.0f); put(img, base, R, ones_n + 1, kb, 1.0f); }
}

// rows [rank N/2, (rank+1) N/2) of a chunked K-major [N x K] image -> chunked K-major [N/2 x K]
void split_rows(uint8_t *dst, const uint8_t *src, int N, int K, int rank)
{
    const int hn = N / 2;
    for (int c = 0; c < K / 8; ++c)
        memcpy(dst + (size_t)c * hn * 16, src + (size_t)c * N * 16 + (size_t)rank * hn * 16, (size_t)hn * 16);
}

}  // namespace

int cn_tc_init(cn_policy *p)
{
    const SarlDims &d = p->d;
    const bool ok = d.in == 13 && d.self_dim == 6 && d.m1[0] == 150 && d.m1[1] == 100 && d.m2[0] == 100 && d.m2[1] == 50 &&
                    d.at[0] == 100 && d.at[1] == 100 && d.at[2] == 1 && d.m3[0] == 150 && d.m3[1] == 100 &&
                    d.m3[2] == 100 && d.m3[3] == 1;
    if (!ok) {
        cn_set_error("CN_PREC_F16_TC is specialised for the default [sarl] dims (150,100 / 100,50 / 100,100,1 / "
                     "150,100,100,1); use CN_PREC_F32 for other shapes");
        return CN_EUNSUPPORTED;
    }
    TcState *t = new TcState();
    memset(t, 0, sizeof(*t));
    cudaDeviceProp prop;
    CN_CUDA_CHECK(cudaGetDeviceProperties(&prop, p->device));
    t->num_sms = prop.multiProcessorCount;
    if (cudaMalloc((void **)&t->img_a, IMG_A_BYTES) != cudaSuccess || cudaMalloc((void **)&t->img_b, IMG_B_BYTES) != cudaSuccess) {
        cn_set_error("cudaMalloc failed for the tensor-core weight images");
        delete t;
        return CN_ENOMEM;
    }
    CN_CUDA_CHECK(cudaFuncSetAttribute(tc_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A_SMEM));
    CN_CUDA_CHECK(cudaFuncSetAttribute(tc_mlp3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B_SMEM));
    CN_CUDA_CHECK(cudaFuncSetAttribute(tc_rows_pair_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q_SMEM));
    CN_CUDA_CHECK(cudaFuncSetAttribute(tc_rows_pair_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q_SMEM));
    CN_CUDA_CHECK(cudaFuncSetAttribute(tc_rows_pair_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q_SMEM));
    CN_CUDA_CHECK(cudaFuncSetAttribute(tc_mlp3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M_SMEM));
    if (cudaMalloc((void **)&t->img_pair, 2 * IMG_H_BYTES) != cudaSuccess ||
        cudaMalloc((void **)&t->img_pair_b, 2 * IMG_HM_BYTES) != cudaSuccess) {
        cn_set_error("cudaMalloc failed for the tensor-core weight images");
        return CN_ENOMEM;
    }
    const char *var = getenv("CN_TC_VARIANT");   // developer switch: "single" = one tile in flight per SM
    t->variant = (var && strcmp(var, "single") == 0) ? 0 : 2;
    p->tc = t;
    return CN_OK;
}

void cn_tc_destroy(cn_policy *p)
{
    TcState *t = (TcState *)p->tc;
    if (!t) return;
    if (t->img_a) cudaFree(t->img_a);
    if (t->img_b) cudaFree(t->img_b);
    if (t->img_pair) cudaFree(t->img_pair);
    if (t->img_pair_b) cudaFree(t->img_pair_b);
    if (t->J) cudaFree(t->J);
    if (t->X) cudaFree(t->X);
    if (t->rew) cudaFree(t->rew);
    if (t->dbg) cudaFree(t->dbg);
    delete t;
    p->tc = nullptr;
}

int cn_tc_load_weights(cn_policy *p, const float *flat, cudaStream_t s)
{
    TcState *t = (TcState *)p->tc;
    const SarlDims &d = p->d;
    HostLinear L[11];
    const int ins[11] = {d.in, d.m1[0], d.m1[1], d.m2[0], 2 * d.m1[1], d.at[0], d.at[1], d.m2[1] + d.self_dim, d.m3[0], d.m3[1], d.m3[2]};
    const int outs[11] = {d.m1[0], d.m1[1], d.m2[0], d.m2[1], d.at[0], d.at[1], d.at[2], d.m3[0], d.m3[1], d.m3[2], d.m3[3]};
    const float *src = flat;
    for (int i = 0; i < 11; ++i) {
        L[i].w = src; src += (size_t)ins[i] * outs[i];
        L[i].b = src; src += outs[i];
        L[i].in = ins[i]; L[i].out = outs[i];
    }
    std::vector<uint8_t> a(IMG_A_BYTES, 0), b(IMG_B_BYTES, 0);
    // mlp1.0: K = [x_hi(13) 1 1 0 | x_lo(13) 0 0 0]; bias at k = 13,14; ones -> n = 150,151
    fill_layer(a, OFF_W1, N_H1, L[0], 0, 13, 0, 13, 150);
    fill_layer(a, OFF_W1, N_H1, L[0], 0, 13, 16, -1, -1);
    fill_layer(a, OFF_W2, N_M1, L[1], 0, 150, 0, 150, 100);          // mlp1.2
    fill_layer(a, OFF_W3, N_M1, L[2], 0, 100, 0, 100, 100);          // mlp2.0
    fill_layer(a, OFF_W4, N_F, L[3], 0, 100, 0, 100, -1);            // mlp2.2 (no ones needed downstream)
    fill_layer(a, OFF_WA1, N_M1, L[4], 0, 100, 0, 100, 100);         // attention.0, mlp1_out half
    fill_layer(a, OFF_WA1, N_M1, L[4], 100, 100, N_M1, -1, -1);      //              group-mean half
    fill_layer(a, OFF_WA2, N_M1, L[5], 0, 100, 0, 100, -1);          // attention.2
    float tail[TAIL_BYTES / 4] = {0};
    for (int k = 0; k < 100; ++k) tail[k] = L[6].w[k];
    tail[100] = L[6].b[0];
    memcpy(&a[OFF_TAILA], tail, TAIL_BYTES);
    memcpy(t->tail_a, tail, sizeof(t->tail_a));
    // mlp3.0 on J = [weighted(50) pad6 | self_hi(6) 1 1 | self_lo(6) 0 0 | pad8]; joint = [self(6), weighted(50)]
    fill_layer(b, OFF_M1, N_H1, L[7], 6, 50, 0, -1, -1);
    fill_layer(b, OFF_M1, N_H1, L[7], 0, 6, 56, 62, 150);
    fill_layer(b, OFF_M1, N_H1, L[7], 0, 6, 64, -1, -1);
    fill_layer(b, OFF_M2, N_M1, L[8], 0, 150, 0, 150, 100);          // mlp3.2
    fill_layer(b, OFF_M3, N_M1, L[9], 0, 100, 0, 100, -1);           // mlp3.4
    for (int k = 0; k < 100; ++k) tail[k] = L[10].w[k];
    tail[100] = L[10].b[0];
    memcpy(&b[OFF_TAILB], tail, TAIL_BYTES);
    memcpy(t->tail_b, tail, sizeof(t->tail_b));
    std::vector<uint8_t> hb(2 * IMG_HM_BYTES, 0);
    for (int rank = 0; rank < 2; ++rank) {
        uint8_t *h = hb.data() + (size_t)rank * IMG_HM_BYTES;
        split_rows(h + HM_M1, b.data() + OFF_M1, N_H1, K_J, rank);
        split_rows(h + HM_M2, b.data() + OFF_M2, N_M1, N_H1, rank);
        split_rows(h + HM_M3, b.data() + OFF_M3, N_M1, N_M1, rank);
    }
    CN_CUDA_CHECK(cudaMemcpyAsync(t->img_pair_b, hb.data(), 2 * IMG_HM_BYTES, cudaMemcpyHostToDevice, s));
    // half images for the CTA-pair kernel: CTA `rank` holds output rows [rank N/2, (rank+1) N/2) of every layer
    std::vector<uint8_t> hp(2 * IMG_H_BYTES, 0);
    for (int rank = 0; rank < 2; ++rank) {
        uint8_t *h = hp.data() + (size_t)rank * IMG_H_BYTES;
        split_rows(h + H_W1, a.data() + OFF_W1, N_H1, K_X, rank);
        split_rows(h + H_W2, a.data() + OFF_W2, N_M1, N_H1, rank);
        // stage 2, N = 224: rank 0 = mlp2.0 (all 112 rows), rank 1 = attention.0 rows on the mlp1_out half of K
        // stage 2 is ONE N = 224 UMMA: rank 0's half = attention.0 (mlp1_out half of K) -> D [0,112), rank 1's = mlp2.0 -> [112,224)
        memcpy(h + H_W3A, rank == 0 ? a.data() + OFF_WA1 : a.data() + OFF_W3, bytes_of(N_M1, N_M1));
        split_rows(h + H_WB, a.data() + OFF_WA1 + bytes_of(N_M1, N_M1), N_M1, N_M1, rank);   // group-mean half of K
        split_rows(h + H_W4, a.data() + OFF_W4, N_F, N_M1, rank);
        split_rows(h + H_WA2, a.data() + OFF_WA2, N_M1, N_M1, rank);
        memcpy(h + H_TAIL, a.data() + OFF_TAILA, TAIL_BYTES);
    }
    CN_CUDA_CHECK(cudaMemcpyAsync(t->img_pair, hp.data(), 2 * IMG_H_BYTES, cudaMemcpyHostToDevice, s));
    CN_CUDA_CHECK(cudaMemcpyAsync(t->img_a, a.data(), IMG_A_BYTES, cudaMemcpyHostToDevice, s));
    CN_CUDA_CHECK(cudaMemcpyAsync(t->img_b, b.data(), IMG_B_BYTES, cudaMemcpyHostToDevice, s));
    CN_CUDA_CHECK(cudaStreamSynchronize(s));
    return CN_OK;
}

int cn_lookahead_tc(cn_policy *p, cn_env *env, int query_env, double epsilon, cudaStream_t s, cudaStream_t tail)
{
    TcState *t = (TcState *)p->tc;
    const EnvDims ed = env->p.d;
    if (ed.H > CN_MAX_HUMANS) { cn_set_error("human_num too large"); return CN_EUNSUPPORTED; }
    int rc = cn_lookahead_prepare(p, env, s);
    if (rc) return rc;
    const int A = p->d.A;
    const size_t NG = (size_t)ed.E * A;
    if (NG > t->cap_groups) {
        if (t->J) cudaFree(t->J);
        if (t->rew) cudaFree(t->rew);
        t->J = nullptr; t->rew = nullptr; t->cap_groups = 0;
        // the CTA-pair mlp3 kernel reads whole rounds of tiles: pad with up to one round of (zero) tiles
        const size_t jtiles = (NG + ROWS - 1) / ROWS + 4 * (size_t)(t->num_sms / 2);
        CN_CUDA_CHECK(cudaMalloc((void **)&t->J, jtiles * J_TILE_BYTES));
        CN_CUDA_CHECK(cudaMalloc((void **)&t->rew, sizeof(double) * NG));
        CN_CUDA_CHECK(cudaMemsetAsync(t->J, 0, jtiles * J_TILE_BYTES, s));
        t->cap_groups = NG;
    }
    int G = ROWS / ed.H;
    if (G > 64) G = 64;                      // keeps the staged environments of a tile within bufA
    const int ntiles_a = (int)((NG + G - 1) / G);
    const int ntiles_b = (int)((NG + ROWS - 1) / ROWS);
    const double gamma_bar = pow(p->cfg.gamma, env->p.time_step * p->cfg.v_pref);
    const int grid_a = ntiles_a < t->num_sms ? ntiles_a : t->num_sms;
    const int grid_b = ntiles_b < t->num_sms ? ntiles_b : t->num_sms;
    if (t->variant == 2) {
        // CTA pairs: every (cluster, rank, context) slot runs the same number of rounds; tiles past the end are all padding
        int nclusters = t->num_sms / 2;
        const int slots_needed = (ntiles_a + 3) / 4;
        if (nclusters > slots_needed) nclusters = slots_needed;
        const int rounds = (ntiles_a + 4 * nclusters - 1) / (4 * nclusters);
        const size_t xtiles = (size_t)rounds * 4 * nclusters;
        if (xtiles > t->cap_xtiles) {
            if (t->X) cudaFree(t->X);
            t->X = nullptr; t->cap_xtiles = 0;
            CN_CUDA_CHECK(cudaMalloc((void **)&t->X, xtiles * X_TILE_BYTES));
            t->cap_xtiles = xtiles;
        }
        TailW tw;
        memcpy(tw.w, t->tail_a, sizeof(tw.w));
        const int ht = (ed.H == 5 && G == ROWS / 5) ? 5 : ((ed.H == 10 && G == ROWS / 10) ? 10 : 0);
        auto feat = ht == 5 ? tc_features_kernel<5> : (ht == 10 ? tc_features_kernel<10> : tc_features_kernel<0>);
        auto kern = ht == 5 ? tc_rows_pair_kernel<5> : (ht == 10 ? tc_rows_pair_kernel<10> : tc_rows_pair_kernel<0>);
        cn_trace_mark("features", s);
        feat<<<(unsigned)xtiles, ROWS, 0, s>>>(env->p, env->state, env->time, env->human_v, p->action_dev, A, query_env, (int)NG, G,
                                              env->theta, t->X, t->J, t->rew);
        CN_LAUNCH_CHECK();
        cn_trace_mark("rows", s);
        kern<<<2 * nclusters, kThreadsPair, Q_SMEM, s>>>(env->p, t->img_pair, t->X, t->J, (int)NG, G, rounds, tw, t->dbg);
    } else {
        if (env->p.kinematics != CN_KIN_HOLONOMIC) {
            cn_set_error("CN_TC_VARIANT=single supports holonomic kinematics only");
            return CN_EUNSUPPORTED;
        }
        tc_rows_kernel<<<grid_a, kThreadsRows, A_SMEM, s>>>(env->p, env->state, env->time, env->human_v, p->action_dev, A,
                                                            query_env, t->img_a, t->J, t->rew, (int)NG, G, ntiles_a, t->dbg);
    }
    CN_LAUNCH_CHECK();
    if (tail && tail != s) {
        // pipelined host steps: the rest of this shard's step runs on a HIGH-priority stream, so that when this row kernel
        // exits its mlp3 is placed before the other shard's (already queued, lower-priority) row kernel takes every SM
        CN_CUDA_CHECK(cudaEventRecord(env->ev_rows, s));
        CN_CUDA_CHECK(cudaStreamWaitEvent(tail, env->ev_rows, 0));
        s = tail;
    }
    if (t->variant == 2) {
        int nclusters = t->num_sms / 2;
        const int slots_needed = (ntiles_b + 3) / 4;
        if (nclusters > slots_needed) nclusters = slots_needed;
        const int rounds = (ntiles_b + 4 * nclusters - 1) / (4 * nclusters);
        TailW tw;
        memcpy(tw.w, t->tail_b, sizeof(tw.w));
        cn_trace_mark("mlp3", s);
        tc_mlp3_pair_kernel<<<2 * nclusters, kThreadsM3, M_SMEM, s>>>(env->p, env->state, t->img_pair_b, t->J, t->rew, A, (int)NG,
                                                                     p->cfg.gamma, gamma_bar, p->cfg.v_pref, p->values, rounds, tw);
    } else {
        tc_mlp3_kernel<<<grid_b, kThreadsTC, B_SMEM, s>>>(env->p, env->state, t->img_b, t->J, t->rew, A, (int)NG, p->cfg.gamma,
                                                   gamma_bar, p->cfg.v_pref, p->values, ntiles_b);
    }
    CN_LAUNCH_CHECK();
    cn_trace_mark("argmax", s);
    return cn_lookahead_argmax(p, env, epsilon, s);
}

static int selftest_impl(int32_t N, int32_t K, const float *a_host, const float *b_host, float *d_host, int device, int mode,
                         int reps, long long *cycles_host, int a_col = 256, int d_col = 256)
{
    const int bmn = mode == 1;
    if (N < 16 || N > 256 || N % 16 || K < 16 || K % 16 || K > 256 || (bmn && K > ROWS)) {
        cn_set_error("N in [16,256] step 16, K in [16,256] step 16 (K <= 128 for the MN-major variant)");
        return CN_EINVAL;
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); cn_set_error("no CUDA device available; no CPU fallback"); return CN_ECUDA; }
    CN_CUDA_CHECK(cudaSetDevice(device));
    std::vector<uint8_t> ai(bytes_of(ROWS, K), 0), bi(bmn ? bytes_of(ROWS, N) : bytes_of(N, K), 0);
    for (int r = 0; r < ROWS; ++r) for (int k = 0; k < K; ++k) put(ai, 0, ROWS, r, k, a_host[(size_t)r * K + k]);
    for (int r = 0; r < N; ++r)
        for (int k = 0; k < K; ++k) {
            if (bmn) put(bi, 0, ROWS, k, r, b_host[(size_t)r * K + k]);   // activation-style image: rows = k, columns = n
            else put(bi, 0, N, r, k, b_host[(size_t)r * K + k]);
        }
    uint8_t *da = nullptr, *db = nullptr;
    float *dd = nullptr;
    long long *dc = nullptr;
    CN_CUDA_CHECK(cudaMalloc((void **)&da, ai.size()));
    CN_CUDA_CHECK(cudaMalloc((void **)&db, bi.size()));
    CN_CUDA_CHECK(cudaMalloc((void **)&dd, sizeof(float) * ROWS * N));
    CN_CUDA_CHECK(cudaMalloc((void **)&dc, sizeof(long long)));
    CN_CUDA_CHECK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
    CN_CUDA_CHECK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
    const size_t smem = ai.size() + bi.size() + 64;
    CN_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem>>>(da, db, dd, N, K, mode, reps, dc, a_col, d_col);
    CN_LAUNCH_CHECK();
    CN_CUDA_CHECK(cudaDeviceSynchronize());
    CN_CUDA_CHECK(cudaMemcpy(d_host, dd, sizeof(float) * ROWS * N, cudaMemcpyDeviceToHost));
    if (cycles_host) CN_CUDA_CHECK(cudaMemcpy(cycles_host, dc, sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dc);
    return CN_OK;
}

extern "C" int cn_selftest_umma(int32_t N, int32_t K, const float *a_host, const float *b_host, float *d_host, int device)
{
    return selftest_impl(N, K, a_host, b_host, d_host, device, 0, 1, nullptr);
}

extern "C" int cn_selftest_umma_bmn(int32_t N, int32_t K, const float *a_host, const float *b_host, float *d_host, int device)
{
    return selftest_impl(N, K, a_host, b_host, d_host, device, 1, 1, nullptr);
}

// CTA-pair building block: D[256 x N] = A[256 x K] * B[N x K]^T with one cta_group::2 UMMA chain (N % 16 == 0)
extern "C" int cn_selftest_umma_pair(int32_t N, int32_t K, const float *a_host, const float *b_host, float *d_host,
                                     int32_t reps, long long *cycles_host, int device)
{
    if (N < 32 || N > 256 || N % 16 || K < 16 || K % 16 || K > 256 || reps < 1) {
        cn_set_error("N in [32,256] step 16, K in [16,256] step 16, reps >= 1");
        return CN_EINVAL;
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); cn_set_error("no CUDA device available; no CPU fallback"); return CN_ECUDA; }
    CN_CUDA_CHECK(cudaSetDevice(device));
    const size_t a_bytes = bytes_of(ROWS, K), b_bytes = bytes_of(N / 2, K);
    std::vector<uint8_t> ai(2 * a_bytes, 0), bi(2 * b_bytes, 0);
    for (int r = 0; r < 2 * ROWS; ++r)
        for (int k = 0; k < K; ++k) put(ai, (uint32_t)((r / ROWS) * a_bytes), ROWS, r % ROWS, k, a_host[(size_t)r * K + k]);
    for (int r = 0; r < N; ++r)
        for (int k = 0; k < K; ++k) put(bi, (uint32_t)((r / (N / 2)) * b_bytes), N / 2, r % (N / 2), k, b_host[(size_t)r * K + k]);
    uint8_t *da = nullptr, *db = nullptr;
    float *dd = nullptr;
    long long *dc = nullptr;
    CN_CUDA_CHECK(cudaMalloc((void **)&da, ai.size()));
    CN_CUDA_CHECK(cudaMalloc((void **)&db, bi.size()));
    CN_CUDA_CHECK(cudaMalloc((void **)&dd, sizeof(float) * 2 * ROWS * N));
    CN_CUDA_CHECK(cudaMalloc((void **)&dc, sizeof(long long)));
    CN_CUDA_CHECK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
    CN_CUDA_CHECK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
    const size_t smem = a_bytes + b_bytes + 64;
    CN_CUDA_CHECK(cudaFuncSetAttribute(umma_pair_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_pair_selftest_kernel<<<2, 160, smem>>>(da, db, dd, N, K, reps, dc);
    CN_LAUNCH_CHECK();
    CN_CUDA_CHECK(cudaDeviceSynchronize());
    CN_CUDA_CHECK(cudaMemcpy(d_host, dd, sizeof(float) * 2 * ROWS * N, cudaMemcpyDeviceToHost));
    if (cycles_host) CN_CUDA_CHECK(cudaMemcpy(cycles_host, dc, sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dc);
    return CN_OK;
}

// Developer diagnostic (not part of the reference surface): clock64() at the phase boundaries of one tile of
// tc_rows_kernel's CTA 0.  First call arms the probes; later calls return the last recorded timestamps.
extern "C" int cn_debug_tc_timing(cn_policy *p, long long *out16)
{
    if (!p || !p->tc) { cn_set_error("policy has no tensor-core state"); return CN_EINVAL; }
    TcState *t = (TcState *)p->tc;
    CN_CUDA_CHECK(cudaSetDevice(p->device));
    if (!t->dbg) {
        CN_CUDA_CHECK(cudaMalloc((void **)&t->dbg, 256 * sizeof(long long)));
        CN_CUDA_CHECK(cudaMemset(t->dbg, 0, 256 * sizeof(long long)));
    }
    CN_CUDA_CHECK(cudaDeviceSynchronize());
    CN_CUDA_CHECK(cudaMemcpy(out16, t->dbg, 256 * sizeof(long long), cudaMemcpyDeviceToHost));
    return CN_OK;
}

// Developer diagnostic: tcgen05.ld throughput.  `nwarps` warps (4 per TMEM lane quarter at most useful) each run
// `iters` iterations of {loads of `cols` fp32 columns, wait::ld}; mode 0 = 32x32b.x32 loads, 1 = .x64, 2 = .x128,
// 3 = x32 pairs followed by the epilogue's cvt + st.shared work.  *cycles = clock64() ticks of the slowest warp.
namespace {
__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t (&r)[64])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
          "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(512, 1)
tmem_ld_bench_kernel(int mode, int iters, long long *__restrict__ cycles, uint32_t *__restrict__ sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tl = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 128;
    uint32_t acc = 0;
    const int row = (warp & 3) * 32 + lane;
    uint8_t *dst = smem + (size_t)(warp >> 2) * 32768;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
            uint32_t v[32], u[32];
            ld32(tl, v); ld32(tl + 32, u);
            wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) acc ^= v[k] ^ u[k];
        } else if (mode == 1) {
            uint32_t v[64];
            ld64(tl, v);
            wait_ld();
#pragma unroll
            for (int k = 0; k < 64; ++k) acc ^= v[k];
        } else if (mode == 3) {
            uint32_t v[32], u[32];
            ld32(tl, v); ld32(tl + 32, u);
            wait_ld();
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) cvt_store8<true>(v + qd * 8, dst + chunk_off(128, row, qd));
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) cvt_store8<true>(u + qd * 8, dst + chunk_off(128, row, 4 + qd));
        } else {
            uint32_t v[16];
            ld16(tl, v);
            wait_ld();
#pragma unroll
            for (int k = 0; k < 16; ++k) acc ^= v[k];
        }
    }
    const long long t1 = clock64();
    if (lane == 0) atomicMax((unsigned long long *)cycles, (unsigned long long)(t1 - t0));
    if (acc == 0x12345678u) sink[tid] = acc;
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, 512);
}
}  // namespace

extern "C" int cn_debug_tmem_bench(int32_t mode, int32_t nwarps, int32_t iters, long long *cycles_host, int device)
{
    if (nwarps < 1 || nwarps > 16 || iters < 1) { cn_set_error("nwarps in [1,16], iters >= 1"); return CN_EINVAL; }
    CN_CUDA_CHECK(cudaSetDevice(device));
    long long *dc = nullptr;
    uint32_t *sink = nullptr;
    CN_CUDA_CHECK(cudaMalloc((void **)&dc, sizeof(long long)));
    CN_CUDA_CHECK(cudaMalloc((void **)&sink, 512 * sizeof(uint32_t)));
    CN_CUDA_CHECK(cudaMemset(dc, 0, sizeof(long long)));
    CN_CUDA_CHECK(cudaFuncSetAttribute(tmem_ld_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768));
    tmem_ld_bench_kernel<<<1, nwarps * 32, 4 * 32768>>>(mode, iters, dc, sink);
    CN_LAUNCH_CHECK();
    CN_CUDA_CHECK(cudaDeviceSynchronize());
    CN_CUDA_CHECK(cudaMemcpy(cycles_host, dc, sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(dc); cudaFree(sink);
    return CN_OK;
}

// Developer diagnostic: same product, selectable operand mode (0 = A and B in smem, 1 = B MN-major, 2 = A in TMEM),
// repeated `reps` times; *cycles = clock64() ticks of the issue+commit+wait loop.
extern "C" int cn_debug_umma_bench(int32_t N, int32_t K, int32_t mode, int32_t reps, const float *a_host,
                                   const float *b_host, float *d_host, long long *cycles, int device)
{
    // mode >= 1000 encodes TS with explicit TMEM columns: mode = 1000 + a_col * 1000 + d_col  (a_col, d_col < 512)
    if (mode >= 1000) return selftest_impl(N, K, a_host, b_host, d_host, device, 7, reps, cycles, (mode - 1000) / 1000, (mode - 1000) % 1000);
    if (mode >= 600 && mode < 1000) return selftest_impl(N, K, a_host, b_host, d_host, device, 6, reps, cycles, 0, mode - 600);   // in-place TS, D at column mode-600
    return selftest_impl(N, K, a_host, b_host, d_host, device, mode, reps, cycles);
}
