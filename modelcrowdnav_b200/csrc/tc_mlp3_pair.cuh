// tc_mlp3_pair.cuh -- mlp3 on the joint states + scoring on CTA pairs (tcgen05 cta_group::2), same dataflow skeleton
// as tc_rows_pair.cuh: per SM two tile contexts, 16 epilogue warps (2 contexts x 2 column halves x 4 TMEM lane
// quarters), warp 16 of the rank-0 CTA issues the UMMAs of the pair, warps 17 / 18 stream the joint-state tiles of
// context 0 / 1 from HBM with TMA bulk copies.  Rows = (env, action); each CTA keeps HALF of the mlp3 weights (42 KB).
//
//   stage  UMMA (M=256 over the pair)     A operand            D           epilogue
//   0      mlp3.0   K=80   N=160          J  (smem, by TMA)    [0,160)     ReLU -> U0, fp16 packed IN PLACE in TMEM
//   1      mlp3.2   K=160  N=112          U0 (TMEM, TS mode)   [120,232)   ReLU -> U1, fp16 packed in place
//   2      mlp3.4   K=112  N=112          U1 (TMEM, TS mode)   [0,112)     mlp3.6 as an fp32 dot on ReLU(acc),
//                                                                           value = reward + gamma_bar * V -> values[E][A]
// The activations never touch shared memory: each column-half warp packs its half of an accumulator in place
// (U0: K 0..79 at [0,40), K 80..159 at [80,120); U1: K 0..63 at [120,152), K 64..111 at [184,208)).
// MODE 1 runs CADRL's whole value network with the same three stages (cadrl.py:22-30: mlp(13 -> 150 -> 100 -> 100 -> 1) on
// every rotated (robot, human) row): stage 0 reads the X tiles of tc_features_kernel (K = 32, inputs split hi + lo) instead of
// the joint-state tiles, and the epilogue leaves one fp32 value per ROW; cadrl_min_kernel then takes the minimum over the
// humans of an (env, action) group (cadrl.py:166-168).
// Reference: crowd_nav/policy/sarl.py:62-64 (mlp3 on the joint state), multi_human_rl.py:52 (scoring); cadrl.py:131-178.

constexpr int kThreadsM3 = 608;
constexpr int M3_CTX_COLS = 256;
constexpr int M3_D1 = 120;                // mlp3.2 accumulator [120,232)

constexpr uint32_t HM_M1 = 0;                                        //  80 x 80
constexpr uint32_t HM_M2 = HM_M1 + bytes_of(N_H1 / 2, K_J);          //  56 x 160
constexpr uint32_t HM_M3 = HM_M2 + bytes_of(N_M1 / 2, N_H1);         //  56 x 112
constexpr uint32_t IMG_HM_BYTES = HM_M3 + bytes_of(N_M1 / 2, N_M1);

constexpr uint32_t M_RA_BYTES = J_TILE_BYTES;                        // 20 KB: J tile
constexpr uint32_t M_CTX_BYTES = M_RA_BYTES;
constexpr uint32_t M_CTX0 = (IMG_HM_BYTES + 127) & ~127u;
constexpr uint32_t M_MISC = M_CTX0 + 2 * M_CTX_BYTES;                // S[2][128] f32 | 10 mbarriers | tmem slot
constexpr uint32_t M_SMEM = M_MISC + 1024 + 96 + 16;
static_assert(M_SMEM <= 232448, "tc_mlp3_pair_kernel exceeds 227 KB of shared memory");

template <int MODE>       // 0: SARL mlp3 on joint-state tiles -> values[E][A];  1: CADRL network on X tiles -> rowv[tiles * 128]
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(88)
tc_mlp3_pair_kernel(EnvParams p, const double *__restrict__ st, const uint8_t *__restrict__ wimg,
                    const uint8_t *__restrict__ J, const double *__restrict__ rew, int A, int NG, double gamma,
                    double gamma_bar_host, double v_pref_host, double *__restrict__ values, int rounds, const TailW tw,
                    float *__restrict__ rowv)
{
    constexpr int K0 = MODE ? K_X : K_J;                                        // K of stage 0
    constexpr uint32_t TILE_BYTES = MODE ? X_TILE_BYTES : J_TILE_BYTES;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, ctx = (warp >> 2) & 1, hf = (warp >> 3) & 1;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int row = q * 32 + lane;
    float *S1 = reinterpret_cast<float *>(smem + M_MISC) + ctx * 128;       // partial dot of the upper column half
    const uint32_t bar0 = smem_u32(smem + M_MISC + 1024);
    const uint32_t req0 = bar0, req1 = bar0 + 8, done0 = bar0 + 16, done1 = bar0 + 24;
    const uint32_t xfull0 = bar0 + 32, xfull1 = bar0 + 40, xfree0 = bar0 + 48, xfree1 = bar0 + 56, xland0 = bar0 + 64, xland1 = bar0 + 72;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + M_MISC + 1024 + 96);

    copy_image_to_smem(smem, wimg + (size_t)rank * IMG_HM_BYTES, IMG_HM_BYTES);
    for (uint32_t i = tid * 16; i < 2 * M_CTX_BYTES; i += kThreadsM3 * 16)
        *reinterpret_cast<uint4 *>(smem + M_CTX0 + i) = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(req0, 16); mbar_init(req1, 16); mbar_init(done0, 1); mbar_init(done1, 1);
        mbar_init(xfull0, 2); mbar_init(xfull1, 2); mbar_init(xfree0, 1); mbar_init(xfree1, 1);
        mbar_init(xland0, 1); mbar_init(xland1, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc_2(smem_u32(tmem_slot), 512);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int tile_stride = 4 * nclusters;
    pdl_wait();                            // (programmatic dependent launch) the prologue above overlapped the producer kernel's tail
    pdl_launch_dependents();               // the argmax grid may be placed as this grid's CTAs exit

    if (warp == 16) {
        // ================= issuer (rank-0 CTA only) =================
        if (rank == 0 && lane == 0) {
            const uint32_t sM1 = smem_u32(smem + HM_M1), sM2 = smem_u32(smem + HM_M2), sM3 = smem_u32(smem + HM_M3);
            const int total = 3 * rounds;
            int stage0 = 0, stage1 = 0;
            uint32_t ph0 = 0, ph1 = 0, phx0 = 0, phx1 = 0;
            uint32_t idle_polls = 0;
            while (stage0 < total || stage1 < total) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    int &stage = c ? stage1 : stage0;
                    if (stage >= total) continue;
                    const int s = stage % 3;
                    uint32_t &ph = c ? ph1 : ph0;
                    if (!mbar_test_wait_cluster(c ? req1 : req0, ph)) continue;
                    if (s == 0) {                                   // stage 0 also needs the joint-state tile
                        uint32_t &phx = c ? phx1 : phx0;
                        if (!mbar_test_wait_cluster(c ? xfull1 : xfull0, phx)) continue;
                        phx ^= 1;
                    }
                    ph ^= 1;
                    fence_after_sync();
                    const uint32_t tm = tmem + (uint32_t)c * M3_CTX_COLS;
                    const uint32_t sRA = smem_u32(smem + M_CTX0 + (uint32_t)c * M_CTX_BYTES);
                    if (s == 0) mma_layer_2(tm, sRA, ROWS, sM1, K0, N_H1, false);
                    else if (s == 1) {
                        mma_steps_2_ts(tm + M3_D1, tm, sM2, 0, 5, N_M1, false);                 // U0, K 0..79
                        mma_steps_2_ts(tm + M3_D1, tm + 80, sM2, 5, 10, N_M1, true);            // U0, K 80..159
                    } else {
                        mma_steps_2_ts(tm, tm + M3_D1, sM3, 0, 4, N_M1, false);                 // U1, K 0..63
                        mma_steps_2_ts(tm, tm + M3_D1 + 64, sM3, 4, 7, N_M1, true);             // U1, K 64..111
                    }
                    commit_2(c ? done1 : done0, 3);
                    ++stage;
                    idle_polls = 0;
                }
                if (++idle_polls > (1u << 28)) __trap();        // protocol bug guard, counted in polls (see tc_rows_pair.cuh)
            }
        }
    } else if (warp > 16) {
        // ================= loader warps: warp 17 + c streams context c's joint-state tiles =================
        if (lane == 0) {
            const int c = warp - 17;
            const uint32_t xl = c ? xland1 : xland0, xfree = c ? xfree1 : xfree0;
            const uint32_t xfull_leader = mapa(c ? xfull1 : xfull0, 0);
            const uint32_t dst = smem_u32(smem + M_CTX0 + (uint32_t)c * M_CTX_BYTES);
            int tile = (cluster_id * 2 + (int)rank) * 2 + c;
            uint32_t phf = 0, phl = 0;
            for (int rnd = 0; rnd < rounds; ++rnd, tile += tile_stride) {
                if (rnd > 0) { mbar_wait_guarded(xfree, phf); phf ^= 1; }              // stage 0 of the previous tile is complete: its J tile is dead
                bulk_load(dst, J + (size_t)tile * TILE_BYTES, TILE_BYTES, xl);
                mbar_wait_guarded(xl, phl); phl ^= 1;
                mbar_arrive_cluster(xfull_leader);
            }
        }
    } else {
        // ================= epilogue warps =================
        const uint32_t done = ctx ? done1 : done0;
        const uint32_t req_leader = mapa(ctx ? req1 : req0, 0);
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)ctx * M3_CTX_COLS;
        uint32_t ph = 0;
        int tile = (cluster_id * 2 + (int)rank) * 2 + ctx;
#define M3_SIGNAL() do { fence_async_smem(); fence_before_sync(); __syncwarp(); if (lane == 0) mbar_arrive_cluster(req_leader); } while (0)
#define M3_WAIT() do { mbar_wait_guarded(done, ph); ph ^= 1; fence_after_sync(); } while (0)
        M3_SIGNAL();                                                               // stage 0 of the first tile
        for (int rnd = 0; rnd < rounds; ++rnd, tile += tile_stride) {
            // reward and robot v_pref of this thread's (env, action): fetched before the first completion wait, so that the two
            // global round trips are not left at the tail of the tile
            const long long g_row = (long long)tile * ROWS + row;
            double rew_g = 0.0, vp_g = v_pref_host;
            if (!MODE && hf == 0 && g_row < NG) {
                rew_g = rew[g_row];
                vp_g = st[st_idx(p.d, F_VPREF, 0, (int)(g_row / A))];
            }
            M3_WAIT();
            if (warp == 4 * ctx && lane == 0) mbar_arrive(ctx ? xfree1 : xfree0);  // J is dead: the next J tile may land
            compact_to_tmem<true, true>(tl, hf * 80, 80, hf * 80, 1.0f);                 // U0, packed in place
            M3_SIGNAL();
            M3_WAIT();
            if (hf == 0) compact_to_tmem<true, true>(tl, M3_D1, 64, M3_D1, 1.0f);        // U1, packed in place
            else compact_to_tmem<true>(tl, M3_D1 + 64, 48, M3_D1 + 64, 1.0f);
            M3_SIGNAL();
            M3_WAIT();
            // mlp3.6 as an fp32 dot over ReLU(mlp3.4), split over the column halves
            float part = 0.0f;
            if (hf == 0) {
                uint32_t v[32];                                 // one 32-column load per wait (see compact_to_tmem)
                ld32(tl, v);
                wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) part = fmaf(fmaxf(__uint_as_float(v[k]), 0.0f), tw.w[k], part);
                ld32(tl + 32, v);
                wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) part = fmaf(fmaxf(__uint_as_float(v[k]), 0.0f), tw.w[32 + k], part);
            } else {
                uint32_t x[32], y[16];
                ld32(tl + 64, x);
                wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) part = fmaf(fmaxf(__uint_as_float(x[k]), 0.0f), tw.w[64 + k], part);
                ld16(tl + 96, y);
                wait_ld();
#pragma unroll
                for (int k = 0; k < 4; ++k) part = fmaf(fmaxf(__uint_as_float(y[k]), 0.0f), tw.w[96 + k], part);
                S1[row] = part;
            }
            // every TMEM read of this tile is done: request stage 0 of the next tile before the scoring below
            if (rnd + 1 < rounds) M3_SIGNAL();
            if (ctx == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
            else asm volatile("bar.sync 2, 256;" ::: "memory");
            if (hf == 0) {
                const float v = part + S1[row] + tw.w[100];
                if (MODE) rowv[g_row] = v;                                          // one value per (env, action, human) row
                else if (g_row < NG) {
                    const double gamma_bar = (vp_g == v_pref_host) ? gamma_bar_host : pow(gamma, p.time_step * vp_g);
                    values[g_row] = rew_g + gamma_bar * (double)v;                // multi_human_rl.py:52
                }
            }
            // S1 is rewritten only after the next tile's three completion waits, which every hf-0 reader precedes
        }
#undef M3_SIGNAL
#undef M3_WAIT
    }
    fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_2(tmem, 512);
}
