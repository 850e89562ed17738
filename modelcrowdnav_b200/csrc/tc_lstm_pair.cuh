// tc_lstm_pair.cuh -- LSTM-RL's recurrence over the humans on CTA pairs (tcgen05 cta_group::2), same dataflow skeleton as
// tc_mlp3_pair.cuh: per SM two tile contexts, 16 epilogue warps (2 contexts x 2 column halves x 4 TMEM lane quarters), warp 16 of
// the rank-0 CTA issues the UMMAs of the pair, warps 17 / 18 stream the input tiles of context 0 / 1 from HBM with TMA bulk copies.
// Included by lookahead_tc.cu inside its anonymous namespace.
//
// Rows = (env, action) groups, 128 per CTA; ONE sequence per row.  Step t of a tile (t = 0 .. H-1, humans in predict()'s
// order: decreasing distance to the robot, lstm_rl.py:99-104):
//     gates[256 x 256] = X_t[256 x 32] * W_ih^T  (+)  h_{t-1}[256 x 112] * W_hh^T        (UMMA, M = 256 over the pair)
//     i, f, o = sigmoid, g = tanh;  c = f c + i g;  h = o tanh(c)                           (fp32, CUDA cores, c in registers)
// X_t is the rotated row of human ord[t] in the layout of the SARL path (K = 32: 13 features split hi + lo, ones columns
// carrying b_ih + b_hh); h goes back to shared memory as the next step's A operand, split hi + lo like every network input
// of this path (K = 112: hi at 0..55, lo at 56..111), so the only fp16 rounding is the weights'.  The 200 gate columns are
// permuted so that each column half (= one CTA's half of B) owns whole cells: half 0 cells 0..23, half 1 cells 24..49, as
// [i(32) | f(32) | g(32) | o(32)] -- a thread reads the four gates of 8 cells with four 8-column TMEM loads.  After the last
// step h_n replaces the weighted feature in the joint-state tile J (its self-state chunks come from tc_features_kernel) and
// tc_mlp3_pair_kernel<0> applies mlp(cat(self_state, h_n)) and the scoring.
// Reference: crowd_nav/policy/lstm_rl.py:9-34 (ValueNetwork1), torch.nn.LSTM gate order (i, f, g, o).

constexpr int kThreadsL = 608;
constexpr int N_LG = 256;                  // gate columns over the pair
constexpr int K_LH = 112;                  // h operand: hi at K 0..55, lo at K 56..111
constexpr int L_CELLS0 = 24;               // cells owned by column half 0; half 1 owns the remaining lstm_hidden - 24 (<= 32)
constexpr int L_HIDDEN = 50;

constexpr uint32_t HL_WIH = 0;                                          // 128 x 32
constexpr uint32_t HL_WHH = HL_WIH + bytes_of(N_LG / 2, K_X);           // 128 x 112
constexpr uint32_t IMG_HL_BYTES = HL_WHH + bytes_of(N_LG / 2, K_LH);

constexpr uint32_t L_H_BYTES = bytes_of(ROWS, K_LH);                    // 28 KB: h_{t-1}
constexpr uint32_t L_CTX_BYTES = X_TILE_BYTES + L_H_BYTES;
constexpr uint32_t L_CTX0 = (IMG_HL_BYTES + 127) & ~127u;
constexpr uint32_t L_MISC = L_CTX0 + 2 * L_CTX_BYTES;                   // 10 mbarriers | tmem slot
constexpr uint32_t L_SMEM = L_MISC + 96 + 16;
static_assert(L_SMEM <= 232448, "tc_lstm_pair_kernel exceeds 227 KB of shared memory");

// sigmoid and tanh from one MUFU.EX2 + one MUFU.RCP each (absolute error ~2e-7; tanh(x) = 2 sigmoid(2x) - 1 loses only RELATIVE
// accuracy near 0, where h = o tanh(c) is small anyway): the cell update is 5 transcendentals per cell and step, and with the libm
// forms the epilogue, not the UMMA chain, bounds the kernel.
__device__ __forceinline__ float sigmoid_f32(float x)
{
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));      // exp(-x); +inf for x << 0 -> 1 / inf = 0
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));                      // (no IEEE slow path: __frcp_rn calls one)
    return r;
}
__device__ __forceinline__ float tanh_f32(float x) { return fmaf(2.0f, sigmoid_f32(2.0f * x), -1.0f); }

// X holds, for tile T and step t, the 8 KB operand tile at (T * H + t); J receives chunks 0..6 of tile T
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsL, 1)
tc_lstm_pair_kernel(const uint8_t *__restrict__ wimg, const uint8_t *__restrict__ X, uint8_t *__restrict__ J, int H, int rounds)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, ctx = (warp >> 2) & 1, hf = (warp >> 3) & 1;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int row = q * 32 + lane;
    const uint32_t bar0 = smem_u32(smem + L_MISC);
    const uint32_t req0 = bar0, req1 = bar0 + 8, done0 = bar0 + 16, done1 = bar0 + 24;
    const uint32_t xfull0 = bar0 + 32, xfull1 = bar0 + 40, xfree0 = bar0 + 48, xfree1 = bar0 + 56, xland0 = bar0 + 64, xland1 = bar0 + 72;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L_MISC + 96);

    copy_image_to_smem(smem, wimg + (size_t)rank * IMG_HL_BYTES, IMG_HL_BYTES);
    for (uint32_t i = tid * 16; i < 2 * L_CTX_BYTES; i += kThreadsL * 16)
        *reinterpret_cast<uint4 *>(smem + L_CTX0 + i) = make_uint4(0, 0, 0, 0);
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

    if (warp == 16) {
        // ================= issuer (rank-0 CTA only): one stage per (tile, step) =================
        if (rank == 0 && lane == 0) {
            const uint32_t sWih = smem_u32(smem + HL_WIH), sWhh = smem_u32(smem + HL_WHH);
            const int total = H * rounds;
            int stage0 = 0, stage1 = 0;
            uint32_t ph0 = 0, ph1 = 0, phx0 = 0, phx1 = 0;
            uint32_t idle_polls = 0;
            while (stage0 < total || stage1 < total) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    int &stage = c ? stage1 : stage0;
                    if (stage >= total) continue;
                    uint32_t &ph = c ? ph1 : ph0;
                    if (!mbar_test_wait_cluster(c ? req1 : req0, ph)) continue;       // h_{t-1} stored, gates of t-1 read
                    uint32_t &phx = c ? phx1 : phx0;
                    if (!mbar_test_wait_cluster(c ? xfull1 : xfull0, phx)) continue;  // X_t of both CTAs has landed
                    phx ^= 1;
                    ph ^= 1;
                    fence_after_sync();
                    const uint32_t tm = tmem + (uint32_t)c * 256;
                    const uint32_t sX = smem_u32(smem + L_CTX0 + (uint32_t)c * L_CTX_BYTES), sH = sX + X_TILE_BYTES;
                    mma_layer_2(tm, sX, ROWS, sWih, K_X, N_LG, false);
                    if (stage % H != 0) mma_layer_2(tm, sH, ROWS, sWhh, K_LH, N_LG, true);     // h_0 = 0 (lstm_rl.py:29-31)
                    commit_2(c ? done1 : done0, 3);
                    ++stage;
                    idle_polls = 0;
                }
                if (++idle_polls > (1u << 28)) __trap();        // protocol bug guard, counted in polls (see tc_rows_pair.cuh)
            }
        }
    } else if (warp > 16) {
        // ================= loader warps: warp 17 + c streams context c's X_t tiles =================
        if (lane == 0) {
            const int c = warp - 17;
            const uint32_t xl = c ? xland1 : xland0, xfree = c ? xfree1 : xfree0;
            const uint32_t xfull_leader = mapa(c ? xfull1 : xfull0, 0);
            const uint32_t dst = smem_u32(smem + L_CTX0 + (uint32_t)c * L_CTX_BYTES);
            int tile = (cluster_id * 2 + (int)rank) * 2 + c;
            uint32_t phf = 0, phl = 0;
            bool first = true;
            for (int rnd = 0; rnd < rounds; ++rnd, tile += tile_stride)
                for (int t = 0; t < H; ++t) {
                    if (!first) { mbar_wait_guarded(xfree, phf); phf ^= 1; }           // the previous step's UMMAs are complete
                    first = false;
                    bulk_load(dst, X + ((size_t)tile * H + t) * X_TILE_BYTES, X_TILE_BYTES, xl);
                    mbar_wait_guarded(xl, phl); phl ^= 1;
                    mbar_arrive_cluster(xfull_leader);
                }
        }
    } else {
        // ================= epilogue warps: context `ctx`, column half `hf` (cells hf * 24 ...) =================
        const uint32_t done = ctx ? done1 : done0;
        const uint32_t req_leader = mapa(ctx ? req1 : req0, 0);
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)ctx * 256 + (uint32_t)hf * 128;
        uint8_t *Hop = smem + L_CTX0 + (uint32_t)ctx * L_CTX_BYTES + X_TILE_BYTES;
        const int nch = hf ? 4 : 3;                     // 8-cell chunks of this half (24 | 26 cells, zero-weight padding to 32)
        uint32_t ph = 0;
        int tile = (cluster_id * 2 + (int)rank) * 2 + ctx;
#define L_SIGNAL() do { fence_async_smem(); fence_before_sync(); __syncwarp(); if (lane == 0) mbar_arrive_cluster(req_leader); } while (0)
#define L_WAIT() do { mbar_wait_guarded(done, ph); ph ^= 1; fence_after_sync(); } while (0)
        L_SIGNAL();                                                                // step 0 of the first tile
        for (int rnd = 0; rnd < rounds; ++rnd, tile += tile_stride) {
            float cst[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) cst[k] = 0.0f;                            // c_0 = 0
            for (int t = 0; t < H; ++t) {
                const bool last = t == H - 1;
                L_WAIT();
                if (warp == 4 * ctx && lane == 0) mbar_arrive(ctx ? xfree1 : xfree0);      // X_t and h_{t-1} are consumed
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8) {
                    if (k8 >= nch) continue;
                    uint32_t gi[8], gf[8], gg[8], go[8];
                    ld8(tl + 8 * k8, gi);
                    ld8(tl + 32 + 8 * k8, gf);
                    ld8(tl + 64 + 8 * k8, gg);
                    ld8(tl + 96 + 8 * k8, go);
                    wait_ld();
                    float hv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float ig = sigmoid_f32(__uint_as_float(gi[j])), fg = sigmoid_f32(__uint_as_float(gf[j]));
                        const float g = tanh_f32(__uint_as_float(gg[j])), og = sigmoid_f32(__uint_as_float(go[j]));
                        const float c = fg * cst[k8 * 8 + j] + ig * g;
                        cst[k8 * 8 + j] = c;
                        hv[j] = og * tanh_f32(c);
                    }
                    float hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) split_hl(hv[j], hi[j], lo[j]);
                    const uint4 chi = make_uint4(h2(hi[0], hi[1]), h2(hi[2], hi[3]), h2(hi[4], hi[5]), h2(hi[6], hi[7]));
                    const int ck = hf * 3 + k8;                                            // K chunk of cells [8 ck, 8 ck + 8)
                    if (!last) {
                        const uint4 clo = make_uint4(h2(lo[0], lo[1]), h2(lo[2], lo[3]), h2(lo[4], lo[5]), h2(lo[6], lo[7]));
                        *reinterpret_cast<uint4 *>(Hop + chunk_off(ROWS, row, ck)) = chi;
                        *reinterpret_cast<uint4 *>(Hop + chunk_off(ROWS, row, 7 + ck)) = clo;
                    } else {
                        // h_n -> columns 0..49 of the joint-state tile (the place of SARL's weighted feature), fp16
                        *reinterpret_cast<uint4 *>(J + (size_t)tile * J_TILE_BYTES + chunk_off(ROWS, row, ck)) = chi;
                    }
                }
                if (!last || rnd + 1 < rounds) L_SIGNAL();     // h_t stored / gates read: the next step (or tile) may be issued
            }
        }
#undef L_SIGNAL
#undef L_WAIT
    }
    fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_2(tmem, 512);
}

// predict()'s human order per env: decreasing distance to the robot, stable (lstm_rl.py:99-104); env order with query_env
// (the next human states then come back from the env unsorted, multi_human_rl.py:37-38)
__global__ void lstm_order_kernel(EnvParams p, const double *__restrict__ st, int query_env, int32_t *__restrict__ ord)
{
    const EnvDims ed = p.d;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ed.E) return;
    const int H = ed.H;
    int o[CN_MAX_HUMANS];
    double dist[CN_MAX_HUMANS];
    const double rx = st[st_idx(ed, F_PX, 0, e)], ry = st[st_idx(ed, F_PY, 0, e)];
    for (int h = 0; h < H; ++h) {
        o[h] = h;
        dist[h] = norm2d(st[st_idx(ed, F_PX, h + 1, e)] - rx, st[st_idx(ed, F_PY, h + 1, e)] - ry);
    }
    if (!query_env)
        for (int i = 1; i < H; ++i) {
            const int oi = o[i];
            const double di = dist[oi];
            int j = i - 1;
            while (j >= 0 && dist[o[j]] < di) { o[j + 1] = o[j]; --j; }
            o[j + 1] = oi;
        }
    for (int t = 0; t < H; ++t) ord[(size_t)e * H + t] = o[t];
}

// X_t operand tiles of tc_lstm_pair_kernel: block (tile, t), thread = group row; the rotated row of human ord[t]
__global__ void __launch_bounds__(ROWS)
tc_features_lstm_kernel(EnvParams p, const double *__restrict__ st, const double *__restrict__ time,
                        const double *__restrict__ human_v, const double *__restrict__ actions, int A, int query_env, int NG,
                        const double *__restrict__ theta, const int32_t *__restrict__ ord, uint8_t *__restrict__ X)
{
    const EnvDims ed = p.d;
    const int H = ed.H;
    const int tile = blockIdx.x / H, t = blockIdx.x - tile * H, r = threadIdx.x;
    const long long g = (long long)tile * ROWS + r;
    const int h = g < NG ? ord[(size_t)(g / A) * H + t] : 0;
    RowInPP in;
    pp_load_inputs(in, ed, st, time, human_v, actions, A, query_env, NG, ROWS, tile, r, h, p.kinematics, theta);
    uint4 c0, c1, c2, c3;
    pair_features(in, p.time_step, p.kinematics, c0, c1, c2, c3);
    uint8_t *xt = X + (size_t)blockIdx.x * X_TILE_BYTES;
    *reinterpret_cast<uint4 *>(xt + chunk_off(ROWS, r, 0)) = c0;
    *reinterpret_cast<uint4 *>(xt + chunk_off(ROWS, r, 1)) = c1;
    *reinterpret_cast<uint4 *>(xt + chunk_off(ROWS, r, 2)) = c2;
    *reinterpret_cast<uint4 *>(xt + chunk_off(ROWS, r, 3)) = c3;
}
