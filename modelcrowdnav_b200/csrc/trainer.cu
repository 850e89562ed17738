// trainer.cu -- one optimisation step of the SARL value network on the device (CN_NET_SARL, FP32):
//   MSE( ValueNetwork(states), targets ) -> backward -> [gradient all-reduce by the caller] -> SGD with momentum.
// Replaces the ~150 kernel launches (+ one host sync) of the torch autograd step the reference runs per batch
// (crowd_nav/utils/trainer.py:61-82: zero_grad, model(inputs), MSELoss, backward, SGD(momentum 0.9).step, loss.item())
// with TWO kernels:
//   trainer_fwd_bwd_kernel   one CTA per kSPC samples of the batch: the whole forward (sarl.py:28-65) and backward of those
//                            samples with every activation in shared memory, weights streamed from L2 (386 kB, resident);
//                            per-CTA partial weight gradients -> gpart[cta][n_params]
//   trainer_reduce_kernel    fixed-order sum of the partials (deterministic) -> gradient; with grad_out == NULL it also
//                            applies SGD momentum in place and refreshes the transposed weight copy the forward reads
//   trainer_apply_kernel     the SGD half alone, for the data-parallel form (gradient -> NCCL all-reduce -> apply)
// FP32 FMA bound: ~3 x 62 k MAC per (sample, human) row; 100 x 5 rows = 0.1 GFLOP per step.
//
// Parameter layout = the torch state-dict order of ValueNetwork (mlp1.0, mlp1.2, mlp2.0, mlp2.2, attention.0, .2, .4,
// mlp3.0, .2, .4, .6; weight [out][in] then bias), the same flat block cn_policy_load_weights takes, so the torch model's
// parameters can be VIEWS of the block this trainer updates.
#include "cn_common.cuh"
#include "dense_f32.cuh"

#include <math.h>
#include <string.h>

#include <vector>

namespace {

using namespace dense_f32;      // kThreads, pad4, dense, weight_grad, relu_mask
constexpr int kSPC = 1;           // samples per CTA: 100 CTAs for the reference batch of 100; H <= 8 rows = one weight pass per layer
constexpr int kMaxH = 16;         // humans per sample supported by the shared-memory plan
constexpr int kLayers = 11;

struct TLayer { int in, out; int w_off, b_off, t_off; };   // offsets into the flat master / transposed blocks

struct TDims {
    TLayer L[kLayers];
    int n_params;
    int in, self_dim;
};

// shared-memory plan (floats); every row stride is a multiple of 4 floats
struct Plan {
    int R, ns;
    int ld_x, ld_a1, ld_e, ld_f1, ld_f, ld_u, ld_t, ld_j, ld_g1, ld_g;
    int x, a1, e, f1, f, u, t1, t2, sc, wt, c, j, g1, g2, g3, v, d0, d1, total;
};

__host__ __device__ inline Plan make_plan(const TDims &d, int H, int ns)
{
    Plan p;
    p.ns = ns; p.R = ns * H;
    const int R = p.R;
    p.ld_x = pad4(d.in); p.ld_a1 = pad4(d.L[0].out); p.ld_e = pad4(d.L[1].out); p.ld_f1 = pad4(d.L[2].out);
    p.ld_f = pad4(d.L[3].out); p.ld_u = pad4(d.L[4].in);
    int wt_ = d.L[4].out > d.L[5].out ? d.L[4].out : d.L[5].out;          // t1 / t2 rows also host de (mlp1's output gradient)
    if (d.L[1].out > wt_) wt_ = d.L[1].out;
    p.ld_t = pad4(wt_);
    p.ld_j = pad4(d.L[7].in); p.ld_g1 = pad4(d.L[7].out); p.ld_g = pad4(d.L[8].out > d.L[9].out ? d.L[8].out : d.L[9].out);
    int o = 0;
    p.x = o; o += R * p.ld_x;
    p.a1 = o; o += R * p.ld_a1;
    p.e = o; o += R * p.ld_e;
    p.f1 = o; o += R * p.ld_f1;
    p.f = o; o += R * p.ld_f;
    p.u = o; o += R * p.ld_u;             // attention input [e_i | mean]
    p.t1 = o; o += R * p.ld_t;
    p.t2 = o; o += R * p.ld_t;
    p.sc = o; o += pad4(R);
    p.wt = o; o += pad4(R);
    p.c = o; o += ns * p.ld_f;
    p.j = o; o += ns * p.ld_j;
    p.g1 = o; o += ns * p.ld_g1;
    p.g2 = o; o += ns * p.ld_g;
    p.g3 = o; o += ns * p.ld_g;
    p.v = o; o += 4;
    // two gradient scratch buffers, each as wide as the widest activation row
    int wide = p.ld_u;
    if (p.ld_a1 > wide) wide = p.ld_a1;
    if (p.ld_g1 > wide) wide = p.ld_g1;
    p.d0 = o; o += R * wide;
    p.d1 = o; o += R * wide;
    p.total = o;
    return p;
}

__global__ void __launch_bounds__(kThreads, 1)
trainer_fwd_bwd_kernel(TDims d, const float *__restrict__ W, const float *__restrict__ Wt, const float *__restrict__ X,
                       const float *__restrict__ target, const int64_t *__restrict__ index, int B, int H,
                       float *__restrict__ gpart, float *__restrict__ loss_part)
{
    extern __shared__ __align__(16) float sm[];
    const int s0 = blockIdx.x * kSPC;
    const int ns = min(kSPC, B - s0);
    const Plan p = make_plan(d, H, ns);
    const int R = p.R, tid = threadIdx.x;
    float *xs = sm + p.x, *a1 = sm + p.a1, *e = sm + p.e, *f1 = sm + p.f1, *f = sm + p.f, *u = sm + p.u, *t1 = sm + p.t1,
          *t2 = sm + p.t2, *sc = sm + p.sc, *wt = sm + p.wt, *c = sm + p.c, *j = sm + p.j, *g1 = sm + p.g1, *g2 = sm + p.g2,
          *g3 = sm + p.g3, *v = sm + p.v, *d0 = sm + p.d0, *d1 = sm + p.d1;
    float *G = gpart + (size_t)blockIdx.x * d.n_params;
    const TLayer *L = d.L;
    const int E1 = L[1].out, F = L[3].out;
#define WT(i) (Wt + L[i].t_off)
#define WM(i) (W + L[i].w_off)
#define BS(i) (W + L[i].b_off)

    // ---------------- forward (sarl.py:28-65) ----------------
    // sample s of the batch = item index[s0 + s] of the replay tensors (index == NULL: the batch is X / target itself)
    for (int idx = tid; idx < R * p.ld_x; idx += kThreads) {
        const int r = idx / p.ld_x, k = idx - r * p.ld_x;
        const size_t item = index ? (size_t)index[s0 + r / H] : (size_t)(s0 + r / H);
        xs[idx] = k < d.in ? X[(item * H + r % H) * d.in + k] : 0.0f;
    }
    __syncthreads();
    dense(xs, p.ld_x, R, L[0].in, WT(0), BS(0), L[0].out, a1, p.ld_a1, true, false);                 // mlp1.0 + ReLU
    __syncthreads();
    dense(a1, p.ld_a1, R, L[1].in, WT(1), BS(1), L[1].out, e, p.ld_e, true, false);                  // mlp1.2 + ReLU (last_relu)
    __syncthreads();
    dense(e, p.ld_e, R, L[2].in, WT(2), BS(2), L[2].out, f1, p.ld_f1, true, false);                  // mlp2.0 + ReLU
    // attention input u = [e_i | mean over the sample's humans]
    for (int idx = tid; idx < ns * E1; idx += kThreads) {
        const int s = idx / E1, k = idx - s * E1;
        float m = 0.0f;
        for (int h = 0; h < H; ++h) m += e[(size_t)(s * H + h) * p.ld_e + k];
        m /= (float)H;
        for (int h = 0; h < H; ++h) {
            u[(size_t)(s * H + h) * p.ld_u + k] = e[(size_t)(s * H + h) * p.ld_e + k];
            u[(size_t)(s * H + h) * p.ld_u + E1 + k] = m;
        }
    }
    __syncthreads();
    dense(f1, p.ld_f1, R, L[3].in, WT(3), BS(3), L[3].out, f, p.ld_f, false, false);                 // mlp2.2
    dense(u, p.ld_u, R, L[4].in, WT(4), BS(4), L[4].out, t1, p.ld_t, true, false);                   // attention.0 + ReLU
    __syncthreads();
    dense(t1, p.ld_t, R, L[5].in, WT(5), BS(5), L[5].out, t2, p.ld_t, true, false);                  // attention.2 + ReLU
    __syncthreads();
    dense(t2, p.ld_t, R, L[6].in, WT(6), BS(6), 1, sc, 1, false, false);                             // attention.4 -> score
    __syncthreads();
    if (tid < ns) {                                                                                   // masked softmax (sarl.py:52-53)
        float z = 0.0f;
        for (int h = 0; h < H; ++h) {
            const float s = sc[tid * H + h];
            const float ex = expf(s) * (s != 0.0f ? 1.0f : 0.0f);
            wt[tid * H + h] = ex; z += ex;
        }
        for (int h = 0; h < H; ++h) wt[tid * H + h] /= z;
    }
    __syncthreads();
    for (int idx = tid; idx < ns * p.ld_j; idx += kThreads) {                                        // joint = [self | weighted feature]
        const int s = idx / p.ld_j, k = idx - s * p.ld_j;
        float val = 0.0f;
        if (k < d.self_dim) val = xs[(size_t)(s * H) * p.ld_x + k];
        else if (k < d.self_dim + F) {
            const int kk = k - d.self_dim;
            for (int h = 0; h < H; ++h) val = fmaf(wt[s * H + h], f[(size_t)(s * H + h) * p.ld_f + kk], val);
            c[(size_t)s * p.ld_f + kk] = val;
        }
        j[idx] = val;
    }
    __syncthreads();
    dense(j, p.ld_j, ns, L[7].in, WT(7), BS(7), L[7].out, g1, p.ld_g1, true, false);                 // mlp3.0
    __syncthreads();
    dense(g1, p.ld_g1, ns, L[8].in, WT(8), BS(8), L[8].out, g2, p.ld_g, true, false);                // mlp3.2
    __syncthreads();
    dense(g2, p.ld_g, ns, L[9].in, WT(9), BS(9), L[9].out, g3, p.ld_g, true, false);                 // mlp3.4
    __syncthreads();
    dense(g3, p.ld_g, ns, L[10].in, WT(10), BS(10), 1, v, 1, false, false);                          // mlp3.6 -> value
    __syncthreads();

    // ---------------- loss and backward ----------------
    // MSELoss(mean): dL/dv_b = 2 (v_b - y_b) / B
    if (tid == 0) {
        float ls = 0.0f;
        for (int s = 0; s < ns; ++s) { const float df = v[s] - target[index ? index[s0 + s] : s0 + s]; ls += df * df; }
        loss_part[blockIdx.x] = ls;
    }
    float *dv = d0;                                                 // [ns][1]
    if (tid < ns) dv[tid] = 2.0f * (v[tid] - target[index ? index[s0 + tid] : s0 + tid]) / (float)B;
    __syncthreads();
    // mlp3.6
    weight_grad(dv, 1, g3, p.ld_g, ns, 1, L[10].in, G + L[10].w_off, G + L[10].b_off);
    float *dg3 = d1;                                                // [ns][ld_g]
    dense(dv, 1, ns, 1, WM(10), nullptr, L[10].in, dg3, p.ld_g, false, false);       // dg3 = dv * W[0][:]  (W as [1][in])
    __syncthreads();
    relu_mask(dg3, p.ld_g, g3, p.ld_g, ns, L[9].out);
    __syncthreads();
    // mlp3.4
    weight_grad(dg3, p.ld_g, g2, p.ld_g, ns, L[9].out, L[9].in, G + L[9].w_off, G + L[9].b_off);
    float *dg2 = d0;
    dense(dg3, p.ld_g, ns, L[9].out, WM(9), nullptr, L[9].in, dg2, p.ld_g, false, false);
    __syncthreads();
    relu_mask(dg2, p.ld_g, g2, p.ld_g, ns, L[8].out);
    __syncthreads();
    // mlp3.2
    weight_grad(dg2, p.ld_g, g1, p.ld_g1, ns, L[8].out, L[8].in, G + L[8].w_off, G + L[8].b_off);
    float *dg1 = d1;
    dense(dg2, p.ld_g, ns, L[8].out, WM(8), nullptr, L[8].in, dg1, p.ld_g1, false, false);
    __syncthreads();
    relu_mask(dg1, p.ld_g1, g1, p.ld_g1, ns, L[7].out);
    __syncthreads();
    // mlp3.0
    weight_grad(dg1, p.ld_g1, j, p.ld_j, ns, L[7].out, L[7].in, G + L[7].w_off, G + L[7].b_off);
    float *dj = d0;                                                 // [ns][ld_j]; the self-state part has no parameters upstream
    dense(dg1, p.ld_g1, ns, L[7].out, WM(7), nullptr, L[7].in, dj, p.ld_j, false, false);
    __syncthreads();
    // weighted feature c = sum_i w_i f_i :  df_i = w_i dc,  dw_i = dc . f_i ;  softmax: ds_i = w_i (dw_i - sum_k w_k dw_k)
    float *df = d1;                                                 // [R][ld_f]
    float *dsc = sm + p.sc;                                         // scores are dead after the softmax: reuse for ds
    if (tid < R) {
        const int s = tid / H;
        float dw = 0.0f;
        for (int k = 0; k < F; ++k) dw = fmaf(dj[(size_t)s * p.ld_j + d.self_dim + k], f[(size_t)tid * p.ld_f + k], dw);
        dsc[tid] = dw;                                              // dw_i for now
    }
    for (int idx = tid; idx < R * F; idx += kThreads) {
        const int r = idx / F, k = idx - r * F;
        df[(size_t)r * p.ld_f + k] = wt[r] * dj[(size_t)(r / H) * p.ld_j + d.self_dim + k];
    }
    __syncthreads();
    if (tid < ns) {
        float dot = 0.0f;
        for (int h = 0; h < H; ++h) dot = fmaf(wt[tid * H + h], dsc[tid * H + h], dot);
        for (int h = 0; h < H; ++h) dsc[tid * H + h] = wt[tid * H + h] * (dsc[tid * H + h] - dot);
    }
    __syncthreads();
    // mlp2.2 (no ReLU after it)
    weight_grad(df, p.ld_f, f1, p.ld_f1, R, L[3].out, L[3].in, G + L[3].w_off, G + L[3].b_off);
    float *df1 = d0;                                                // [R][ld_f1]  (dj is dead)
    dense(df, p.ld_f, R, L[3].out, WM(3), nullptr, L[3].in, df1, p.ld_f1, false, false);
    // attention.4
    weight_grad(dsc, 1, t2, p.ld_t, R, 1, L[6].in, G + L[6].w_off, G + L[6].b_off);
    __syncthreads();
    relu_mask(df1, p.ld_f1, f1, p.ld_f1, R, L[2].out);
    __syncthreads();
    // mlp2.0: de (accumulated in u's first half? no -- a dedicated buffer: reuse f (dead after df) is too narrow, use t-space later)
    weight_grad(df1, p.ld_f1, e, p.ld_e, R, L[2].out, L[2].in, G + L[2].w_off, G + L[2].b_off);
    float *dt2 = d1;                                                // [R][ld_t]  (df is dead after mlp2.2's two reads above)
    __syncthreads();
    dense(dsc, 1, R, 1, WM(6), nullptr, L[6].in, dt2, p.ld_t, false, false);         // dt2 = ds * W_att4[0][:]
    __syncthreads();
    relu_mask(dt2, p.ld_t, t2, p.ld_t, R, L[5].out);
    __syncthreads();
    // attention.2
    weight_grad(dt2, p.ld_t, t1, p.ld_t, R, L[5].out, L[5].in, G + L[5].w_off, G + L[5].b_off);
    float *dt1 = t2;                                                // t2 is dead once its mask and attention.4's gradient are taken
    __syncthreads();
    dense(dt2, p.ld_t, R, L[5].out, WM(5), nullptr, L[5].in, dt1, p.ld_t, false, false);
    __syncthreads();
    relu_mask(dt1, p.ld_t, t1, p.ld_t, R, L[4].out);
    __syncthreads();
    // attention.0 on u = [e | mean]
    weight_grad(dt1, p.ld_t, u, p.ld_u, R, L[4].out, L[4].in, G + L[4].w_off, G + L[4].b_off);
    float *du = d1;                                                 // [R][ld_u]  (dt2 is dead)
    __syncthreads();
    dense(dt1, p.ld_t, R, L[4].out, WM(4), nullptr, L[4].in, du, p.ld_u, false, false);
    __syncthreads();
    // de_i = du_i[:E1] + (1/H) sum_k du_k[E1:]  + mlp2.0's path (df1 W_20)
    float *de = t1;                                                 // t1 is dead; ld_t >= E1? e and t rows: use ld_e-compatible indexing below
    for (int idx = tid; idx < ns * E1; idx += kThreads) {
        const int s = idx / E1, k = idx - s * E1;
        float m = 0.0f;
        for (int h = 0; h < H; ++h) m += du[(size_t)(s * H + h) * p.ld_u + E1 + k];
        m /= (float)H;
        for (int h = 0; h < H; ++h) de[(size_t)(s * H + h) * p.ld_t + k] = du[(size_t)(s * H + h) * p.ld_u + k] + m;
    }
    __syncthreads();
    dense(df1, p.ld_f1, R, L[2].out, WM(2), nullptr, L[2].in, de, p.ld_t, false, true);   // += df1 * W_mlp2.0
    __syncthreads();
    relu_mask(de, p.ld_t, e, p.ld_e, R, L[1].out);
    __syncthreads();
    // mlp1.2
    weight_grad(de, p.ld_t, a1, p.ld_a1, R, L[1].out, L[1].in, G + L[1].w_off, G + L[1].b_off);
    float *da1 = d1;                                                // [R][ld_a1]  (du is dead)
    dense(de, p.ld_t, R, L[1].out, WM(1), nullptr, L[1].in, da1, p.ld_a1, false, false);
    __syncthreads();
    relu_mask(da1, p.ld_a1, a1, p.ld_a1, R, L[0].out);
    __syncthreads();
    // mlp1.0 (no gradient w.r.t. the input)
    weight_grad(da1, p.ld_a1, xs, p.ld_x, R, L[0].out, L[0].in, G + L[0].w_off, G + L[0].b_off);
#undef WT
#undef WM
#undef BS
}

// gradient = fixed-order sum of the per-CTA partials; loss = sum of the partial squared errors / B.
// grad_out != NULL: write the gradient (the caller all-reduces it, then trainer_apply_kernel).
// grad_out == NULL: apply SGD with momentum in place (torch.optim.SGD: buf = mu * buf + g; w -= lr * buf) and refresh Wt.
__global__ void trainer_reduce_kernel(int n_params, int nparts, const float *__restrict__ gpart, const float *__restrict__ loss_part,
                                      int B, float *__restrict__ grad_out, float *__restrict__ W, float *__restrict__ Wt,
                                      float *__restrict__ mom, const int32_t *__restrict__ tmap, float lr, float mu,
                                      float *__restrict__ loss_out, int loss_accumulate)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && loss_out) {
        float s = 0.0f;
        for (int c = 0; c < nparts; ++c) s += loss_part[c];
        *loss_out = (loss_accumulate ? *loss_out : 0.0f) + s / (float)B;
    }
    if (i >= n_params) return;
    float g = 0.0f;
    for (int c = 0; c < nparts; ++c) g += gpart[(size_t)c * n_params + i];
    if (grad_out) { grad_out[i] = g; return; }
    const float b = mu * mom[i] + g;
    mom[i] = b;
    const float w = W[i] - lr * b;
    W[i] = w;
    Wt[tmap[i]] = w;
}

__global__ void trainer_apply_kernel(int n_params, const float *__restrict__ grad, float scale, float *__restrict__ W,
                                     float *__restrict__ Wt, float *__restrict__ mom, const int32_t *__restrict__ tmap, float lr, float mu)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_params) return;
    const float b = mu * mom[i] + grad[i] * scale;
    mom[i] = b;
    const float w = W[i] - lr * b;
    W[i] = w;
    Wt[tmap[i]] = w;
}

__global__ void trainer_transpose_kernel(int n_params, const float *__restrict__ W, float *__restrict__ Wt, const int32_t *__restrict__ tmap)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_params) Wt[tmap[i]] = W[i];
}

}  // namespace

struct cn_trainer {
    int device;
    TDims d;
    int max_batch, max_humans;
    float *Wt, *mom, *gpart, *loss_part;
    int32_t *tmap;
    int nparts_cap;
    size_t smem_cap;
};

extern "C" {

int cn_trainer_create(const cn_sarl_cfg *cfg, int device, int32_t max_batch, int32_t max_humans, cn_trainer **out)
{
    if (!cfg || !out) { cn_set_error("null argument"); return CN_EINVAL; }
    if (cfg->network != CN_NET_SARL || cfg->with_om) {
        cn_set_error("the fused trainer covers the SARL value network without occupancy maps; use torch autograd for the others");
        return CN_EUNSUPPORTED;
    }
    if (max_batch < 1 || max_humans < 1 || max_humans > kMaxH) { cn_set_error("1 <= max_humans <= %d, max_batch >= 1", kMaxH); return CN_EINVAL; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); cn_set_error("no CUDA device available; no CPU fallback"); return CN_ECUDA; }
    if (device < 0 || device >= n) { cn_set_error("device %d out of range", device); return CN_EINVAL; }
    CN_CUDA_CHECK(cudaSetDevice(device));
    cn_trainer *t = new cn_trainer();
    memset(t, 0, sizeof(*t));
    t->device = device; t->max_batch = max_batch; t->max_humans = max_humans;
    TDims &d = t->d;
    const int ins[kLayers] = {cfg->input_dim, cfg->mlp1_dims[0], cfg->mlp1_dims[1], cfg->mlp2_dims[0], 2 * cfg->mlp1_dims[1],
                              cfg->attn_dims[0], cfg->attn_dims[1], cfg->mlp2_dims[1] + cfg->self_state_dim, cfg->mlp3_dims[0],
                              cfg->mlp3_dims[1], cfg->mlp3_dims[2]};
    const int outs[kLayers] = {cfg->mlp1_dims[0], cfg->mlp1_dims[1], cfg->mlp2_dims[0], cfg->mlp2_dims[1], cfg->attn_dims[0],
                               cfg->attn_dims[1], cfg->attn_dims[2], cfg->mlp3_dims[0], cfg->mlp3_dims[1], cfg->mlp3_dims[2],
                               cfg->mlp3_dims[3]};
    int off = 0;
    for (int i = 0; i < kLayers; ++i) {
        d.L[i].in = ins[i]; d.L[i].out = outs[i];
        d.L[i].w_off = off; d.L[i].t_off = off; off += ins[i] * outs[i];
        d.L[i].b_off = off; off += outs[i];
    }
    d.n_params = off; d.in = cfg->input_dim; d.self_dim = cfg->self_state_dim;
    if (outs[6] != 1 || outs[10] != 1) { delete t; cn_set_error("attention and mlp3 must end in 1 unit"); return CN_EINVAL; }
    // flat index -> index in the transposed block ([in][out] per layer; biases stay where they are)
    std::vector<int32_t> tmap(d.n_params);
    for (int i = 0; i < kLayers; ++i) {
        for (int o = 0; o < outs[i]; ++o)
            for (int k = 0; k < ins[i]; ++k) tmap[d.L[i].w_off + o * ins[i] + k] = d.L[i].t_off + k * outs[i] + o;
        for (int o = 0; o < outs[i]; ++o) tmap[d.L[i].b_off + o] = d.L[i].b_off + o;
    }
    t->nparts_cap = (max_batch + kSPC - 1) / kSPC;
    const Plan pl = make_plan(d, max_humans, kSPC);
    t->smem_cap = sizeof(float) * (size_t)pl.total;
    if (t->smem_cap > 232448) { delete t; cn_set_error("fused trainer needs %zu bytes of shared memory", t->smem_cap); return CN_EUNSUPPORTED; }
    if (cudaMalloc((void **)&t->Wt, sizeof(float) * d.n_params) != cudaSuccess ||
        cudaMalloc((void **)&t->mom, sizeof(float) * d.n_params) != cudaSuccess ||
        cudaMalloc((void **)&t->gpart, sizeof(float) * (size_t)d.n_params * t->nparts_cap) != cudaSuccess ||
        cudaMalloc((void **)&t->loss_part, sizeof(float) * t->nparts_cap) != cudaSuccess ||
        cudaMalloc((void **)&t->tmap, sizeof(int32_t) * d.n_params) != cudaSuccess) {
        cn_set_error("cudaMalloc failed in cn_trainer_create");
        cn_trainer_destroy(t);
        return CN_ENOMEM;
    }
    CN_CUDA_CHECK(cudaMemset(t->mom, 0, sizeof(float) * d.n_params));
    CN_CUDA_CHECK(cudaMemcpy(t->tmap, tmap.data(), sizeof(int32_t) * d.n_params, cudaMemcpyHostToDevice));
    CN_CUDA_CHECK(cudaFuncSetAttribute(trainer_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t->smem_cap));
    *out = t;
    return CN_OK;
}

int cn_trainer_destroy(cn_trainer *t)
{
    if (!t) return CN_OK;
    cudaSetDevice(t->device);
    void *ptrs[] = {t->Wt, t->mom, t->gpart, t->loss_part, t->tmap};
    for (void *q : ptrs) if (q) cudaFree(q);
    delete t;
    return CN_OK;
}

int64_t cn_trainer_param_count(const cn_trainer *t) { return t ? t->d.n_params : 0; }

int cn_trainer_sync_weights(cn_trainer *t, const float *w_dev, int zero_momentum, void *stream)
{
    if (!t || !w_dev) { cn_set_error("null argument"); return CN_EINVAL; }
    CN_CUDA_CHECK(cudaSetDevice(t->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int n = t->d.n_params;
    trainer_transpose_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, w_dev, t->Wt, t->tmap);
    CN_LAUNCH_CHECK();
    if (zero_momentum) CN_CUDA_CHECK(cudaMemsetAsync(t->mom, 0, sizeof(float) * n, s));
    return CN_OK;
}

static int trainer_step_impl(cn_trainer *t, float *w_dev, const float *states_dev, const float *targets_dev,
                             const int64_t *index_dev, int32_t batch, int32_t human_num, float lr, float momentum,
                             float *grad_out_dev, float *loss_dev, int loss_accumulate, void *stream)
{
    if (!t || !w_dev || !states_dev || !targets_dev) { cn_set_error("null argument"); return CN_EINVAL; }
    if (batch < 1 || batch > t->max_batch || human_num < 1 || human_num > t->max_humans) {
        cn_set_error("batch %d / human_num %d outside the trainer's capacity (%d, %d)", batch, human_num, t->max_batch, t->max_humans);
        return CN_EINVAL;
    }
    CN_CUDA_CHECK(cudaSetDevice(t->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int nparts = (batch + kSPC - 1) / kSPC;
    const Plan pl = make_plan(t->d, human_num, kSPC);
    const size_t smem = sizeof(float) * (size_t)pl.total;
    trainer_fwd_bwd_kernel<<<nparts, kThreads, smem, s>>>(t->d, w_dev, t->Wt, states_dev, targets_dev, index_dev, batch, human_num,
                                                         t->gpart, t->loss_part);
    CN_LAUNCH_CHECK();
    const int n = t->d.n_params;
    trainer_reduce_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, nparts, t->gpart, t->loss_part, batch, grad_out_dev, w_dev, t->Wt, t->mom,
                                                         t->tmap, lr, momentum, loss_dev, loss_accumulate);
    CN_LAUNCH_CHECK();
    return CN_OK;
}

int cn_trainer_step(cn_trainer *t, float *w_dev, const float *states_dev, const float *targets_dev, int32_t batch,
                    int32_t human_num, float lr, float momentum, float *grad_out_dev, float *loss_dev, void *stream)
{
    return trainer_step_impl(t, w_dev, states_dev, targets_dev, nullptr, batch, human_num, lr, momentum, grad_out_dev, loss_dev, 0,
                             stream);
}

int cn_trainer_step_indexed(cn_trainer *t, float *w_dev, const float *memory_states_dev, const float *memory_values_dev,
                            const int64_t *index_dev, int32_t batch, int32_t human_num, float lr, float momentum,
                            float *grad_out_dev, float *loss_sum_dev, void *stream)
{
    if (!index_dev) { cn_set_error("index_dev is null"); return CN_EINVAL; }
    return trainer_step_impl(t, w_dev, memory_states_dev, memory_values_dev, index_dev, batch, human_num, lr, momentum, grad_out_dev,
                             loss_sum_dev, 1, stream);
}

int cn_trainer_apply(cn_trainer *t, float *w_dev, const float *grad_dev, float grad_scale, float lr, float momentum, void *stream)
{
    if (!t || !w_dev || !grad_dev) { cn_set_error("null argument"); return CN_EINVAL; }
    CN_CUDA_CHECK(cudaSetDevice(t->device));
    const int n = t->d.n_params;
    trainer_apply_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, grad_dev, grad_scale, w_dev, t->Wt, t->mom, t->tmap, lr,
                                                                          momentum);
    CN_LAUNCH_CHECK();
    return CN_OK;
}

}  // extern "C"
