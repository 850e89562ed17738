// lookahead_tc.cu -- SARL lookahead on tcgen05 tensor cores (CN_PREC_F16_TC). Placeholder until the
// UMMA kernel lands: creating a policy with this precision fails loudly, nothing falls back.
#include "cn_common.cuh"

int cn_tc_init(cn_policy *p) { (void)p; cn_set_error("CN_PREC_F16_TC is not built yet"); return CN_EUNSUPPORTED; }
void cn_tc_destroy(cn_policy *p) { (void)p; }
int cn_tc_load_weights(cn_policy *p, const float *flat_host, cudaStream_t s) { (void)p; (void)flat_host; (void)s; return CN_EUNSUPPORTED; }
int cn_lookahead_tc(cn_policy *p, cn_env *env, int query_env, double epsilon, cudaStream_t s)
{ (void)p; (void)env; (void)query_env; (void)epsilon; (void)s; return CN_EUNSUPPORTED; }
extern "C" int cn_selftest_umma(int32_t N, int32_t K, const float *a, const float *b, float *d, int device)
{ (void)N; (void)K; (void)a; (void)b; (void)d; (void)device; cn_set_error("not built yet"); return CN_EUNSUPPORTED; }
