#include "common.cuh"
namespace rsem_b200 {
int gibbs_run(rsem_b200_ctx*, const rsem_b200_gibbs_params*, rsem_b200_gibbs_out*) { set_error("K5 not built yet"); return RSEM_B200_ERR_UNSUPPORTED; }
}
