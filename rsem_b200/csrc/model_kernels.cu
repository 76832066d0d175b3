#include "common.cuh"
namespace rsem_b200 {
int model_launch_conprb(rsem_b200_ctx*) { set_error("K1 not built yet"); return RSEM_B200_ERR_UNSUPPORTED; }
int model_launch_update(rsem_b200_ctx*) { set_error("K3 not built yet"); return RSEM_B200_ERR_UNSUPPORTED; }
}
