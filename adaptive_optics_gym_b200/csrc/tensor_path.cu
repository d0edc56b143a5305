// Tensor-core (tcgen05) arithmetic of the step path -- placeholder until the kernels land.
#include "tensor_path.cuh"

int aog_tensor_create(aog_env* env) { AOG_FAIL(AOG_ERR_UNSUPPORTED, "tensor precision path not built yet"); }
void aog_tensor_destroy(aog_env*) {}
int aog_tensor_table_updated(aog_env*, int, const void*) { return AOG_OK; }
int aog_tensor_screens_updated(aog_env*) { return AOG_OK; }
int aog_tensor_optics(aog_env* env, bool, bool, const aog_outputs&, cudaStream_t) {
  AOG_FAIL(AOG_ERR_UNSUPPORTED, "tensor precision path not built yet");
}
