// tcgen05 / TMEM MLP path (placeholder until the kernel lands).
#include "mn_model.cuh"

size_t mn_mlp_tc_workspace(const mn_model*, int64_t, int) { return 0; }
int mn_mlp_tc_pack(mn_ctx*, mn_model*, int, cudaStream_t) { return MN_OK; }
int mn_mlp_tc_launch(mn_ctx* ctx, mn_model*, const MlpArgs&, int64_t, int, void*, size_t, cudaStream_t) {
    return mn_fail(ctx, MN_ERR_UNSUPPORTED, "tensor-core MLP path not built yet");
}
