// placeholder until the Xception-UQ path lands (next commit)
#include "common.cuh"
extern "C" {
int bq_model_create(bq_ctx* ctx, const bq_model_config*, bq_model** out) { if (out) *out = nullptr; return bq_fail(ctx, BQ_ERR_STATE, "model path not built yet"); }
void bq_model_destroy(bq_model*) {}
int bq_model_load_weights(bq_model*, const bq_named_tensor*, int32_t) { return BQ_ERR_STATE; }
int bq_predict_uq(bq_model*, const uint8_t*, int64_t, int32_t, uint64_t, uint64_t, const uint8_t*, float*, float*, float*) { return BQ_ERR_STATE; }
int bq_model_debug_stage(bq_model*, const uint8_t*, int64_t, const char*, float*, int64_t, int64_t*) { return BQ_ERR_STATE; }
int bq_model_set_profiling(bq_model*, int) { return BQ_ERR_STATE; }
int bq_model_last_stage_ms(bq_model*, float*) { return BQ_ERR_STATE; }
}
