// dist_stub.cu — placeholder entry points of the multi-GPU path until dist.cu lands.
#include "common.cuh"
using namespace pbf;
extern "C" {
int pbf_dist_unique_id(uint8_t *) { return PBF_ERR_STATE; }
int pbf_dist_init(pbf_ctx *ctx, const uint8_t *, int, int) { return fail(ctx, PBF_ERR_STATE, "pbf_dist_init", "multi-GPU path not built"); }
int pbf_dist_upload(pbf_ctx *ctx, const pbf_params *, const pbf_particle *, uint64_t) { return fail(ctx, PBF_ERR_STATE, "pbf_dist_upload", "multi-GPU path not built"); }
int pbf_dist_step(pbf_ctx *ctx, const pbf_params *) { return fail(ctx, PBF_ERR_STATE, "pbf_dist_step", "multi-GPU path not built"); }
int pbf_dist_download(pbf_ctx *ctx, pbf_particle *, uint64_t, uint64_t *) { return fail(ctx, PBF_ERR_STATE, "pbf_dist_download", "multi-GPU path not built"); }
int pbf_dist_stats_read(pbf_ctx *ctx, pbf_dist_stats *) { return fail(ctx, PBF_ERR_STATE, "pbf_dist_stats_read", "multi-GPU path not built"); }
int pbf_host_plan_splits(const uint64_t *, uint32_t, uint32_t, int, uint32_t *) { return PBF_ERR_STATE; }
}
