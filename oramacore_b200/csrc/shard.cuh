// shard.cuh — K5: cross-shard exchange + merge (document-sharded index).
// (first slice: not built yet — returns OC_ERR_UNSUPPORTED)
#pragma once
static int run_sharded_merge(oc_ctx *c, const oc_search_params *p, const oc::FuseParams &fp, const oc::Bm25Params &bp,
                             bool has_ft, bool has_v, uint32_t B, uint32_t n_keep, uint32_t vlimit) {
    (void)c; (void)p; (void)fp; (void)bp; (void)has_ft; (void)has_v; (void)B; (void)n_keep; (void)vlimit;
    return fail(OC_ERR_UNSUPPORTED, "sharded merge not built yet");
}
