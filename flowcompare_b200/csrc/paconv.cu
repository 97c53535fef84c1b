// PAConv (PointNet2SSGSeg) embedder -- not built yet: every entry reports FC_ERR_UNSUPPORTED.
#include "model.cuh"
int fc_paconv_create(FcCursor&, const int32_t*, int, fc_embedder*) { return FC_ERR_UNSUPPORTED; }
void fc_paconv_destroy(fc_embedder*) {}
int64_t fc_paconv_workspace_bytes(const fc_embedder*, int, int) { return FC_ERR_UNSUPPORTED; }
int fc_paconv_embed(const fc_embedder*, const float*, float*, int, int, void*, int64_t, int, cudaStream_t) { return FC_ERR_UNSUPPORTED; }
