// llk_internal.h -- what llk_engine.cu and llk_ingest.cu share inside libvb2llk.so (not part of the C ABI).
#ifndef VB2_LLK_INTERNAL_H_
#define VB2_LLK_INTERNAL_H_

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "llk_pack.h"
#include "vb2_llk.h"

namespace vb2 {

struct CreateParams {  // the fields of vb2_llk_desc that do not describe the sample
  int device;
  void *stream;
  uint32_t flags;
  int panel_dtype;
  double min_af, max_af;
};

vb2_llk_ctx *ctx_new();
int ctx_open(const CreateParams &cp, vb2_llk_ctx *ctx);               // device, stream, wait mode
PackConfig ctx_pack_config(const vb2_llk_ctx *ctx, const CreateParams &cp);
cudaStream_t ctx_stream(const vb2_llk_ctx *ctx);
// adopt an image that already sits in device memory (d_blob belongs to the context from here on)
int ctx_adopt_image(vb2_llk_ctx *ctx, const CreateParams &cp, const PackedSample &meta, uint8_t *d_blob);
int ctx_read_image(vb2_llk_ctx *ctx, uint8_t *dst, uint64_t n);
int ctx_error(vb2_llk_ctx *ctx, int code, const std::string &msg);    // records the message, returns code

}  // namespace vb2
#endif
