// extern "C" entry points for hot path 2 (token -> waveform).  Kernels: decoder_kernels.cuh
#include "common.cuh"

using namespace ctb;

struct ctb_decoder { int dummy; };

extern "C" int64_t ctb_dvae_blob_floats(const ctb_convstack_config*) { return 0; }
extern "C" int64_t ctb_vocos_blob_floats(const ctb_vocos_config*) { return 0; }
extern "C" int ctb_decoder_create(const ctb_convstack_config*, const float*, const ctb_vocos_config*, const float*,
                                  int32_t, int32_t, ctb_decoder**) {
  return set_err(CTB_ERR_STATE, "decoder path not built yet");
}
extern "C" int ctb_decoder_destroy(ctb_decoder*) { return CTB_OK; }
extern "C" int ctb_dvae_decode(ctb_decoder*, const void*, int32_t, int32_t, int32_t, float*, void*) {
  return set_err(CTB_ERR_STATE, "decoder path not built yet");
}
extern "C" int ctb_vocos_decode(ctb_decoder*, const float*, int32_t, int32_t, float*, void*) {
  return set_err(CTB_ERR_STATE, "decoder path not built yet");
}
