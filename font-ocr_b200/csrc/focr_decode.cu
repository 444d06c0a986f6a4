// focr_decode.cu -- placeholder: the focr entries report FOCR_ERR_UNSUPPORTED until the kernels land.
#include <string>
#include "common.cuh"
extern "C" int focr_glyph_bank_create(focr_ctx *, const uint8_t *, size_t, const focr_glyph_raster *, const float *,
                                      uint32_t, focr_glyph_bank **) { return FOCR_ERR_UNSUPPORTED; }
extern "C" void focr_glyph_bank_destroy(focr_glyph_bank *) {}
extern "C" int focr_decode_pages(focr_ctx *, const focr_glyph_bank *, const uint8_t *, size_t, uint32_t, uint32_t,
                                 uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t,
                                 uint16_t *, uint32_t *, uint32_t *, uint32_t *) { return FOCR_ERR_UNSUPPORTED; }
extern "C" int focr_sum_of_squares(focr_ctx *, const uint8_t *, const uint8_t *, size_t, uint32_t, int64_t *)
{ return FOCR_ERR_UNSUPPORTED; }
