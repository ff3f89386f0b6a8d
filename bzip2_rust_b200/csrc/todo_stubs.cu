// todo_stubs.cu -- entry points declared in include/bz2b200.h that are not implemented yet.
// Each returns an error (never silently falls back to the CPU).  Shrinks as stages land.
#include "common.cuh"

#define NOT_YET(ctx) do { if (ctx) (ctx)->err = std::string(__func__) + ": not implemented yet"; return BZ2B200_E_ARG; } while (0)

extern "C" {
int bz2b200_bwt_decode(bz2b200_ctx *ctx, uint32_t, const uint8_t *, uint32_t, uint8_t *) { NOT_YET(ctx); }
int bz2b200_decompress_stream(bz2b200_ctx *ctx, const uint8_t *, size_t, uint8_t *, size_t, size_t *) { NOT_YET(ctx); }
}
