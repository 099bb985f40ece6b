// placeholder: model-level entry points are added next
#include "../../include/speedy_b200.h"
#include "ctx.h"
#include "abi_util.h"
namespace spd {
void model_create(speedy_ctx*) {}
void model_destroy(speedy_ctx*) {}
void upload_level_consts(speedy_ctx*) {}
}
extern "C" {
size_t speedy_output_len(const speedy_ctx* ctx) { return (size_t)(5 * ctx->d.kx + 1) * ctx->d.ngrid(); }
size_t speedy_state_len(const speedy_ctx* ctx) { return 0; }
const char* speedy_field_names(void) { return ""; }
}
