// shared helpers for the extern "C" layers
#pragma once
#include <string>
#include <stdexcept>
struct speedy_ctx;
namespace spd {
std::string& last_error();
void model_create(speedy_ctx* ctx);    // model.cu
void model_destroy(speedy_ctx* ctx);
}
#define API_BEGIN try {
#define API_END } catch (const std::exception& e_) { spd::last_error() = e_.what(); return -1; } return 0;
