// Shared helpers for libkukeon_gpuload: status codes, error strings, dtype table.
// Error convention mirrors the reference's sentinel + wrapped-message style
// (internal/errdefs/errdefs.go:23-, wrapped with %w at internal/ctr/container.go:561,586):
// a stable code plus a human-readable chain, never an exception across the C ABI.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/kukeon_gpuload.h"

namespace kk {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& msg) : std::runtime_error(msg), code(c) {}
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
inline void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

void set_last_error(const std::string& s);
const char* get_last_error();

struct DtypeInfo {
  const char* name;     // safetensors / ggml spelling
  uint32_t block_elems; // elements per storage block (1 for plain types)
  uint32_t block_bytes; // bytes per storage block; 0 => sub-byte type we do not size
  bool is_float;        // converted to bf16 in the pool
};

// Returns nullptr for unknown codes.
const DtypeInfo* dtype_info(uint32_t dt);
// safetensors "dtype" string -> kk_dtype, or -1.
int dtype_from_safetensors(const std::string& s);
// ggml type id -> kk_dtype, or -1 when this build does not handle it.
int dtype_from_ggml(uint32_t ggml_type);

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

}  // namespace kk
