// Small driver for host/modelhub.hpp.
//   modelhub_test pull <path>                    -> JSON index on stdout (CPU only)
//   modelhub_test load <path> <container_dir>    -> Load on device 0, Mount, print {"checksum0":..,"stats":..} (GPU)
#include <cinttypes>
#include <cstdio>

#include "modelhub.hpp"

using namespace kukeon;

static const char* dtype_name(uint32_t d) {
  static const char* names[] = {"BOOL", "F4", "F6_E2M3", "F6_E3M2", "U8", "I8", "F8_E5M2", "F8_E4M3", "F8_E8M0", "I16", "U16",
                                "F16", "BF16", "I32", "U32", "F32", "C64", "F64", "I64", "U64"};
  if (d < 20) return names[d];
  switch (d) {
    case KK_Q4_0: return "Q4_0"; case KK_Q4_1: return "Q4_1"; case KK_Q5_0: return "Q5_0"; case KK_Q5_1: return "Q5_1";
    case KK_Q8_0: return "Q8_0"; case KK_Q2_K: return "Q2_K"; case KK_Q3_K: return "Q3_K"; case KK_Q4_K: return "Q4_K";
    case KK_Q5_K: return "Q5_K"; case KK_Q6_K: return "Q6_K"; case KK_Q8_K: return "Q8_K";
    default: return "?";
  }
}

int main(int argc, char** argv) {
  try {
    if (argc >= 3 && std::string(argv[1]) == "pull") {
      modelhub::TensorIndex ix = modelhub::Pull(argv[2]);
      printf("{\"shards\":%zu,\"file_bytes\":%" PRIu64 ",\"tensors\":[", ix.shards.size(), ix.file_bytes());
      for (size_t i = 0; i < ix.tensors.size(); ++i) {
        const kk_tensor_meta& t = ix.tensors[i];
        printf("%s{\"name\":\"%s\",\"dtype\":\"%s\",\"shape\":[", i ? "," : "", t.name, dtype_name(t.dtype));
        for (uint32_t d = 0; d < t.n_dims; ++d) printf("%s%" PRIu64, d ? "," : "", t.shape[d]);
        printf("],\"shard\":%u,\"file_offset\":%" PRIu64 ",\"nbytes\":%" PRIu64 "}", t.shard, t.file_offset, t.nbytes);
      }
      printf("]}\n");
      return 0;
    }
    if (argc >= 4 && std::string(argv[1]) == "load") {
      gpupool::Config cfg;
      cfg.staging_buffers = 2;
      cfg.staging_buffer_bytes = 4 << 20;
      cfg.reader_threads = 1;
      gpupool::Pool pool(cfg);
      gpupool::Model m = modelhub::Load(pool, argv[2]);
      gpupool::Model again = modelhub::Load(pool, argv[2]);  // second "session": same resident copy
      int refs = m.Info().refcount;
      modelhub::MountSpec spec = modelhub::Mount(m, 0, argv[3]);
      kk_model_info mi = m.Info();
      uint64_t sum = m.Checksum(0, 0, mi.pool_bytes / 8 * 8);
      printf("{\"refcount_two_sessions\":%d,\"same_handle\":%s,\"pool_bytes\":%" PRIu64 ",\"checksum\":%" PRIu64
             ",\"mount_source\":\"%s\",\"env0\":\"%s\",\"stats\":%s}\n",
             refs, m.handle() == again.handle() ? "true" : "false", mi.pool_bytes, sum, spec.mounts[0].source.c_str(), spec.env[0].c_str(),
             m.Stats().c_str());
      again.Release();
      m.Release();
      pool.Close();
      return 0;
    }
    fprintf(stderr, "usage: %s pull <path> | load <path> <container_dir>\n", argv[0]);
    return 2;
  } catch (const errdefs::Error& e) {
    fprintf(stderr, "error(%d): %s\n", e.code, e.what());
    return 1;
  }
}
