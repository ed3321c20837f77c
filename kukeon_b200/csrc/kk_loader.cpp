// Host side of the loader: staging pipeline (pread -> pinned ring -> H2D -> convert/fan-out kernel),
// pool allocation, refcounted model registry, IPC export, resident-image measurement mode.
//
// Pipeline per ingesting device:
//   R reader threads, each with its own CUDA stream and >= 2 pinned slots (+ matching device staging
//   buffers).  A thread claims the next chunk, waits for its slot's previous kernel (event), preads the
//   chunk's file ranges into the pinned slot, enqueues H2D + the convert kernel on its stream and moves
//   on.  Disk/page-cache reads, PCIe DMA and the kernels of different threads overlap.  The kernel writes
//   bf16 into the local pool and, in BROADCAST mode, into every peer pool over NVLink in the same pass.
#include "kk_loader.hpp"

#include <fcntl.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/vfs.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <exception>
#include <sstream>
#include <thread>

#include "kk_json.hpp"

namespace kk {

namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct FdSet {
  enum MapPolicy { kMapNone = 0, kMapTmpfs, kMapAll };  // which shards get a read-only mapping next to their descriptor
  std::vector<int> fds;
  std::vector<const uint8_t*> maps;  // read-only mappings (nullptr = none): long ranges are copied out of them with streaming stores, see read_chunk
  std::vector<uint64_t> sizes;
  explicit FdSet(const std::vector<std::string>& paths, MapPolicy policy = kMapNone) {
    for (auto& p : paths) {
      int fd = ::open(p.c_str(), O_RDONLY | O_CLOEXEC);
      if (fd < 0) {
        int e = errno;
        cleanup();
        fail(e == ENOENT ? KK_ENOENT : KK_EIO, "open %s: %s", p.c_str(), strerror(e));
      }
      fds.push_back(fd);
      const uint8_t* mp = nullptr;
      uint64_t sz = 0;
      bool want = policy == kMapAll;
      if (policy == kMapTmpfs) {
        struct statfs sf;
        want = fstatfs(fd, &sf) == 0 && (unsigned long)sf.f_type == 0x01021994ul;  // TMPFS_MAGIC: every byte is in memory, a mapping never waits for a disk
      }
      if (want) {
        struct stat st;
        if (fstat(fd, &st) == 0 && st.st_size > 0) {
          void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_SHARED, fd, 0);
          if (m != MAP_FAILED) { mp = (const uint8_t*)m; sz = (uint64_t)st.st_size; }
        }
      }
      maps.push_back(mp);
      sizes.push_back(sz);
    }
  }
  // Bytes of shard i that may be read through its mapping right now: touching a mapped page past the end of a file that has been truncated since
  // it was mapped raises SIGBUS, where pread reports a short read.  Asked once per chunk (one fstat per 16 MiB), this narrows that window from the
  // whole load to one chunk's copy; a range beyond the returned size takes the pread path and its KK_EIO.
  uint64_t live_size(uint32_t i) const {
    struct stat st;
    if (fstat(fds[i], &st) != 0 || st.st_size < 0) return 0;
    return std::min<uint64_t>((uint64_t)st.st_size, sizes[i]);
  }
  void cleanup() {
    for (size_t i = 0; i < maps.size(); ++i)
      if (maps[i]) munmap((void*)maps[i], sizes[i]);
    for (int f : fds) ::close(f);
    maps.clear();
    fds.clear();
  }
  ~FdSet() { cleanup(); }
  FdSet(const FdSet&) = delete;
  FdSet& operator=(const FdSet&) = delete;
};

void pread_full(int fd, uint8_t* dst, uint64_t len, uint64_t off, const std::string& name) {
  while (len) {
    ssize_t r = ::pread(fd, dst, len > (1ull << 30) ? (1ull << 30) : len, (off_t)off);
    if (r < 0) {
      if (errno == EINTR) continue;
      fail(KK_EIO, "pread %s @%llu: %s", name.c_str(), (unsigned long long)off, strerror(errno));
    }
    if (r == 0) fail(KK_EIO, "pread %s @%llu: unexpected end of file", name.c_str(), (unsigned long long)off);
    dst += r;
    off += (uint64_t)r;
    len -= (uint64_t)r;
  }
}

// CPUs of the NUMA node the device hangs off (empty when sysfs does not say).
std::vector<int> numa_cpus_of_device(int ordinal) {
  std::vector<int> cpus;
  char bdf[32] = {0};
  if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, ordinal) != cudaSuccess) { cudaGetLastError(); return cpus; }
  for (char* c = bdf; *c; ++c) *c = (char)tolower(*c);
  char path[256];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
  FILE* f = fopen(path, "r");
  if (!f) return cpus;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  if (node < 0) return cpus;
  snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
  f = fopen(path, "r");
  if (!f) return cpus;
  char buf[4096] = {0};
  if (fgets(buf, sizeof buf, f)) {
    char* save = nullptr;  // strtok_r: kk_open may run on several threads at once
    for (char* tok = strtok_r(buf, ",\n", &save); tok; tok = strtok_r(nullptr, ",\n", &save)) {
      int a, b;
      if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int i = a; i <= b; ++i) cpus.push_back(i); }
      else if (sscanf(tok, "%d", &a) == 1) cpus.push_back(a);
    }
  }
  fclose(f);
  return cpus;
}

void pin_this_thread(const std::vector<int>& cpus) {
  if (cpus.empty()) return;
  cpu_set_t set;
  CPU_ZERO(&set);
  for (int c : cpus)
    if (c >= 0 && c < CPU_SETSIZE) CPU_SET(c, &set);
  pthread_setaffinity_np(pthread_self(), sizeof set, &set);  // best effort
}

struct ErrorSink {
  std::mutex mu;
  std::exception_ptr first;
  std::atomic<bool> stop{false};
  void capture() {
    std::lock_guard<std::mutex> g(mu);
    if (!first) first = std::current_exception();
    stop = true;
  }
  void rethrow() {
    if (first) std::rethrow_exception(first);
  }
};

// CUDA events / scratch device memory that must not leak when a KK_CUDA check throws half way through a function.
struct EventSet {
  std::vector<cudaEvent_t> ev;
  explicit EventSet(size_t n = 0) : ev(n, nullptr) {}
  cudaEvent_t& operator[](size_t i) { return ev[i]; }
  size_t size() const { return ev.size(); }
  void create_all() {
    for (auto& e : ev) KK_CUDA(cudaEventCreate(&e));
  }
  ~EventSet() {
    for (auto e : ev)
      if (e) cudaEventDestroy(e);
  }
  EventSet(const EventSet&) = delete;
  EventSet& operator=(const EventSet&) = delete;
  EventSet(EventSet&& o) noexcept : ev(std::move(o.ev)) { o.ev.clear(); }
};
struct DevScratch {
  void* p = nullptr;
  ~DevScratch() {
    if (p) cudaFree(p);
  }
};

// How a chunk's long contiguous ranges travel from the page cache into the pinned slot.  Measured on the B200 box, Llama-3-8B (16 GB) from tmpfs,
// 16 reader threads (profiles/r02/e2e_read_modes_{o,p,q}.jsonl):
//   pread                       the kernel's copy_to_user: 0.31-0.33 s of copying per thread, the load is reader-bound at 0.33-0.34 s (47-49 GB/s)
//   mapping + memcpy            no faster than pread, and the final munmap of a checkpoint-sized mapping is ONE thread tearing down 4 M page-table
//                               entries: + 0.33 s per load
//   mapping + streaming stores  0.17-0.19 s per thread (no read-for-ownership of the slot's lines, which only the DMA engine reads next) — with the
//                               same munmap bill
//   ... + MADV_DONTNEED of each range right after its copy (the reader threads drop their own page-table entries in parallel, the final munmap
//                               finds none: 1 ms): the load becomes H2D-bound, 0.303-0.307 s = 52.5 GB/s = 0.96 of the pinned-H2D probe.  DEFAULT for
//                               shards on tmpfs; other file systems keep pread (a mapping of a cold file would fault page by page into the disk).
//   MAP_POPULATE mapping per range: 0.9 s per thread (every mmap / munmap takes the process's mm lock exclusively).
//   cache-resident bounce ring (4 x 256 KiB pinned pieces per reader between mapping and device slot, so that neither the CPU's stores nor the copy
//                               engine's reads reach DRAM): CPU copy 2x faster again, but 64 K small H2D copies starve the copy engines — 2 GPUs on one
//                               socket 288.6 -> 382.2 ms per step (profiles/r02/gpu_call_w.log).
// The measured losers are deleted (git history: ed69c4b..50a97f8 carry them); what is left is the default and one switch:
// KUKEON_GPULOAD_READ = pread (never map) | mapped (map every shard, whatever the file system).  Measurement knob, not API.
enum ReadMode { kReadAuto = 0, kReadPread, kReadMapped };
ReadMode read_mode() {
  static const ReadMode mode = [] {
    const char* e = getenv("KUKEON_GPULOAD_READ");
    if (!e) return kReadAuto;
    if (!strcmp(e, "pread")) return kReadPread;
    if (!strcmp(e, "mapped")) return kReadMapped;
    return kReadAuto;
  }();
  return mode;
}
FdSet::MapPolicy map_policy(const kk_model* m) {
  switch (read_mode()) {
    case kReadAuto: return FdSet::kMapTmpfs;
    case kReadPread: return FdSet::kMapNone;
    default: return FdSet::kMapAll;
  }
}
#if defined(__x86_64__)
__attribute__((target("avx2"))) void copy_nt_avx2(uint8_t* dst, const uint8_t* src, size_t n) {
  size_t head = (32u - ((uintptr_t)dst & 31u)) & 31u;
  if (head > n) head = n;
  if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
    _mm256_stream_si256((__m256i*)(dst + i), a);
    _mm256_stream_si256((__m256i*)(dst + i + 32), b);
    _mm256_stream_si256((__m256i*)(dst + i + 64), c);
    _mm256_stream_si256((__m256i*)(dst + i + 96), d);
  }
  if (i < n) memcpy(dst + i, src + i, n - i);
  _mm_sfence();  // the streaming stores must be globally visible before the H2D copy of the slot is enqueued
}
#endif
void copy_nt(uint8_t* dst, const uint8_t* src, size_t n) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) { copy_nt_avx2(dst, src, n); return; }
#endif
  memcpy(dst, src, n);
}

// Fill the slot with the chunk's file bytes.
void read_chunk(const Chunk& c, const FdSet& fds, const Index& ix, uint8_t* pinned) {
  static const uint64_t pg = (uint64_t)sysconf(_SC_PAGESIZE);
  const uint8_t* mp = fds.maps.empty() ? nullptr : fds.maps[c.shard];
  const uint64_t msz = mp ? fds.live_size(c.shard) : 0;
  const ReadMode mode = read_mode();
  for (auto& r : c.reads) {
    const bool in_map = mp && r.file_off <= msz && r.len <= msz - r.file_off;  // (a file that shrank since it was mapped takes pread's error path)
    if (in_map && r.len > (256u << 10) && mode != kReadPread) {
      copy_nt(pinned + r.buf_off, mp + r.file_off, r.len);
      const uint64_t a = (r.file_off + pg - 1) & ~(pg - 1), e = (r.file_off + r.len) & ~(pg - 1);  // the whole pages inside the range
      if (e > a) madvise((void*)(mp + a), (size_t)(e - a), MADV_DONTNEED);  // page-table entries only: the pages stay in the page cache
      continue;
    }
    // short ranges (column-slice rows of a SCATTER load: thousands of 2-7 KB runs per chunk) stay on pread: copied out of the mapping the 70B
    // scatter load measured 1.99 s against 1.51 s — a first-touch page fault per row costs more than the syscall (round 1, 8 x B200)
    pread_full(fds.fds[c.shard], pinned + r.buf_off, r.len, r.file_off, ix.shards[c.shard]);
  }
}

uint64_t seg_base_of(const Plan& P, int part) {
  uint64_t b = 0;
  for (int i = 0; i < part; ++i) b += P.parts[(size_t)i].segs.size();
  return b;
}

// Test hook: KUKEON_GPULOAD_TEST_NDST=<n> pads the destination list to n entries by repeating the ones it has,
// so that the kernel's n-destination store paths (the 8-GPU fan-out ladder) can be exercised on a box with fewer
// GPUs.  Writing the same bytes to the same pool several times is harmless.
void pad_dsts_for_test(ConvertLaunch& L) {
  static const int want = [] {
    const char* e = getenv("KUKEON_GPULOAD_TEST_NDST");
    return e ? atoi(e) : 0;
  }();
  if (want <= 0) return;
  const uint32_t have = L.n_dst;
  while (L.n_dst < (uint32_t)want && L.n_dst < KK_MAX_DST) { L.dst[L.n_dst] = L.dst[L.n_dst % have]; L.n_dst++; }
}

bool is_nvls(const kk_model* m) { return m->opts.fanout == KK_FANOUT_NVLS && m->plan.mode == KK_MODE_BROADCAST && m->nvls; }

// multimem.st exists for 4-, 8- and 16-byte accesses only: a plan qualifies for KK_LAUNCH_MULTIMEM when no segment ever needs a 1- or
// 2-byte store — every tile of it is whole 16-byte output vectors.  (Block dequantisers and the FP8/F16/F32 casts of whole groups
// always are; verbatim copies need a multiple of 16 bytes; transposes store single elements at tile edges and are excluded.)
bool plan_allows_multimem(const Plan& P, std::string* why) {
  for (auto& pp : P.parts)
    for (auto& s : pp.segs) {
      bool ok;
      switch (s.op) {
        case KK_OP_COPY: case KK_OP_F8E4M3_BF16: case KK_OP_F8E5M2_BF16: ok = s.units % 16 == 0; break;
        case KK_OP_F32_BF16: case KK_OP_F16_BF16: ok = s.units % 8 == 0; break;
        default: ok = kk_block_geom(s.op).block_bytes != 0; break;
      }
      if (!ok) {
        if (why) *why = "a tensor's size is not a whole number of 16-byte output vectors (or the load transposes); multimem.st cannot store its tail";
        return false;
      }
    }
  return true;
}

bool is_pull(const kk_model* m) { return m->opts.fanout == KK_FANOUT_PULL && m->plan.mode == KK_MODE_BROADCAST; }

// Pool bytes [lo, hi) that plan part `part` produces.  Pool order equals file order, so for every op that writes its output
// linearly this is one contiguous range and the ranges of different parts are disjoint.  Returns false when the part carries ops
// whose output is not linear in the pool (transposes, the scatter row exchange): such plans cannot be pulled slice by slice.
bool part_pool_range(const Plan& P, int part, uint64_t& lo, uint64_t& hi) {
  lo = UINT64_MAX;
  hi = 0;
  for (auto& s : P.parts[(size_t)part].segs) {
    uint64_t e;
    switch (s.op) {
      case KK_OP_COPY: e = s.dst_off + s.units; break;
      case KK_OP_F32_BF16: case KK_OP_F16_BF16: case KK_OP_F8E4M3_BF16: case KK_OP_F8E5M2_BF16: e = s.dst_off + s.units * 2; break;
      default: {
        const KKBlockGeom g = kk_block_geom(s.op);
        if (!g.block_bytes) return false;
        e = s.dst_off + s.units * g.out_bytes;
      }
    }
    if (s.dst_off < lo) lo = s.dst_off;
    if (e > hi) hi = e;
  }
  if (lo == UINT64_MAX) lo = hi = 0;
  return true;
}

// Destination pools a convert launch on local device `li` writes to.
void fill_dsts(kk_model* m, int li, ConvertLaunch& L) {
  L.n_dst = 0;
  L.flags = 0;
  for (auto& d : L.dst) d = nullptr;
  L.dst[L.n_dst++] = m->pools[(size_t)li];
  L.n_xdst = 0;
  for (auto& d : L.xdst) d = nullptr;
  if (m->plan.mode == KK_MODE_SCATTER && (m->plan.flags & KK_LOAD_SCATTER_EXCHANGE) && m->plan.n_parts > 1) {
    // all-to-all destinations of KK_OP_ROWSPLIT segments: the pool of every rank, indexed by rank
    const int n = m->plan.n_parts;
    for (int j = 0; j < n; ++j) {
      uint8_t* p = nullptr;
      if (m->opts.part_count > 1) p = (j == m->opts.part_index) ? m->pools[0] : (uint8_t*)m->peer_ptr[j];
      else if (m->ctx->peer_ok) p = m->pools[(size_t)j];
      if (!p) fail(KK_ESTATE, "KK_LOAD_SCATTER_EXCHANGE: the pool of rank %d is not reachable (attach it with kk_peer_attach, or enable peer access)", j);
      L.xdst[j] = p;
    }
    L.n_xdst = (uint32_t)n;
  }
  if (is_nvls(m)) {  // one multimem.st per vector reaches every pool through the switch
    L.dst[0] = m->nvls->multicast();
    L.n_dst = 1;
    L.flags |= KK_LAUNCH_MULTIMEM;
    return;
  }
  if (is_pull(m)) {  // stage 1 of a pull load: the same bytes into the slice buffer the peers will read (pool offset -> slice_buf - slice_base)
    if (m->slice_buf) L.dst[L.n_dst++] = (uint8_t*)((uintptr_t)m->slice_buf - (uintptr_t)m->slice_base);
    return;
  }
  if (m->plan.mode != KK_MODE_BROADCAST || m->opts.fanout == KK_FANOUT_NONE) {
    pad_dsts_for_test(L);
    return;
  }
  if (m->opts.part_count > 1) {
    for (int r = 0; r < KK_MAX_DEVICES; ++r)
      if (m->peer_ptr[r]) L.dst[L.n_dst++] = (uint8_t*)m->peer_ptr[r];
  } else if (m->ctx->peer_ok) {
    for (size_t j = 0; j < m->pools.size(); ++j)
      if ((int)j != li) L.dst[L.n_dst++] = m->pools[j];
  }
  pad_dsts_for_test(L);
}

// Tile scheduling of a launch: dynamic (counters of the launching stream) except for transposing loads.  Their 16-byte column stores fill a
// 32-byte sector only together with the neighbouring row group's tile, and under static round-robin those two tiles run at about the same time
// on different SMs and meet in L2; with dynamic draws GPT-2-small measured 0.173 ms against 0.146 ms static (profiles/r02/gpt2_quick_{e,h}.json).
uint32_t* sched_for(const kk_model* m, uint32_t* stream_counters) {
  // measurement knob (A/B of the two schedulers on one box; not part of the API): KUKEON_GPULOAD_SCHED=static | dynamic forces one
  static const int forced = [] {
    const char* e = getenv("KUKEON_GPULOAD_SCHED");
    return !e ? 0 : !strcmp(e, "static") ? 1 : !strcmp(e, "dynamic") ? 2 : 0;
  }();
  if (forced == 1) return nullptr;
  if (forced == 2) return stream_counters;
  return (m->plan.flags & KK_LOAD_GPT2_CONV1D_T) ? nullptr : stream_counters;
}

// Ingest plan part `part` on local device `li`.
void run_part(kk_model* m, int li, int part, const FdSet& fds) {
  kk_ctx* c = m->ctx;
  Device& dev = c->devs[(size_t)m->dev_idx[(size_t)li]];
  const PartPlan& pp = m->plan.parts[(size_t)part];
  if (pp.chunks.empty()) return;
  std::lock_guard<std::mutex> pipeline(*dev.load_mu);
  const KKSeg* d_segs = m->d_segs[(size_t)li] + seg_base_of(m->plan, part);
  ConvertLaunch base{};
  fill_dsts(m, li, base);
  const bool zerocopy = (c->cfg.flags & KK_CFG_ZEROCOPY) != 0;
  std::atomic<size_t> next{0};
  ErrorSink sink;
  auto ns = [] { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  auto worker = [&](Reader* rd) {
    try {
      if (!(c->cfg.flags & KK_CFG_NO_NUMA_PIN)) pin_this_thread(dev.numa_cpus);
      KK_CUDA(cudaSetDevice(dev.ordinal));
      size_t k = 0;
      uint64_t w_ns = 0, p_ns = 0, i_ns = 0;
      for (;;) {
        if (sink.stop) break;
        size_t ci = next.fetch_add(1);
        if (ci >= pp.chunks.size()) break;
        const Chunk& ch = pp.chunks[ci];
        Slot& s = rd->slots[k++ % rd->slots.size()];
        const uint64_t t0 = ns();
        KK_CUDA(cudaEventSynchronize(s.done));
        const uint64_t t1 = ns();
        read_chunk(ch, fds, m->plan.index, s.pinned);
        const uint64_t t2 = ns();
        ConvertLaunch L = base;
        if (zerocopy) {
          L.src = s.pinned;
        } else {
          KK_CUDA(cudaMemcpyAsync(s.dev, s.pinned, ch.buf_bytes, cudaMemcpyHostToDevice, rd->stream));
          L.src = s.dev;
        }
        L.segs = d_segs + ch.seg_begin;
        L.n_segs = ch.seg_count;
        L.n_tiles = ch.n_tiles;
        L.sched = sched_for(m, rd->sched);
        KK_CUDA(launch_convert(L, dev.sm_count, rd->stream));
        KK_CUDA(cudaEventRecord(s.done, rd->stream));
        w_ns += t1 - t0; p_ns += t2 - t1; i_ns += ns() - t2;
      }
      const uint64_t t3 = ns();
      KK_CUDA(cudaStreamSynchronize(rd->stream));
      m->rd_drain_ns += ns() - t3;
      m->rd_wait_ns += w_ns; m->rd_pread_ns += p_ns; m->rd_issue_ns += i_ns;
    } catch (...) {
      sink.capture();
    }
  };
  m->rd_wait_ns = 0; m->rd_pread_ns = 0; m->rd_issue_ns = 0; m->rd_drain_ns = 0;
  m->rd_threads = (uint32_t)dev.readers.size();
  std::vector<std::thread> th;
  // every reader runs on its own (NUMA-pinned) thread; the caller's thread affinity is left alone
  for (size_t r = 0; r < dev.readers.size(); ++r) th.emplace_back(worker, &dev.readers[r]);
  for (auto& t : th) t.join();
  if (sink.first) {
    // leave the streams quiet before reporting
    cudaSetDevice(dev.ordinal);
    for (auto& r : dev.readers) cudaStreamSynchronize(r.stream);
  }
  sink.rethrow();
}

// ---- KK_FANOUT_RAW ------------------------------------------------------------------------------
bool is_raw(const kk_model* m) { return m->opts.fanout == KK_FANOUT_RAW && m->plan.mode == KK_MODE_BROADCAST; }

// Peer raw images a stage-1 copy launch on local device li writes to (never the local image: it is the source).
void fill_raw_dsts(kk_model* m, int li, ConvertLaunch& L) {
  L.n_dst = 0;
  L.flags = 0;
  for (auto& d : L.dst) d = nullptr;
  if (m->opts.part_count > 1) {
    for (int r = 0; r < KK_MAX_DEVICES; ++r)
      if (m->peer_raw_ptr[r]) L.dst[L.n_dst++] = (uint8_t*)m->peer_raw_ptr[r];
  } else {
    for (size_t j = 0; j < m->raw.size(); ++j)
      if ((int)j != li) L.dst[L.n_dst++] = m->raw[j].image;
  }
}

// Stage 1 of a RAW load on local device li: file bytes of plan part `part` -> local raw image (H2D straight into
// place) -> peers' raw images (COPY launch per chunk, bulk TMA stores over NVLink).  fan_out=false: H2D only.
void run_part_raw(kk_model* m, int li, int part, const FdSet& fds, bool fan_out) {
  kk_ctx* c = m->ctx;
  Device& dev = c->devs[(size_t)m->dev_idx[(size_t)li]];
  const PartPlan& pp = m->plan.parts[(size_t)part];
  if (pp.chunks.empty()) return;
  std::lock_guard<std::mutex> pipeline(*dev.load_mu);
  kk_model::Raw& R = m->raw[(size_t)li];
  ConvertLaunch base{};
  fill_raw_dsts(m, li, base);
  const bool do_fan = fan_out && base.n_dst > 0;
  if (do_fan && !c->peer_ok && m->opts.part_count <= 1) fail(KK_EUNSUPPORTED, "KK_FANOUT_RAW needs peer access between the context's devices");
  std::atomic<size_t> next{0};
  ErrorSink sink;
  auto worker = [&](Reader* rd) {
    try {
      if (!(c->cfg.flags & KK_CFG_NO_NUMA_PIN)) pin_this_thread(dev.numa_cpus);
      KK_CUDA(cudaSetDevice(dev.ordinal));
      size_t k = 0;
      for (;;) {
        if (sink.stop) break;
        size_t ci = next.fetch_add(1);
        if (ci >= pp.chunks.size()) break;
        const Chunk& ch = pp.chunks[ci];
        Slot& s = rd->slots[k++ % rd->slots.size()];
        KK_CUDA(cudaEventSynchronize(s.done));
        read_chunk(ch, fds, m->plan.index, s.pinned);
        KK_CUDA(cudaMemcpyAsync(R.image + m->img_off[(size_t)part][ci], s.pinned, ch.buf_bytes, cudaMemcpyHostToDevice, rd->stream));
        if (do_fan) {
          ConvertLaunch L = base;
          L.src = R.image;
          L.segs = R.d_copy_segs + m->chunk_base[(size_t)part] + ci;
          L.n_segs = 1;
          L.n_tiles = (uint32_t)kk_seg_tiles(KK_OP_COPY, align_up(ch.buf_bytes, 16), 0);
          L.sched = sched_for(m, rd->sched);
        KK_CUDA(launch_convert(L, dev.sm_count, rd->stream));
        }
        KK_CUDA(cudaEventRecord(s.done, rd->stream));
      }
      KK_CUDA(cudaStreamSynchronize(rd->stream));
    } catch (...) {
      sink.capture();
    }
  };
  std::vector<std::thread> th;
  for (size_t r = 0; r < dev.readers.size(); ++r) th.emplace_back(worker, &dev.readers[r]);
  for (auto& t : th) t.join();
  if (sink.first) {
    cudaSetDevice(dev.ordinal);
    for (auto& r : dev.readers) cudaStreamSynchronize(r.stream);
  }
  sink.rethrow();
}

// Allocate the raw images and build the stage-1 (copy) and stage-2 (convert) segment tables.
void setup_raw(kk_model* m) {
  const Plan& P = m->plan;
  m->img_off.assign(P.parts.size(), {});
  m->chunk_base.assign(P.parts.size(), 0);
  uint64_t tot = 0;
  uint32_t nchunks = 0;
  for (size_t p = 0; p < P.parts.size(); ++p) {
    m->chunk_base[p] = nchunks;
    for (auto& ch : P.parts[p].chunks) {
      m->img_off[p].push_back(tot);
      tot += align_up(ch.buf_bytes + 64, 256);
      nchunks++;
    }
  }
  std::vector<KKSeg> copy_segs, conv_segs;
  std::vector<kk_model::Resident::Launch> launches;
  int last_shard = -1;
  for (size_t p = 0; p < P.parts.size(); ++p) {
    const PartPlan& pp = P.parts[p];
    for (size_t ci = 0; ci < pp.chunks.size(); ++ci) {
      const Chunk& ch = pp.chunks[ci];
      KKSeg cs{};
      cs.src_off = cs.dst_off = m->img_off[p][ci];
      cs.units = align_up(ch.buf_bytes, 16);
      cs.op = KK_OP_COPY;
      copy_segs.push_back(cs);
      bool fresh = launches.empty() || last_shard != (int)ch.shard || launches.back().n_segs + ch.seg_count > kMaxSegsPerLaunch ||
                   (uint64_t)launches.back().n_tiles + ch.n_tiles > 0xFFFFFFF0ull;
      if (fresh) launches.push_back({(uint32_t)conv_segs.size(), 0, 0, 0, 0});
      last_shard = (int)ch.shard;
      auto& L = launches.back();
      for (uint32_t j = 0; j < ch.seg_count; ++j) {
        KKSeg sg = pp.segs[ch.seg_begin + j];
        sg.src_off += m->img_off[p][ci];
        sg.tile_begin += L.n_tiles;
        conv_segs.push_back(sg);
      }
      L.n_segs += ch.seg_count;
      L.n_tiles += ch.n_tiles;
      L.src_bytes += ch.src_bytes;
      L.out_bytes += ch.out_bytes;
    }
  }
  m->raw.resize(m->dev_idx.size());
  for (size_t li = 0; li < m->dev_idx.size(); ++li) {
    Device& d = m->ctx->devs[(size_t)m->dev_idx[li]];
    KK_CUDA(cudaSetDevice(d.ordinal));
    auto& R = m->raw[li];
    R.bytes = align_up(tot ? tot : 256, 2u << 20);  // 2 MiB multiples: cheap to map over CUDA IPC (see the slice buffer in model_load)
    cudaError_t e = cudaMalloc((void**)&R.image, R.bytes);
    if (e != cudaSuccess) { cudaGetLastError(); R.image = nullptr; fail(KK_ENOMEM, "device %d: cudaMalloc(%llu) for the raw image failed", d.ordinal, (unsigned long long)R.bytes); }
    if (!copy_segs.empty()) {
      KK_CUDA(cudaMalloc((void**)&R.d_copy_segs, copy_segs.size() * sizeof(KKSeg)));
      KK_CUDA(cudaMemcpy(R.d_copy_segs, copy_segs.data(), copy_segs.size() * sizeof(KKSeg), cudaMemcpyHostToDevice));
    }
    if (!conv_segs.empty()) {
      KK_CUDA(cudaMalloc((void**)&R.d_conv_segs, conv_segs.size() * sizeof(KKSeg)));
      KK_CUDA(cudaMemcpy(R.d_conv_segs, conv_segs.data(), conv_segs.size() * sizeof(KKSeg), cudaMemcpyHostToDevice));
    }
    R.conv_launches = launches;
  }
}

void free_raw(kk_model* m) {
  for (size_t li = 0; li < m->raw.size(); ++li) {
    auto& R = m->raw[li];
    cudaSetDevice(m->ctx->devs[(size_t)m->dev_idx[li]].ordinal);
    if (R.image) cudaFree(R.image);
    if (R.d_copy_segs) cudaFree(R.d_copy_segs);
    if (R.d_conv_segs) cudaFree(R.d_conv_segs);
  }
  m->raw.clear();
}

// Stage 2: every local device converts the whole gathered image into its own pool.
void convert_local_all(kk_model* m, float* ms_total) {
  kk_ctx* c = m->ctx;
  const size_t nl = m->dev_idx.size();
  EventSet e0(nl), e1(nl);
  for (size_t li = 0; li < nl; ++li) {
    Device& dev = c->devs[(size_t)m->dev_idx[li]];
    auto& R = m->raw[li];
    KK_CUDA(cudaSetDevice(dev.ordinal));
    KK_CUDA(cudaEventCreate(&e0[li]));
    KK_CUDA(cudaEventCreate(&e1[li]));
    KK_CUDA(cudaEventRecord(e0[li], dev.stream));
    for (auto& La : R.conv_launches) {
      ConvertLaunch L{};
      L.src = R.image;
      L.segs = R.d_conv_segs + La.seg_begin;
      L.n_segs = La.n_segs;
      L.n_tiles = La.n_tiles;
      L.n_dst = 1;
      L.dst[0] = m->pools[li];
      L.sched = sched_for(m, dev.sched);
      KK_CUDA(launch_convert(L, dev.sm_count, dev.stream));
    }
    KK_CUDA(cudaEventRecord(e1[li], dev.stream));
  }
  float worst = 0.f;
  for (size_t li = 0; li < nl; ++li) {
    Device& dev = c->devs[(size_t)m->dev_idx[li]];
    KK_CUDA(cudaSetDevice(dev.ordinal));
    KK_CUDA(cudaStreamSynchronize(dev.stream));
    float ms = 0.f;
    KK_CUDA(cudaEventElapsedTime(&ms, e0[li], e1[li]));
    if (ms > worst) worst = ms;
  }
  if (ms_total) *ms_total = worst;
}

// Stage 2 of a KK_FANOUT_PULL load: one COPY launch whose segments read every attached peer's slice buffer over NVLink (bulk loads)
// and write the local pool.  Segments start with the peer after this rank so that the N ranks do not all read GPU 0 first.
void pull_slices(kk_model* m, float* ms_total) {
  kk_ctx* c = m->ctx;
  Device& dev = c->devs[(size_t)m->dev_idx[0]];
  const int n = m->opts.part_count, me = m->opts.part_index;
  std::vector<int> peers;
  std::lock_guard<std::mutex> peer_lock(m->peer_mu);
  // Every rank owns a slice buffer (never smaller than 256 B) and every peer's must be attached before stage 2, whether or not the
  // planner gave that rank any bytes: a rank that skipped the exchange is a protocol error of the caller, and reporting success for it
  // only because its part happened to be empty (a checkpoint smaller than one staging chunk) would hide the same mistake at full size.
  for (int k = 1; k < n; ++k) {
    const int r = (me + k) % n;
    if (!m->peer_slice_ptr[r]) fail(KK_ESTATE, "KK_FANOUT_PULL: the slice buffer of rank %d is not attached (kk_peer_attach_buffer, KK_BUF_SLICE)", r);
    if (m->part_range[(size_t)r].second > m->part_range[(size_t)r].first) peers.push_back(r);
  }
  KK_CUDA(cudaSetDevice(dev.ordinal));
  EventSet ev(2);
  ev.create_all();
  if (peers.empty()) {
    KK_CUDA(cudaEventRecord(ev[0], dev.stream));
    KK_CUDA(cudaEventRecord(ev[1], dev.stream));
  } else {
    // one src base for the launch: the numerically lowest peer mapping; every segment's src_off is its distance from it
    uintptr_t base = UINTPTR_MAX;
    for (int r : peers) base = std::min(base, (uintptr_t)m->peer_slice_ptr[r]);
    base &= ~(uintptr_t)15;
    std::vector<KKSeg> segs;
    uint64_t tiles = 0;
    for (int r : peers) {
      const uint64_t lo = m->part_range[(size_t)r].first, hi = m->part_range[(size_t)r].second;
      KKSeg sg{};
      sg.src_off = (uint64_t)((uintptr_t)m->peer_slice_ptr[r] - base) + (lo - (lo & ~(uint64_t)255));  // peers allocate from slice_base = lo & ~255
      sg.dst_off = lo;
      sg.units = hi - lo;
      sg.op = KK_OP_COPY;
      sg.tile_begin = (uint32_t)tiles;
      tiles += kk_seg_tiles(KK_OP_COPY, sg.units, 0);
      segs.push_back(sg);
    }
    if (tiles > 0xFFFFFFF0ull) fail(KK_EUNSUPPORTED, "too many tiles for one pull launch");
    DevScratch tmp;
    KK_CUDA(cudaMalloc(&tmp.p, segs.size() * sizeof(KKSeg)));
    KK_CUDA(cudaMemcpyAsync(tmp.p, segs.data(), segs.size() * sizeof(KKSeg), cudaMemcpyHostToDevice, dev.stream));
    ConvertLaunch L{};
    L.src = (const uint8_t*)base;
    L.segs = (const KKSeg*)tmp.p;
    L.n_segs = (uint32_t)segs.size();
    L.n_tiles = (uint32_t)tiles;
    L.n_dst = 1;
    L.dst[0] = m->pools[0];
    KK_CUDA(cudaEventRecord(ev[0], dev.stream));
    L.sched = sched_for(m, dev.sched);
      KK_CUDA(launch_convert(L, dev.sm_count, dev.stream));
    KK_CUDA(cudaEventRecord(ev[1], dev.stream));
    KK_CUDA(cudaStreamSynchronize(dev.stream));  // tmp outlives the launch
  }
  KK_CUDA(cudaStreamSynchronize(dev.stream));
  float ms = 0.f;
  KK_CUDA(cudaEventElapsedTime(&ms, ev[0], ev[1]));
  if (ms_total) *ms_total = ms;
}

void free_resident(kk_model* m) {
  for (size_t li = 0; li < m->resident.size(); ++li) {
    auto& R = m->resident[li];
    if (R.image || R.d_segs) {
      cudaSetDevice(m->ctx->devs[(size_t)m->dev_idx[li]].ordinal);
      if (R.image) cudaFree(R.image);
      if (R.d_segs) cudaFree(R.d_segs);
    }
  }
  m->resident.clear();
}

void destroy_model(kk_model* m) {
  kk_ctx* c = m->ctx;
  free_resident(m);
  for (int r = 0; r < KK_MAX_DEVICES; ++r)
    if (m->peer_raw_ptr[r]) {
      cudaSetDevice(c->devs[(size_t)m->dev_idx[0]].ordinal);
      cudaIpcCloseMemHandle(m->peer_raw_ptr[r]);
      m->peer_raw_ptr[r] = nullptr;
    }
  free_raw(m);
  for (int r = 0; r < KK_MAX_DEVICES; ++r)
    if (m->peer_slice_ptr[r]) {
      cudaSetDevice(c->devs[(size_t)m->dev_idx[0]].ordinal);
      if (m->peer_slice_is_ipc[r]) cudaIpcCloseMemHandle(m->peer_slice_ptr[r]);
      m->peer_slice_ptr[r] = nullptr;
    }
  if (m->slice_buf) {
    cudaSetDevice(c->devs[(size_t)m->dev_idx[0]].ordinal);
    cudaFree(m->slice_buf);
    m->slice_buf = nullptr;
  }
  for (int r = 0; r < KK_MAX_DEVICES; ++r)
    if (m->peer_ptr[r]) {
      cudaSetDevice(c->devs[(size_t)m->dev_idx[0]].ordinal);
      if (m->peer_is_ipc[r]) cudaIpcCloseMemHandle(m->peer_ptr[r]);
      m->peer_ptr[r] = nullptr;
    }
  for (size_t i = 0; i < m->pools.size(); ++i) {
    Device& d = c->devs[(size_t)m->dev_idx[i]];
    cudaSetDevice(d.ordinal);
    if (m->pools[i]) {
      if (!m->nvls && m->vmm.empty()) cudaFree(m->pools[i]);  // NVLS / VMM pools are unmapped and released by their owners below
      std::lock_guard<std::mutex> g(c->mu);  // model_load checks the budget under the same lock
      d.pool_in_use -= m->pool_bytes[i];
    }
    if (i < m->d_segs.size() && m->d_segs[i]) cudaFree(m->d_segs[i]);
  }
  m->nvls.reset();
  m->vmm.clear();
  delete m;
}

std::string canon(const std::string& p) {
  char buf[4096];
  if (realpath(p.c_str(), buf)) return buf;
  return p;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
kk_ctx* ctx_open(const kk_config& cfg_in) {
  kk_config cfg = cfg_in;
  if (cfg.n_devices < 1 || cfg.n_devices > KK_MAX_DEVICES) fail(KK_EINVAL, "n_devices %d out of range 1..%d", cfg.n_devices, KK_MAX_DEVICES);
  for (int i = 0; i < cfg.n_devices; ++i)
    for (int j = 0; j < i; ++j)
      if (cfg.devices[i] == cfg.devices[j]) fail(KK_EINVAL, "device %d listed twice", cfg.devices[i]);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    fail(KK_ECUDA, "no usable CUDA device (%s); this library has no CPU path", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  for (int i = 0; i < cfg.n_devices; ++i)
    if (cfg.devices[i] < 0 || cfg.devices[i] >= count) fail(KK_EINVAL, "device ordinal %d not present (%d devices)", cfg.devices[i], count);
  if (cfg.n_reader_threads == 0) cfg.n_reader_threads = 16;  // 16 x ~4 GB/s of page-cache pread saturates a Gen5 x16 link (profiles/r01)
  if (cfg.n_staging_buffers == 0) cfg.n_staging_buffers = 2 * cfg.n_reader_threads;
  if (cfg.n_reader_threads > 64) fail(KK_EINVAL, "n_reader_threads %u too large", cfg.n_reader_threads);
  if (cfg.n_staging_buffers < cfg.n_reader_threads) cfg.n_staging_buffers = cfg.n_reader_threads;
  if (cfg.staging_buffer_bytes == 0) cfg.staging_buffer_bytes = 16ull << 20;
  if (cfg.staging_buffer_bytes < (1ull << 20)) fail(KK_EINVAL, "staging_buffer_bytes must be at least 1 MiB");

  std::unique_ptr<kk_ctx> c(new kk_ctx);
  c->cfg = cfg;
  c->slot_bytes = align_up(cfg.staging_buffer_bytes, 2ull << 20);
  c->devs.resize((size_t)cfg.n_devices);
  try {
    // One thread per device, all at once: the pinned ring (cudaHostAlloc pins and maps 0.5 GiB per device by default) dominates kk_open, and eight
    // devices set up one after the other cost 4.7 s in the one-process shape (profiles/r02/bench_n8_head.json: single_process.kk_open_s).  Each
    // thread is bound to its device's NUMA node, so the slots are first-touched / pinned there.
    std::vector<std::exception_ptr> errs((size_t)cfg.n_devices);
    std::vector<std::thread> setup;
    for (int i = 0; i < cfg.n_devices; ++i) {
      c->devs[(size_t)i].ordinal = cfg.devices[i];
      setup.emplace_back([&, i] {
        try {
          Device& d = c->devs[(size_t)i];
          KK_CUDA(cudaSetDevice(d.ordinal));
          cudaDeviceProp prop;
          KK_CUDA(cudaGetDeviceProperties(&prop, d.ordinal));
          if (prop.major < 10) fail(KK_EUNSUPPORTED, "device %d is sm_%d%d; this build carries sm_100a code only", d.ordinal, prop.major, prop.minor);
          d.sm_count = prop.multiProcessorCount;
          KK_CUDA(kernels_init_device());
          d.kernels_ready = true;
          KK_CUDA(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
          KK_CUDA(cudaMalloc((void**)&d.sched, 256));
          KK_CUDA(cudaMemset(d.sched, 0, 256));
          d.numa_cpus = numa_cpus_of_device(d.ordinal);
          d.readers.resize(cfg.n_reader_threads);
          if (!(cfg.flags & KK_CFG_NO_NUMA_PIN)) pin_this_thread(d.numa_cpus);
          for (uint32_t r = 0; r < cfg.n_reader_threads; ++r) {
            Reader& rd = d.readers[r];
            KK_CUDA(cudaStreamCreateWithFlags(&rd.stream, cudaStreamNonBlocking));
            KK_CUDA(cudaMalloc((void**)&rd.sched, 256));
            KK_CUDA(cudaMemset(rd.sched, 0, 256));
            uint32_t ns = cfg.n_staging_buffers / cfg.n_reader_threads + (r < cfg.n_staging_buffers % cfg.n_reader_threads ? 1 : 0);
            rd.slots.resize(ns);
            for (auto& s : rd.slots) {
              KK_CUDA(cudaHostAlloc((void**)&s.pinned, c->slot_bytes + 256, cudaHostAllocPortable | cudaHostAllocMapped));
              if (!(cfg.flags & KK_CFG_ZEROCOPY)) KK_CUDA(cudaMalloc((void**)&s.dev, c->slot_bytes + 256));
              KK_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
            }
          }
        } catch (...) {
          errs[(size_t)i] = std::current_exception();
        }
      });
    }
    for (auto& t : setup) t.join();
    for (auto& e : errs)
      if (e) std::rethrow_exception(e);
    if (cfg.flags & KK_CFG_PEER_ALL) {
      for (int i = 0; i < cfg.n_devices; ++i) {
        KK_CUDA(cudaSetDevice(cfg.devices[i]));
        for (int j = 0; j < count; ++j) {
          if (j == cfg.devices[i]) continue;
          int can = 0;
          if (cudaDeviceCanAccessPeer(&can, cfg.devices[i], j) != cudaSuccess || !can) { cudaGetLastError(); continue; }
          cudaError_t pe = cudaDeviceEnablePeerAccess(j, 0);
          if (pe != cudaSuccess) cudaGetLastError();  // already enabled or refused: kk_peer_attach will report a real failure
        }
      }
    }
    c->peer_ok = cfg.n_devices > 1 && !(cfg.flags & KK_CFG_NO_PEER_ACCESS);
    if (c->peer_ok) {
      for (int i = 0; i < cfg.n_devices && c->peer_ok; ++i)
        for (int j = 0; j < cfg.n_devices; ++j) {
          if (i == j) continue;
          int can = 0;
          KK_CUDA(cudaDeviceCanAccessPeer(&can, cfg.devices[i], cfg.devices[j]));
          if (!can) { c->peer_ok = false; break; }
          KK_CUDA(cudaSetDevice(cfg.devices[i]));
          cudaError_t pe = cudaDeviceEnablePeerAccess(cfg.devices[j], 0);
          if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
          else if (pe != cudaSuccess) { cudaGetLastError(); c->peer_ok = false; break; }
        }
    }
  } catch (...) {
    ctx_close(c.release());
    throw;
  }
  return c.release();
}

void ctx_close(kk_ctx* c) {
  if (!c) return;
  {
    std::lock_guard<std::mutex> g(c->mu);
    if (!c->models.empty()) fail(KK_EBUSY, "%zu model(s) still referenced", c->models.size());
  }
  for (auto& d : c->devs) {
    if (d.ordinal < 0) continue;
    cudaSetDevice(d.ordinal);
    for (auto& rd : d.readers) {
      for (auto& s : rd.slots) {
        if (s.done) cudaEventDestroy(s.done);
        if (s.dev) cudaFree(s.dev);
        if (s.pinned) cudaFreeHost(s.pinned);
      }
      if (rd.stream) cudaStreamDestroy(rd.stream);
      if (rd.sched) cudaFree(rd.sched);
    }
    if (d.stream) cudaStreamDestroy(d.stream);
    if (d.sched) cudaFree(d.sched);
  }
  delete c;
}

// ---------------------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------------------
static void do_load(kk_model* m) {
  const double t0 = now_s();
  FdSet fds(m->plan.index.shards, map_policy(m));
  m->t_files_open = now_s() - t0;
  const size_t nl = m->dev_idx.size();
  m->t_part.assign(nl, 0.0);
  const bool multi_proc = m->opts.part_count > 1;
  const bool replicas = !multi_proc && m->plan.mode == KK_MODE_BROADCAST && nl > 1 &&
                        (m->opts.fanout == KK_FANOUT_NONE || !m->ctx->peer_ok);
  ErrorSink sink;
  auto per_dev = [&](int li) {
    try {
      const double a = now_s();
      if (is_raw(m)) {
        run_part_raw(m, li, m->local_parts[(size_t)li], fds, true);
      } else if (replicas) {
        for (int p = 0; p < m->plan.n_parts; ++p) run_part(m, li, p, fds);
      } else {
        run_part(m, li, m->local_parts[(size_t)li], fds);
      }
      m->t_part[(size_t)li] = now_s() - a;
    } catch (...) {
      sink.capture();
    }
  };
  std::vector<std::thread> th;
  for (size_t li = 1; li < nl; ++li) th.emplace_back(per_dev, (int)li);
  per_dev(0);
  for (auto& t : th) t.join();
  sink.rethrow();
  if (is_raw(m)) {
    m->raw_staged = true;
    if (!multi_proc) {  // one process owns every device: all stage-1 work is done, convert now
      convert_local_all(m, nullptr);
      m->raw_staged = false;
    }
  }
  const double tc = now_s();
  fds.cleanup();  // part of the load: unmapping a checkpoint-sized mapping is not free
  m->t_files_close = now_s() - tc;
  m->t_load = now_s() - t0;
  m->n_loads++;
}

kk_model* model_load(kk_ctx* c, const std::string& path, const kk_load_opts& opts_in) {
  kk_load_opts opts = opts_in;
  if (opts.part_count <= 1) { opts.part_count = 1; opts.part_index = 0; }
  if (opts.part_index < 0 || opts.part_index >= opts.part_count || opts.part_count > KK_MAX_DEVICES)
    fail(KK_EINVAL, "part %d of %d out of range", opts.part_index, opts.part_count);
  if (opts.mode < KK_MODE_SINGLE || opts.mode > KK_MODE_SCATTER) fail(KK_EINVAL, "unknown mode %d", opts.mode);
  if (opts.fanout < KK_FANOUT_P2P || opts.fanout > KK_FANOUT_PULL) fail(KK_EINVAL, "unknown fanout %d", opts.fanout);
  if (opts.fanout == KK_FANOUT_NVLS) {
    // every "cannot do NVLS here" is KK_EUNSUPPORTED (the contract since the first ABI version), whatever the reason
    if (opts.mode != KK_MODE_BROADCAST) fail(KK_EUNSUPPORTED, "fan-out NVLS only applies to KK_MODE_BROADCAST");
    if (opts.part_count > 1) fail(KK_EUNSUPPORTED, "fan-out NVLS needs one process owning all devices (sharing a multicast object across processes is not implemented)");
    if (c->cfg.n_devices < 2) fail(KK_EUNSUPPORTED, "fan-out NVLS needs at least two devices in the context");
    std::vector<int> ords(c->cfg.devices, c->cfg.devices + c->cfg.n_devices);
    std::string why;
    if (!NvlsPools::supported(ords, &why)) fail(KK_EUNSUPPORTED, "fan-out NVLS: this host does not expose it: %s", why.c_str());
  }
  if (opts.fanout == KK_FANOUT_RAW && opts.mode != KK_MODE_BROADCAST) fail(KK_EINVAL, "KK_FANOUT_RAW only applies to KK_MODE_BROADCAST");
  const bool multi_proc = opts.part_count > 1;
  if (opts.fanout == KK_FANOUT_PULL && opts.mode != KK_MODE_BROADCAST) fail(KK_EINVAL, "KK_FANOUT_PULL only applies to KK_MODE_BROADCAST");
  if (opts.fanout == KK_FANOUT_PULL && !multi_proc)
    fail(KK_EINVAL, "KK_FANOUT_PULL is for one-process-per-GPU operation (part_count > 1); one process owning the GPUs fans out with P2P stores");
  if (multi_proc && c->cfg.n_devices != 1) fail(KK_EINVAL, "multi-process parts need a one-device context (got %d devices)", c->cfg.n_devices);
  if (multi_proc && opts.mode == KK_MODE_SINGLE) fail(KK_EINVAL, "KK_MODE_SINGLE cannot be split into parts");

  std::ostringstream ks;
  ks << canon(path) << "|m" << opts.mode << "|f" << opts.fanout << "|x" << (opts.flags & ~KK_LOAD_DEFER) << "|p" << opts.part_index << "/" << opts.part_count;
  const std::string key = ks.str();

  kk_model* m = nullptr;
  {
    std::unique_lock<std::mutex> lk(c->mu);
    for (;;) {
      auto it = c->models.find(key);
      if (it == c->models.end()) break;
      if (it->second->loading) { c->cv.wait(lk); continue; }
      it->second->refcount++;
      return it->second;
    }
    m = new kk_model;
    m->ctx = c;
    m->key = key;
    m->opts = opts;
    m->refcount = 1;
    m->loading = true;
    c->models[key] = m;
  }
  try {
    double t0 = now_s();
    Index ix = index_path(path);
    m->t_index = now_s() - t0;
    t0 = now_s();
    int n_parts = 1;
    if (opts.mode != KK_MODE_SINGLE) n_parts = multi_proc ? opts.part_count : c->cfg.n_devices;
    int mode = opts.mode;
    m->plan = build_plan(std::move(ix), mode, opts.flags & ~KK_LOAD_DEFER, n_parts, c->slot_bytes);
    m->t_plan = now_s() - t0;
    t0 = now_s();
    if (multi_proc || opts.mode == KK_MODE_SINGLE) {
      m->dev_idx = {0};
      m->local_parts = {multi_proc ? opts.part_index : 0};
    } else {
      for (int i = 0; i < c->cfg.n_devices; ++i) { m->dev_idx.push_back(i); m->local_parts.push_back(i); }
    }
    // concatenated segment table of all parts
    std::vector<KKSeg> all;
    for (auto& pp : m->plan.parts) all.insert(all.end(), pp.segs.begin(), pp.segs.end());
    m->pools.assign(m->dev_idx.size(), nullptr);
    m->pool_bytes.assign(m->dev_idx.size(), 0);
    m->d_segs.assign(m->dev_idx.size(), nullptr);
    if (opts.fanout == KK_FANOUT_NVLS) {
      std::string why;
      if (!plan_allows_multimem(m->plan, &why)) fail(KK_EUNSUPPORTED, "fan-out NVLS: %s", why.c_str());
      std::vector<int> ords;
      for (int di : m->dev_idx) ords.push_back(c->devs[(size_t)di].ordinal);
      KK_CUDA(cudaSetDevice(ords[0]));  // the driver entry points below want a current context on the calling thread
      m->nvls.reset(new NvlsPools);
      m->nvls->create(ords, m->plan.pool_bytes_of_part(0));
    }
    for (size_t li = 0; li < m->dev_idx.size(); ++li) {
      Device& d = c->devs[(size_t)m->dev_idx[li]];
      KK_CUDA(cudaSetDevice(d.ordinal));
      const uint64_t pb = m->plan.pool_bytes_of_part(m->local_parts[li]);
      {
        std::lock_guard<std::mutex> g(c->mu);
        if (c->cfg.pool_bytes_per_device && d.pool_in_use + pb > c->cfg.pool_bytes_per_device)
          fail(KK_ENOMEM, "device %d: pool budget exceeded (%llu in use + %llu > %llu)", d.ordinal, (unsigned long long)d.pool_in_use,
               (unsigned long long)pb, (unsigned long long)c->cfg.pool_bytes_per_device);
        d.pool_in_use += pb;
        m->pool_bytes[li] = pb;
      }
      cudaError_t e = cudaSuccess;
      if (m->nvls) {
        m->pools[li] = m->nvls->pool(li);  // already allocated, bound and mapped
      } else if (c->cfg.flags & KK_CFG_VMM_POOLS) {
        // cuMemCreate memory: exported as a POSIX fd (kk_export_fd) that another process maps READ-ONLY, which a cudaIpcMemHandle cannot offer.
        // Mapped read-write here for this device and, in a peer-enabled context, for the other devices (fan-out stores land in it).
        std::vector<int> acc;
        if (c->peer_ok)
          for (auto& dv : c->devs) acc.push_back(dv.ordinal);
        if (m->vmm.size() < m->dev_idx.size()) m->vmm.resize(m->dev_idx.size());
        try {
          m->vmm[li].reset(new VmmAlloc);
          m->vmm[li]->create(d.ordinal, pb, acc);
          m->pools[li] = m->vmm[li]->ptr();
        } catch (...) {
          m->vmm[li].reset();
          std::lock_guard<std::mutex> g(c->mu);
          d.pool_in_use -= pb;
          m->pool_bytes[li] = 0;
          throw;
        }
      } else {
        // whole 2 MiB multiples: the driver sub-allocates smaller requests out of shared 2 MiB blocks, and an IPC handle maps the whole block —
        // a pool that owns its blocks outright cannot expose a neighbouring allocation through its handle (round-1 review)
        e = cudaMalloc((void**)&m->pools[li], align_up(pb ? pb : 1, 2u << 20));
      }
      if (e != cudaSuccess) {
        cudaGetLastError();
        m->pools[li] = nullptr;
        std::lock_guard<std::mutex> g(c->mu);
        d.pool_in_use -= pb;
        m->pool_bytes[li] = 0;
        fail(KK_ENOMEM, "device %d: cudaMalloc(%llu) for the pool failed: %s", d.ordinal, (unsigned long long)pb, cudaGetErrorString(e));
      }
      if (!all.empty()) {
        KK_CUDA(cudaMalloc((void**)&m->d_segs[li], all.size() * sizeof(KKSeg)));
        KK_CUDA(cudaMemcpy(m->d_segs[li], all.data(), all.size() * sizeof(KKSeg), cudaMemcpyHostToDevice));
      }
    }
    if (is_raw(m)) setup_raw(m);
    if (is_pull(m)) {
      m->part_range.assign((size_t)m->plan.n_parts, {0, 0});
      for (int p = 0; p < m->plan.n_parts; ++p)
        if (!part_pool_range(m->plan, p, m->part_range[(size_t)p].first, m->part_range[(size_t)p].second))
          fail(KK_EUNSUPPORTED, "KK_FANOUT_PULL cannot be combined with transposing loads (their output is not one contiguous pool range per rank)");
      const auto& mine = m->part_range[(size_t)opts.part_index];
      m->slice_base = mine.first & ~(uint64_t)255;
      const uint64_t sb = mine.second > m->slice_base ? mine.second - m->slice_base : 0;
      KK_CUDA(cudaSetDevice(c->devs[(size_t)m->dev_idx[0]].ordinal));
      // whole 2 MiB multiples, like the pools: cudaIpcOpenMemHandle of such an allocation is ~50x cheaper than of one the driver carved out of
      // shared blocks (measured at N = 8: seven 16 GB pools map in 0.07 s, seven 2 GB slice buffers of odd size took 0.66 s — and in round 1,
      // before the pools were rounded, seven pools took 3.6 s; profiles/r02/bench_n8_*.json)
      cudaError_t se = cudaMalloc((void**)&m->slice_buf, align_up(sb ? sb : 256, 2u << 20));
      if (se != cudaSuccess) { cudaGetLastError(); m->slice_buf = nullptr; fail(KK_ENOMEM, "cudaMalloc(%llu) for the slice buffer failed", (unsigned long long)sb); }
    }
    m->t_alloc = now_s() - t0;
    if (!(opts.flags & KK_LOAD_DEFER)) {
      do_load(m);
      m->loaded = !((is_raw(m) || is_pull(m)) && multi_proc);
      if (is_raw(m) && !multi_proc) free_raw(m);  // the gathered file bytes are not needed once the pools are built
    }
  } catch (...) {
    {
      std::lock_guard<std::mutex> g(c->mu);
      c->models.erase(key);
    }
    c->cv.notify_all();
    destroy_model(m);
    throw;
  }
  {
    std::lock_guard<std::mutex> g(c->mu);
    m->loading = false;
  }
  c->cv.notify_all();
  return m;
}

void model_load_part(kk_model* m) {
  std::lock_guard<std::mutex> op(m->op_mu);
  if (is_raw(m) && m->raw.empty()) fail(KK_ESTATE, "the raw image of this model has been released");
  do_load(m);
  std::lock_guard<std::mutex> g(m->ctx->mu);
  if (!((is_raw(m) || is_pull(m)) && m->opts.part_count > 1)) m->loaded = true;  // multi-process RAW / PULL: loaded after kk_convert_local
}

void model_convert_local(kk_model* m, float* ms_total) {
  std::lock_guard<std::mutex> op(m->op_mu);
  if (is_pull(m)) {
    pull_slices(m, ms_total);
    std::lock_guard<std::mutex> g(m->ctx->mu);
    m->loaded = true;
    return;
  }
  if (!is_raw(m) || m->raw.empty()) fail(KK_ESTATE, "kk_convert_local only applies to KK_FANOUT_RAW models with a live raw image and to KK_FANOUT_PULL models");
  convert_local_all(m, ms_total);
  std::lock_guard<std::mutex> g(m->ctx->mu);
  m->loaded = true;
  m->raw_staged = false;
}

void model_export_raw(kk_model* m, int li, void* handle_out) {
  if (!is_raw(m) || m->raw.empty()) fail(KK_ESTATE, "this model has no raw image");
  KK_CUDA(cudaSetDevice(m->ctx->devs[(size_t)m->dev_idx[(size_t)li]].ordinal));
  cudaIpcMemHandle_t h;
  KK_CUDA(cudaIpcGetMemHandle(&h, m->raw[(size_t)li].image));
  memcpy(handle_out, &h, sizeof h);
}

void model_peer_attach_raw(kk_model* m, int rank, const void* handle) {
  std::lock_guard<std::mutex> op(m->op_mu);  // the destination tables are read by a running load
  if (!is_raw(m)) fail(KK_ESTATE, "raw peer attach needs a KK_FANOUT_RAW model");
  if (m->opts.part_count <= 1) fail(KK_ESTATE, "peer attach needs a multi-process (part_count > 1) model");
  if (rank < 0 || rank >= m->opts.part_count || rank == m->opts.part_index) fail(KK_EINVAL, "bad peer rank %d", rank);
  if (m->peer_raw_ptr[rank]) fail(KK_ESTATE, "peer rank %d already attached", rank);
  Device& d = m->ctx->devs[(size_t)m->dev_idx[0]];
  KK_CUDA(cudaSetDevice(d.ordinal));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  void* p = nullptr;
  KK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  m->peer_raw_ptr[rank] = p;
}

void model_export_slice(kk_model* m, void* handle_out, bool as_pointer) {
  if (!is_pull(m) || !m->slice_buf) fail(KK_ESTATE, "this model has no slice buffer (KK_FANOUT_PULL only)");
  if (as_pointer) {
    void* p = m->slice_buf;
    memcpy(handle_out, &p, sizeof p);
    return;
  }
  KK_CUDA(cudaSetDevice(m->ctx->devs[(size_t)m->dev_idx[0]].ordinal));
  cudaIpcMemHandle_t h;
  KK_CUDA(cudaIpcGetMemHandle(&h, m->slice_buf));
  memcpy(handle_out, &h, sizeof h);
}

void model_peer_attach_slice(kk_model* m, int rank, const void* handle, bool is_ipc) {
  // Not under op_mu: stage 1 of a PULL load (kk_load_part) writes only this rank's own pool and slice buffer, so the caller may map the
  // peers' slice buffers on another thread while it runs.  The mapping itself happens outside peer_mu (it is the slow part; several
  // ranks may be attached concurrently), only the table update is locked.
  if (!is_pull(m)) fail(KK_ESTATE, "slice attach needs a KK_FANOUT_PULL model");
  if (rank < 0 || rank >= m->opts.part_count || rank == m->opts.part_index) fail(KK_EINVAL, "bad peer rank %d", rank);
  {
    std::lock_guard<std::mutex> g(m->peer_mu);
    if (m->peer_slice_ptr[rank]) fail(KK_ESTATE, "peer rank %d already attached", rank);
  }
  KK_CUDA(cudaSetDevice(m->ctx->devs[(size_t)m->dev_idx[0]].ordinal));
  void* p = nullptr;
  if (is_ipc) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    KK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  } else {
    memcpy(&p, handle, sizeof p);
    if (!p) fail(KK_EINVAL, "null device pointer");
  }
  std::lock_guard<std::mutex> g(m->peer_mu);
  if (m->peer_slice_ptr[rank]) {  // lost a race against another attach of the same rank
    if (is_ipc) cudaIpcCloseMemHandle(p);
    fail(KK_ESTATE, "peer rank %d already attached", rank);
  }
  m->peer_slice_ptr[rank] = p;
  m->peer_slice_is_ipc[rank] = is_ipc;
}

// NVLink probe (measurement only): copy-engine read of `nbytes` from the attached buffer of rank `rank` (its pool, or its slice buffer
// in a PULL load) into a local scratch, one untimed pass first, CUDA events on the device's stream.  Run by every rank at once against
// its ring neighbour this is the per-GPU ingress rate the fan-out's roofline is quoted against — measured in the run, not assumed.
void model_probe_peer(kk_model* m, int rank, int which, uint64_t& nbytes, float* ms) {
  std::lock_guard<std::mutex> op(m->op_mu);
  if (rank < 0 || rank >= KK_MAX_DEVICES) fail(KK_EINVAL, "bad peer rank %d", rank);
  const uint8_t* src = nullptr;
  uint64_t avail = 0;
  if (which == KK_BUF_SLICE) {
    std::lock_guard<std::mutex> g(m->peer_mu);
    src = (const uint8_t*)m->peer_slice_ptr[rank];
    if (src && (size_t)rank < m->part_range.size()) {
      const auto& pr = m->part_range[(size_t)rank];
      avail = pr.second > (pr.first & ~(uint64_t)255) ? pr.second - (pr.first & ~(uint64_t)255) : 0;
    }
  } else if (which == KK_BUF_POOL) {
    src = (const uint8_t*)m->peer_ptr[rank];
    avail = m->pool_bytes.empty() ? 0 : m->pool_bytes[0];
  } else if (which == KK_BUF_RAW) {
    src = (const uint8_t*)m->peer_raw_ptr[rank];
    avail = m->raw.empty() ? 0 : m->raw[0].bytes;
  } else {
    fail(KK_EINVAL, "probe: buffer kind %d (KK_BUF_POOL, KK_BUF_RAW or KK_BUF_SLICE)", which);
  }
  if (!src) fail(KK_ESTATE, "probe: rank %d has no attached buffer of that kind", rank);
  nbytes = std::min(nbytes, avail) & ~(uint64_t)255;
  if (!nbytes) fail(KK_EINVAL, "probe: nothing to copy from rank %d", rank);
  Device& dev = m->ctx->devs[(size_t)m->dev_idx[0]];
  KK_CUDA(cudaSetDevice(dev.ordinal));
  DevScratch tmp;
  if (cudaMalloc(&tmp.p, nbytes) != cudaSuccess) { cudaGetLastError(); fail(KK_ENOMEM, "probe: cudaMalloc(%llu) failed", (unsigned long long)nbytes); }
  EventSet ev(2);
  ev.create_all();
  KK_CUDA(cudaMemcpyAsync(tmp.p, src, nbytes, cudaMemcpyDeviceToDevice, dev.stream));
  KK_CUDA(cudaEventRecord(ev[0], dev.stream));
  KK_CUDA(cudaMemcpyAsync(tmp.p, src, nbytes, cudaMemcpyDeviceToDevice, dev.stream));
  KK_CUDA(cudaEventRecord(ev[1], dev.stream));
  KK_CUDA(cudaStreamSynchronize(dev.stream));
  KK_CUDA(cudaEventElapsedTime(ms, ev[0], ev[1]));
  if (*ms < 0.f) *ms = 0.f;
}

void model_release(kk_model* m) {
  kk_ctx* c = m->ctx;
  {
    std::lock_guard<std::mutex> g(c->mu);
    if (m->refcount <= 0) fail(KK_ESTATE, "release of a model with refcount %d", m->refcount);
    if (--m->refcount > 0) return;
    c->models.erase(m->key);
  }
  destroy_model(m);
}

void model_peer_attach(kk_model* m, int rank, const void* handle, bool is_ipc) {
  std::lock_guard<std::mutex> op(m->op_mu);  // the destination tables are read by a running load
  if (m->opts.part_count <= 1) fail(KK_ESTATE, "peer attach needs a multi-process (part_count > 1) model");
  if (rank < 0 || rank >= m->opts.part_count || rank == m->opts.part_index) fail(KK_EINVAL, "bad peer rank %d", rank);
  if (m->plan.mode != KK_MODE_BROADCAST && !(m->plan.mode == KK_MODE_SCATTER && (m->plan.flags & KK_LOAD_SCATTER_EXCHANGE)))
    fail(KK_ESTATE, "peer attach only applies to BROADCAST models and to SCATTER models loaded with KK_LOAD_SCATTER_EXCHANGE");
  if (m->peer_ptr[rank]) fail(KK_ESTATE, "peer rank %d already attached", rank);
  Device& d = m->ctx->devs[(size_t)m->dev_idx[0]];
  KK_CUDA(cudaSetDevice(d.ordinal));
  void* p = nullptr;
  if (is_ipc) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    KK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  } else {
    memcpy(&p, handle, sizeof p);
    if (!p) fail(KK_EINVAL, "null device pointer");
  }
  m->peer_ptr[rank] = p;
  m->peer_is_ipc[rank] = is_ipc;
}

void model_peer_detach_all(kk_model* m) {
  std::lock_guard<std::mutex> op(m->op_mu);  // the destination tables are read by a running load
  std::lock_guard<std::mutex> pg(m->peer_mu);
  Device& d = m->ctx->devs[(size_t)m->dev_idx[0]];
  cudaSetDevice(d.ordinal);
  for (int r = 0; r < KK_MAX_DEVICES; ++r) {
    if (m->peer_ptr[r]) {
      if (m->peer_is_ipc[r]) cudaIpcCloseMemHandle(m->peer_ptr[r]);
      m->peer_ptr[r] = nullptr;
    }
    if (m->peer_raw_ptr[r]) {
      cudaIpcCloseMemHandle(m->peer_raw_ptr[r]);
      m->peer_raw_ptr[r] = nullptr;
    }
    if (m->peer_slice_ptr[r]) {
      if (m->peer_slice_is_ipc[r]) cudaIpcCloseMemHandle(m->peer_slice_ptr[r]);
      m->peer_slice_ptr[r] = nullptr;
    }
  }
}

int model_local_device(kk_model* m, int ordinal) {
  for (size_t li = 0; li < m->dev_idx.size(); ++li)
    if (m->ctx->devs[(size_t)m->dev_idx[li]].ordinal == ordinal) return (int)li;
  fail(KK_EINVAL, "device %d holds no pool of this model", ordinal);
}

void model_pool_ipc_handle(kk_model* m, int li, void* handle_out) {
  std::lock_guard<std::mutex> g(m->export_mu);
  if (m->pool_handle_cache.size() < m->pools.size()) m->pool_handle_cache.resize(m->pools.size());
  auto& c = m->pool_handle_cache[(size_t)li];
  if (c.empty()) {
    KK_CUDA(cudaSetDevice(m->ctx->devs[(size_t)m->dev_idx[(size_t)li]].ordinal));
    cudaIpcMemHandle_t h;
    KK_CUDA(cudaIpcGetMemHandle(&h, m->pools[(size_t)li]));
    static_assert(sizeof h == KK_IPC_HANDLE_BYTES, "ipc handle size");
    c.assign((const uint8_t*)&h, (const uint8_t*)&h + sizeof h);
  }
  memcpy(handle_out, c.data(), c.size());
}

static std::string build_manifest(kk_model* m, int li);
std::string model_manifest(kk_model* m, int li) {
  std::lock_guard<std::mutex> g(m->export_mu);
  if (m->manifest_cache.size() < m->dev_idx.size()) m->manifest_cache.resize(m->dev_idx.size());
  auto& c = m->manifest_cache[(size_t)li];
  if (c.empty()) c = build_manifest(m, li);
  return c;
}

static std::string build_manifest(kk_model* m, int li) {
  const auto& pl = m->plan.placement_of_part(m->local_parts[(size_t)li]);
  const auto& T = m->plan.index.tensors;
  std::ostringstream o;
  // "device" is the ordinal in THIS process (diagnostics only); a consumer in another process or container finds the GPU by its UUID / PCI bus id
  const int ordinal = m->ctx->devs[(size_t)m->dev_idx[(size_t)li]].ordinal;
  char bus[32] = "", uuid[48] = "";
  {
    cudaDeviceProp pr;
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof bus, ordinal) != cudaSuccess) { cudaGetLastError(); bus[0] = 0; }
    for (char* c = bus; *c; ++c)
      if (*c >= 'A' && *c <= 'F') *c = (char)(*c - 'A' + 'a');
    if (cudaGetDeviceProperties(&pr, ordinal) == cudaSuccess) {
      const unsigned char* b = (const unsigned char*)pr.uuid.bytes;
      snprintf(uuid, sizeof uuid, "GPU-%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x", b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7],
               b[8], b[9], b[10], b[11], b[12], b[13], b[14], b[15]);
    } else {
      cudaGetLastError();
    }
  }
  o << "{\"apiVersion\":\"kukeon.gpupool/v1\",\"kind\":\"PoolManifest\",\"device\":" << ordinal << ",\"deviceUUID\":\"" << uuid << "\",\"pciBusId\":\"" << bus
    << "\",\"poolBytes\":" << m->pool_bytes[(size_t)li] << ",\"mode\":" << m->plan.mode << ",\"format\":\"" << m->plan.index.format
    << "\",\"align\":" << KK_POOL_ALIGN << ",\"tensors\":[";
  for (size_t i = 0; i < T.size(); ++i) {
    const DtypeInfo* di = dtype_info(pl[i].dtype);
    if (i) o << ",";
    o << "{\"name\":\"" << json_escape(T[i].name) << "\",\"dtype\":\"" << (di ? di->name : "?") << "\",\"shape\":[";
    for (size_t d = 0; d < pl[i].shape.size(); ++d) o << (d ? "," : "") << pl[i].shape[d];
    o << "],\"offset\":" << pl[i].pool_offset << ",\"nbytes\":" << pl[i].nbytes;
    if (pl[i].slice_dim != kNoSlice) o << ",\"sliceDim\":" << pl[i].slice_dim << ",\"sliceBegin\":" << pl[i].slice_begin;
    o << "}";
  }
  o << "]}";
  return o.str();
}

std::string model_stats(kk_model* m) {
  std::ostringstream o;
  o.precision(9);
  uint64_t src = 0, out = 0;
  for (size_t li = 0; li < m->local_parts.size(); ++li) {
    const PartPlan& pp = m->plan.parts[(size_t)m->local_parts[li]];
    src += pp.src_bytes;
    out += pp.out_bytes;
  }
  o << "{\"n_tensors\":" << m->plan.index.tensors.size() << ",\"n_shards\":" << m->plan.index.shards.size()
    << ",\"file_bytes\":" << m->plan.file_bytes << ",\"pool_bytes\":" << m->plan.pool_bytes_of_part(m->local_parts[0])
    << ",\"n_parts\":" << m->plan.n_parts << ",\"local_src_bytes\":" << src << ",\"local_out_bytes\":" << out
    << ",\"index_s\":" << m->t_index << ",\"plan_s\":" << m->t_plan << ",\"alloc_s\":" << m->t_alloc << ",\"load_s\":" << m->t_load
    << ",\"n_loads\":" << m->n_loads << ",\"readers\":{\"threads\":" << m->rd_threads << ",\"slot_wait_s\":" << m->rd_wait_ns.load() / 1e9 << ",\"pread_s\":" << m->rd_pread_ns.load() / 1e9
    << ",\"issue_s\":" << m->rd_issue_ns.load() / 1e9 << ",\"drain_s\":" << m->rd_drain_ns.load() / 1e9 << ",\"files_open_s\":" << m->t_files_open << ",\"files_close_s\":" << m->t_files_close << "},\"load_gbps\":" << (m->t_load > 0 ? (double)src / m->t_load / 1e9 : 0.0) << ",\"parts\":[";
  for (size_t li = 0; li < m->local_parts.size(); ++li) {
    const int part = m->local_parts[li];
    const PartPlan& pp = m->plan.parts[(size_t)part];
    // pool range this part produces (contiguous for SINGLE/BROADCAST because pool order == file order)
    uint64_t lo = UINT64_MAX, hi = 0;
    for (auto& s : pp.segs) {
      uint64_t b = s.dst_off, e;
      switch (s.op) {
        case KK_OP_ROWSPLIT: continue;  // lands in every pool; not part of this rank's contiguous range
        case KK_OP_COPY: e = b + s.units; break;
        case KK_OP_F32_BF16: case KK_OP_F16_BF16: case KK_OP_F8E4M3_BF16: case KK_OP_F8E5M2_BF16: e = b + s.units * 2; break;
        case KK_OP_T_B32: e = b + (uint64_t)s.p0 * s.p1 * 4; break;
        case KK_OP_T_F32_BF16: case KK_OP_T_F16_BF16: case KK_OP_T_B16: e = b + (uint64_t)s.p0 * s.p1 * 2; break;
        default: e = b + s.units * kk_block_geom(s.op).out_bytes; break;  // block-dequantising ops
      }
      if (b < lo) lo = b;
      if (e > hi) hi = e;
    }
    if (lo == UINT64_MAX) lo = 0;
    uint64_t tiles = 0;
    for (auto& ch : pp.chunks) tiles += ch.n_tiles;
    if (li) o << ",";
    o << "{\"part\":" << part << ",\"device\":" << m->ctx->devs[(size_t)m->dev_idx[li]].ordinal << ",\"chunks\":" << pp.chunks.size()
      << ",\"segs\":" << pp.segs.size() << ",\"tiles\":" << tiles << ",\"src_bytes\":" << pp.src_bytes << ",\"out_bytes\":" << pp.out_bytes
      << ",\"pool_lo\":" << lo << ",\"pool_hi\":" << hi << ",\"seconds\":" << (li < m->t_part.size() ? m->t_part[li] : 0.0) << "}";
  }
  o << "]}";
  return o.str();
}

// ---------------------------------------------------------------------------------------------
// resident image: kernel-stage measurement with the source bytes already in HBM
// ---------------------------------------------------------------------------------------------
void model_stage_resident(kk_model* m) {
  std::lock_guard<std::mutex> op(m->op_mu);
  if (is_raw(m)) {  // RAW: the resident image IS the raw image; stage this process's own part(s), no fan-out
    if (m->raw.empty()) fail(KK_ESTATE, "the raw image of this model has been released");
    FdSet fds(m->plan.index.shards);
    for (size_t li = 0; li < m->dev_idx.size(); ++li) run_part_raw(m, (int)li, m->local_parts[li], fds, false);
    return;
  }
  free_resident(m);
  kk_ctx* c = m->ctx;
  FdSet fds(m->plan.index.shards, map_policy(m));
  m->resident.resize(m->dev_idx.size());
  for (size_t li = 0; li < m->dev_idx.size(); ++li) {
    Device& dev = c->devs[(size_t)m->dev_idx[li]];
    KK_CUDA(cudaSetDevice(dev.ordinal));
    const PartPlan& pp = m->plan.parts[(size_t)m->local_parts[li]];
    auto& R = m->resident[li];
    std::vector<uint64_t> img_off(pp.chunks.size());
    uint64_t tot = 0;
    for (size_t i = 0; i < pp.chunks.size(); ++i) {
      img_off[i] = tot;
      tot += align_up(pp.chunks[i].buf_bytes + 64, 256);
    }
    R.image_bytes = tot ? tot : 256;
    cudaError_t e = cudaMalloc((void**)&R.image, R.image_bytes);
    if (e != cudaSuccess) { cudaGetLastError(); R.image = nullptr; fail(KK_ENOMEM, "device %d: cudaMalloc(%llu) for the resident image failed", dev.ordinal, (unsigned long long)R.image_bytes); }
    // copy the chunk bytes exactly as the streaming path stages them
    std::lock_guard<std::mutex> pipeline(*dev.load_mu);
    Reader& rd = dev.readers[0];
    size_t k = 0;
    for (size_t i = 0; i < pp.chunks.size(); ++i) {
      Slot& s = rd.slots[k++ % rd.slots.size()];
      KK_CUDA(cudaEventSynchronize(s.done));
      read_chunk(pp.chunks[i], fds, m->plan.index, s.pinned);
      KK_CUDA(cudaMemcpyAsync(R.image + img_off[i], s.pinned, pp.chunks[i].buf_bytes, cudaMemcpyHostToDevice, rd.stream));
      KK_CUDA(cudaEventRecord(s.done, rd.stream));
    }
    KK_CUDA(cudaStreamSynchronize(rd.stream));
    // one launch per shard (split only if a launch would exceed the segment-table limit)
    std::vector<KKSeg> segs;
    for (size_t i = 0; i < pp.chunks.size(); ++i) {
      const Chunk& ch = pp.chunks[i];
      bool fresh = R.launches.empty() || pp.chunks[i - 1].shard != ch.shard ||
                   R.launches.back().n_segs + ch.seg_count > kMaxSegsPerLaunch ||
                   (uint64_t)R.launches.back().n_tiles + ch.n_tiles > 0xFFFFFFF0ull;
      if (fresh) R.launches.push_back({(uint32_t)segs.size(), 0, 0, 0, 0});
      auto& L = R.launches.back();
      for (uint32_t j = 0; j < ch.seg_count; ++j) {
        KKSeg s = pp.segs[ch.seg_begin + j];
        s.src_off += img_off[i];
        s.tile_begin += L.n_tiles;
        segs.push_back(s);
      }
      L.n_segs += ch.seg_count;
      L.n_tiles += ch.n_tiles;
      L.src_bytes += ch.src_bytes;
      L.out_bytes += ch.out_bytes;
    }
    if (!segs.empty()) {
      KK_CUDA(cudaMalloc((void**)&R.d_segs, segs.size() * sizeof(KKSeg)));
      KK_CUDA(cudaMemcpy(R.d_segs, segs.data(), segs.size() * sizeof(KKSeg), cudaMemcpyHostToDevice));
    }
  }
}

void model_unstage_resident(kk_model* m) {
  std::lock_guard<std::mutex> op(m->op_mu);
  if (is_raw(m)) return;  // the raw image lives as long as a deferred RAW model does
  free_resident(m);
}

// RAW: time the stage-1 fan-out only (own part of the image -> every peer image), one launch per local device.
static void raw_fanout_resident(kk_model* m, float* ms_total, float* ms_per_launch, size_t cap, size_t* n_launches) {
  kk_ctx* c = m->ctx;
  const size_t nl = m->dev_idx.size();
  EventSet e0(nl), e1(nl);
  size_t launched = 0;
  for (size_t li = 0; li < nl; ++li) {
    Device& dev = c->devs[(size_t)m->dev_idx[li]];
    KK_CUDA(cudaSetDevice(dev.ordinal));
    KK_CUDA(cudaEventCreate(&e0[li]));
    KK_CUDA(cudaEventCreate(&e1[li]));
    ConvertLaunch L{};
    fill_raw_dsts(m, (int)li, L);
    const int part = m->local_parts[li];
    const PartPlan& pp = m->plan.parts[(size_t)part];
    KK_CUDA(cudaEventRecord(e0[li], dev.stream));
    if (L.n_dst > 0 && !pp.chunks.empty()) {
      // one launch over all of this part's chunks (their COPY segments are consecutive in d_copy_segs)
      std::vector<KKSeg> segs(pp.chunks.size());
      uint32_t tiles = 0;
      for (size_t ci = 0; ci < pp.chunks.size(); ++ci) {
        KKSeg sg{};
        sg.src_off = sg.dst_off = m->img_off[(size_t)part][ci];
        sg.units = align_up(pp.chunks[ci].buf_bytes, 16);
        sg.op = KK_OP_COPY;
        sg.tile_begin = tiles;
        tiles += (uint32_t)kk_seg_tiles(KK_OP_COPY, sg.units, 0);
        segs[ci] = sg;
      }
      if (segs.size() > kMaxSegsPerLaunch) fail(KK_EUNSUPPORTED, "too many chunks for one raw fan-out launch");
      DevScratch tmp;
      KK_CUDA(cudaMalloc(&tmp.p, segs.size() * sizeof(KKSeg)));
      KKSeg* d_tmp = (KKSeg*)tmp.p;
      KK_CUDA(cudaMemcpyAsync(d_tmp, segs.data(), segs.size() * sizeof(KKSeg), cudaMemcpyHostToDevice, dev.stream));
      KK_CUDA(cudaEventRecord(e0[li], dev.stream));
      L.src = m->raw[li].image;
      L.segs = d_tmp;
      L.n_segs = (uint32_t)segs.size();
      L.n_tiles = tiles;
      L.sched = sched_for(m, dev.sched);
      KK_CUDA(launch_convert(L, dev.sm_count, dev.stream));
      KK_CUDA(cudaEventRecord(e1[li], dev.stream));
      KK_CUDA(cudaStreamSynchronize(dev.stream));  // tmp is freed on scope exit, after the launch that reads it has finished
      launched = 1;
    } else {
      KK_CUDA(cudaEventRecord(e1[li], dev.stream));
    }
  }
  float worst = 0.f;
  for (size_t li = 0; li < nl; ++li) {
    Device& dev = c->devs[(size_t)m->dev_idx[li]];
    KK_CUDA(cudaSetDevice(dev.ordinal));
    KK_CUDA(cudaStreamSynchronize(dev.stream));
    float ms = 0.f;
    KK_CUDA(cudaEventElapsedTime(&ms, e0[li], e1[li]));
    if (ms > worst) worst = ms;
  }
  if (ms_total) *ms_total = worst;
  if (n_launches) *n_launches = launched;
  if (ms_per_launch && cap && launched) ms_per_launch[0] = worst;
}

void model_convert_resident(kk_model* m, float* ms_total, float* ms_per_launch, size_t cap, size_t* n_launches) {
  std::lock_guard<std::mutex> op(m->op_mu);
  if (is_raw(m)) {
    if (m->raw.empty()) fail(KK_ESTATE, "the raw image of this model has been released");
    raw_fanout_resident(m, ms_total, ms_per_launch, cap, n_launches);
    return;
  }
  if (m->resident.size() != m->dev_idx.size()) fail(KK_ESTATE, "kk_stage_resident has not been called");
  kk_ctx* c = m->ctx;
  const size_t nl = m->dev_idx.size();
  std::vector<EventSet> evs;
  evs.reserve(nl);
  size_t max_launches = 0;
  // enqueue on every local device first (they run concurrently), then collect
  for (size_t li = 0; li < nl; ++li) {
    Device& dev = c->devs[(size_t)m->dev_idx[li]];
    auto& R = m->resident[li];
    KK_CUDA(cudaSetDevice(dev.ordinal));
    evs.emplace_back(R.launches.size() + 1);
    evs[li].create_all();
    ConvertLaunch base{};
    fill_dsts(m, (int)li, base);
    for (size_t k = 0; k < R.launches.size(); ++k) {
      KK_CUDA(cudaEventRecord(evs[li][k], dev.stream));
      ConvertLaunch L = base;
      L.src = R.image;
      L.segs = R.d_segs + R.launches[k].seg_begin;
      L.n_segs = R.launches[k].n_segs;
      L.n_tiles = R.launches[k].n_tiles;
      L.sched = sched_for(m, dev.sched);
      KK_CUDA(launch_convert(L, dev.sm_count, dev.stream));
    }
    KK_CUDA(cudaEventRecord(evs[li][R.launches.size()], dev.stream));
    if (R.launches.size() > max_launches) max_launches = R.launches.size();
  }
  float worst = 0.f;
  std::vector<float> per(max_launches, 0.f);
  for (size_t li = 0; li < nl; ++li) {
    Device& dev = c->devs[(size_t)m->dev_idx[li]];
    KK_CUDA(cudaSetDevice(dev.ordinal));
    KK_CUDA(cudaStreamSynchronize(dev.stream));
    const size_t n = evs[li].size() - 1;
    float tot = 0.f;
    if (n) KK_CUDA(cudaEventElapsedTime(&tot, evs[li][0], evs[li][n]));
    if (tot > worst) worst = tot;
    for (size_t k = 0; k < n; ++k) {
      float ms = 0.f;
      KK_CUDA(cudaEventElapsedTime(&ms, evs[li][k], evs[li][k + 1]));
      if (ms > per[k]) per[k] = ms;
    }
  }
  if (ms_total) *ms_total = worst;
  if (n_launches) *n_launches = max_launches;
  if (ms_per_launch)
    for (size_t k = 0; k < max_launches && k < cap; ++k) ms_per_launch[k] = per[k];
}

}  // namespace kk
