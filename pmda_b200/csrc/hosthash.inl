// hosthash.inl -- host-side content hash of a block of trajectory frames.
//
// The device-resident frame cache (pmda_b200/engine.py) may only serve a block
// of frames if the host data it was uploaded from is unchanged: the reference
// re-reads every frame on every run (pmda/parallel.py:434), so an in-place
// edit of an in-memory trajectory, or a rewritten file, must show up in the
// next run().  A sparse sample cannot guarantee that; this hashes EVERY byte.
//
// Design: the block is `n_rows` rows of `row_bytes` bytes, `row_stride` bytes
// apart (a strided frame selection of one big array is hashed in place).  The
// byte stream (rows concatenated) is cut into fixed 256 KiB chunks; every
// chunk gets a 64-bit hash from four independent multiply-rotate lanes (32
// bytes per step, memory-bandwidth bound), chunks are hashed by `nthreads`
// host threads, and the chunk hashes are folded in order, so the result does
// not depend on the thread count.  Not cryptographic: it detects edits, it
// does not resist an adversary.

#include <pthread.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

// ---------------------------------------------------------------------------
// A small persistent pool for the host-side helpers of this file (hash,
// gather, InterRDF_s epilogue): they run once or several times per run() on
// blocks of a few MB, where creating 8-16 threads per call costs as much as
// the work.  One job at a time; a caller that finds the pool busy (the hash
// of a cached block runs beside the gather of the next batch) simply uses
// threads of its own.
// ---------------------------------------------------------------------------
namespace hostpool {

struct Pool {
  std::mutex job_mutex;              // one parallel_for at a time
  std::mutex m;
  std::condition_variable cv_start, cv_done;
  std::vector<std::thread> workers;
  const std::function<void(size_t)> *fn = nullptr;
  size_t n_tasks = 0, next = 0, running = 0;
  unsigned long long generation = 0;
  bool stop = false;

  void worker() {
    unsigned long long seen = 0;
    std::unique_lock<std::mutex> lk(m);
    for (;;) {
      cv_start.wait(lk, [&] { return stop || generation != seen; });
      if (stop) return;
      seen = generation;
      while (next < n_tasks) {
        const size_t t = next++;
        lk.unlock();
        (*fn)(t);
        lk.lock();
      }
      if (--running == 0) cv_done.notify_all();
    }
  }
  void ensure(size_t n) {
    while (workers.size() < n) {
      try {
        workers.emplace_back([this] { worker(); });
      } catch (...) {
        break;
      }
    }
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv_start.notify_all();
    for (auto &t : workers) t.join();
  }
};

static Pool *g_pool = nullptr;
static std::once_flag g_pool_once;
// after fork() the child has the Pool object but none of its threads: start
// over with a fresh one (the old object is leaked on purpose -- destroying
// std::thread handles of threads that do not exist would terminate)
static void after_fork_in_child() { g_pool = new Pool(); }

static Pool &pool() {
  std::call_once(g_pool_once, [] {
    g_pool = new Pool();   // never destroyed: no join at process exit
    pthread_atfork(nullptr, nullptr, after_fork_in_child);
  });
  return *g_pool;
}

// fn(t) for t in [0, n_tasks) on up to `nthreads` threads (the caller's
// included)
static void parallel_for(size_t n_tasks, int nthreads,
                         const std::function<void(size_t)> &fn) {
  size_t nt = nthreads > 0 ? (size_t)nthreads : 1;
  if (nt > n_tasks) nt = n_tasks;
  if (nt > 64) nt = 64;
  if (nt <= 1) {
    for (size_t t = 0; t < n_tasks; ++t) fn(t);
    return;
  }
  Pool &p = pool();
  std::unique_lock<std::mutex> job(p.job_mutex, std::try_to_lock);
  if (!job.owns_lock()) {
    // pool busy with another caller's job: threads of our own
    std::vector<std::thread> own;
    std::mutex m2;
    size_t next = 0;
    auto run = [&] {
      for (;;) {
        size_t t;
        {
          std::lock_guard<std::mutex> lk(m2);
          if (next >= n_tasks) return;
          t = next++;
        }
        fn(t);
      }
    };
    for (size_t i = 1; i < nt; ++i) {
      try {
        own.emplace_back(run);
      } catch (...) {
        break;
      }
    }
    run();
    for (auto &t : own) t.join();
    return;
  }
  p.ensure(nt - 1);
  const size_t helpers = p.workers.size() < nt - 1 ? p.workers.size() : nt - 1;
  {
    std::lock_guard<std::mutex> lk(p.m);
    p.fn = &fn;
    p.n_tasks = n_tasks;
    p.next = 0;
    p.running = p.workers.size();     // every worker wakes up and reports
    ++p.generation;
  }
  (void)helpers;
  p.cv_start.notify_all();
  {
    std::unique_lock<std::mutex> lk(p.m);
    while (p.next < p.n_tasks) {       // the caller works too
      const size_t t = p.next++;
      lk.unlock();
      fn(t);
      lk.lock();
    }
    p.cv_done.wait(lk, [&] { return p.running == 0; });
    p.fn = nullptr;
  }
}

}  // namespace hostpool

namespace hosthash {

constexpr uint64_t P1 = 0x9E3779B185EBCA87ull;
constexpr uint64_t P2 = 0xC2B2AE3D27D4EB4Full;
constexpr uint64_t P3 = 0x165667B19E3779F9ull;
constexpr uint64_t P4 = 0x85EBCA77C2B2AE63ull;
constexpr uint64_t P5 = 0x27D4EB2F165667C5ull;
constexpr uint64_t P32 = 0x9E3779B1ull;
constexpr size_t CHUNK = 256u << 10;

static inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t avalanche(uint64_t h) {
  h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
  return h;
}

// Eight 64-bit accumulators over 64-byte stripes, in the manner of XXH3's
// long-input loop: acc[i] += lo32(d ^ k) * hi32(d ^ k), acc[i ^ 1] += d.  No
// multiply sits on a loop-carried dependency, so the compiler vectorises the
// stripe (pmuludq) and the loop runs at memory speed; every 1 KiB the
// accumulators are scrambled.  Every input bit reaches the result through
// the additive term and through the product.
struct Chunker {
  uint64_t acc[8], key[8];
  unsigned char tail[64];
  size_t ntail, nstripes;
  uint64_t total;
  explicit Chunker(uint64_t seed) : ntail(0), nstripes(0), total(0) {
    static const uint64_t K[8] = {
        0xbe4ba423396cfeb8ull, 0x1cad21f72c81017cull, 0xdb979083e96dd4deull,
        0x1f67b3b7a4a44072ull, 0x78e5c0cc4ee679cbull, 0x2172ffcc7dd05a82ull,
        0x8e2443f7744608b8ull, 0x4c263a81e69035e0ull};
    for (int i = 0; i < 8; ++i) {
      key[i] = K[i] ^ avalanche(seed + (uint64_t)i * P1);
      acc[i] = (i & 1 ? P2 : P1) + (uint64_t)i * P5;
    }
  }
  inline void stripe(const unsigned char *p) {
    uint64_t d[8];
    memcpy(d, p, 64);
    for (int i = 0; i < 8; ++i) {
      const uint64_t k = d[i] ^ key[i];
      acc[i ^ 1] += d[i];
      acc[i] += (k & 0xffffffffull) * (k >> 32);
    }
    if ((++nstripes & 15) == 0)
      for (int i = 0; i < 8; ++i)
        acc[i] = (acc[i] ^ (acc[i] >> 47) ^ key[i]) * P32;
  }
  void feed(const unsigned char *p, size_t n) {
    total += n;
    if (ntail) {
      const size_t take = n < 64 - ntail ? n : 64 - ntail;
      memcpy(tail + ntail, p, take);
      ntail += take; p += take; n -= take;
      if (ntail < 64) return;
      stripe(tail);
      ntail = 0;
    }
    for (; n >= 64; p += 64, n -= 64) stripe(p);
    if (n) { memcpy(tail, p, n); ntail = n; }
  }
  uint64_t finish() {
    if (ntail) {
      memset(tail + ntail, 0, 64 - ntail);
      stripe(tail);
    }
    uint64_t h = total * P5;
    for (int i = 0; i < 8; ++i)
      h = rotl(h ^ avalanche(acc[i] + key[i ^ 3]), 23) * P1 + P4;
    return avalanche(h);
  }
};

// hash of bytes [lo, hi) of the concatenated row stream
static uint64_t hash_span(const unsigned char *base, size_t row_bytes,
                          long long row_stride, uint64_t lo, uint64_t hi,
                          uint64_t seed) {
  Chunker c(seed);
  uint64_t at = lo;
  while (at < hi) {
    const uint64_t row = at / row_bytes, off = at - row * row_bytes;
    uint64_t n = row_bytes - off;
    if (n > hi - at) n = hi - at;
    c.feed(base + (long long)row * row_stride + (long long)off, (size_t)n);
    at += n;
  }
  return c.finish();
}

}  // namespace hosthash

extern "C" unsigned long long rdfk_host_hash64_rows(const void *base_host,
                                                    size_t n_rows,
                                                    size_t row_bytes,
                                                    long long row_stride,
                                                    unsigned long long seed,
                                                    int nthreads) {
  using namespace hosthash;
  uint64_t h = avalanche(seed ^ P3) ^ ((uint64_t)n_rows * P4) ^
               ((uint64_t)row_bytes * P5);
  if (!base_host || n_rows == 0 || row_bytes == 0) return avalanche(h);
  const unsigned char *base = (const unsigned char *)base_host;
  const uint64_t total = (uint64_t)n_rows * row_bytes;
  const size_t nchunks = (size_t)((total + CHUNK - 1) / CHUNK);
  std::vector<uint64_t> ch(nchunks);
  auto work = [&](size_t c0, size_t c1) {
    for (size_t c = c0; c < c1; ++c) {
      const uint64_t lo = (uint64_t)c * CHUNK;
      const uint64_t hi = lo + CHUNK < total ? lo + CHUNK : total;
      ch[c] = hash_span(base, row_bytes, row_stride, lo, hi, seed + c);
    }
  };
  // tasks of 8 chunks (2 MiB): coarse enough for the pool's bookkeeping
  const size_t per = 8, n_tasks = (nchunks + per - 1) / per;
  hostpool::parallel_for(n_tasks, nthreads, [&](size_t t) {
    const size_t c0 = t * per, c1 = c0 + per < nchunks ? c0 + per : nchunks;
    work(c0, c1);
  });
  for (size_t c = 0; c < nchunks; ++c) h = rotl(h ^ ch[c], 27) * P1 + P4;
  return avalanche(h);
}

// ---------------------------------------------------------------------------
// Host-side gather of the atoms an analysis references, frame by frame, into
// a (page-locked) staging buffer: InterRDF_s touches 64 x (1 + 256) atoms of a
// 50,000-atom frame, and the reference gathers `ag.positions` before anything
// else too (pmda/rdf.py:296-301).  Shipping only those rows cuts the PCIe
// bytes 3x; the gather itself is a DRAM-bound pass over the frames, so it is
// spread over `nthreads` host threads (frames are independent).
//   dst[f][k][0..2] = src[f][idx[k]][0..2]        float32 triples
// `frame_stride` = bytes between consecutive source frames.
// ---------------------------------------------------------------------------
extern "C" int rdfk_host_gather_rows(const void *src_host, size_t n_frames,
                                     long long frame_stride, size_t natoms,
                                     const int *idx_host, size_t n_idx,
                                     float *dst_host, int nthreads) {
  if (n_frames == 0 || n_idx == 0) return RDFK_OK;
  if (!src_host || !idx_host || !dst_host)
    return set_err(RDFK_EINVAL, "null pointer argument%s%s");
  for (size_t k = 0; k < n_idx; ++k)
    if (idx_host[k] < 0 || (size_t)idx_host[k] >= natoms)
      return set_err(RDFK_EINVAL, "atom index out of range%s%s");
  const unsigned char *src = (const unsigned char *)src_host;
  auto frame = [&](size_t f) {
    const float *s = (const float *)(src + (long long)f * frame_stride);
    float *d = dst_host + f * n_idx * 3;
    // (indices are sorted: the hardware prefetchers follow; software
    // prefetching made this loop slower)
    // The destination is a staging buffer nobody reads before the DMA
    // engine does: non-temporal stores keep it out of the caches and spare
    // the read-for-ownership of every destination line (this pass is bound
    // by the host's memory bandwidth, which it shares with the uploads).
    int *di = (int *)d;
    for (size_t k = 0; k < n_idx; ++k) {
      const int *p = (const int *)(s + (size_t)idx_host[k] * 3);
#if defined(__SSE2__)
      __builtin_ia32_movnti(di + 3 * k + 0, p[0]);
      __builtin_ia32_movnti(di + 3 * k + 1, p[1]);
      __builtin_ia32_movnti(di + 3 * k + 2, p[2]);
#else
      di[3 * k + 0] = p[0]; di[3 * k + 1] = p[1]; di[3 * k + 2] = p[2];
#endif
    }
#if defined(__SSE2__)
    __builtin_ia32_sfence();
#endif
  };
  if (n_frames * n_idx < (1u << 16)) nthreads = 1;
  hostpool::parallel_for(n_frames, nthreads, frame);
  return RDFK_OK;
}

// ---------------------------------------------------------------------------
// InterRDF_s epilogue on the host (pmda/rdf.py:311-334), threaded: the uint32
// counts the kernels produced become the float64 tensors the reference
// returns per block (`count[i][a, b, :]`, pmda/rdf.py:294-295), their sum over
// the blocks (`np.sum(self._results[:, 0])`; exact: integers below 2^53) and
//   rdf = count / den[bin],   den = density * vol * n_frames   ([nbins])
// -- one IEEE double division per cell, what numpy does, hence numpy's bits.
// Replaces three numpy passes over [cells, nbins] (astype, sum, divide).
// ---------------------------------------------------------------------------
extern "C" int rdfk_host_sites_finish(const unsigned int *count_host,
                                      int n_blocks, long long n, int nbins,
                                      const double *den_host,
                                      double *count_f64_host,
                                      double *total_host, double *rdf_host,
                                      int nthreads) {
  if (n_blocks < 0 || n < 0 || nbins < 1)
    return set_err(RDFK_EINVAL, "negative size or nbins < 1%s%s");
  if (n == 0 || n_blocks == 0) return RDFK_OK;
  if (!count_host || (rdf_host && !den_host) ||
      (!count_f64_host && !total_host && !rdf_host))
    return set_err(RDFK_EINVAL, "null pointer argument%s%s");
  if (n % nbins) return set_err(RDFK_EINVAL, "n is not a multiple of nbins%s%s");
  const long long cells = n / nbins;
  const long long per = 2048;                    // cells per task
  const size_t n_tasks = (size_t)((cells + per - 1) / per);
  if (n < (1 << 16)) nthreads = 1;
  hostpool::parallel_for(n_tasks, nthreads, [&](size_t t) {
    const long long c0 = (long long)t * per;
    const long long c1 = c0 + per < cells ? c0 + per : cells;
    for (long long c = c0; c < c1; ++c) {
      for (int k = 0; k < nbins; ++k) {
        const long long i = c * nbins + k;
        double sum = 0.0;
        for (int b = 0; b < n_blocks; ++b) {
          const double v = (double)count_host[(size_t)b * n + i];
          if (count_f64_host) count_f64_host[(size_t)b * n + i] = v;
          sum += v;
        }
        if (total_host) total_host[i] = sum;
        if (rdf_host) rdf_host[i] = sum / den_host[k];
      }
    }
  });
  return RDFK_OK;
}
