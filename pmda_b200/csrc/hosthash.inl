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

#include <thread>
#include <vector>

namespace hosthash {

constexpr uint64_t P1 = 0x9E3779B185EBCA87ull;
constexpr uint64_t P2 = 0xC2B2AE3D27D4EB4Full;
constexpr uint64_t P3 = 0x165667B19E3779F9ull;
constexpr uint64_t P4 = 0x85EBCA77C2B2AE63ull;
constexpr uint64_t P5 = 0x27D4EB2F165667C5ull;
constexpr size_t CHUNK = 256u << 10;

static inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t lane(uint64_t acc, uint64_t w) {
  return rotl(acc + w * P2, 31) * P1;
}
static inline uint64_t avalanche(uint64_t h) {
  h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
  return h;
}

struct Chunker {
  uint64_t a0, a1, a2, a3;
  unsigned char tail[32];
  size_t ntail;
  uint64_t total;
  explicit Chunker(uint64_t seed)
      : a0(seed + P1 + P2), a1(seed + P2), a2(seed), a3(seed - P1), ntail(0),
        total(0) {}
  inline void step(const unsigned char *p) {
    uint64_t w[4];
    memcpy(w, p, 32);
    a0 = lane(a0, w[0]); a1 = lane(a1, w[1]);
    a2 = lane(a2, w[2]); a3 = lane(a3, w[3]);
  }
  void feed(const unsigned char *p, size_t n) {
    total += n;
    if (ntail) {
      const size_t take = n < 32 - ntail ? n : 32 - ntail;
      memcpy(tail + ntail, p, take);
      ntail += take; p += take; n -= take;
      if (ntail < 32) return;
      step(tail);
      ntail = 0;
    }
    for (; n >= 32; p += 32, n -= 32) step(p);
    if (n) { memcpy(tail, p, n); ntail = n; }
  }
  uint64_t finish() {
    if (ntail) {
      memset(tail + ntail, 0, 32 - ntail);
      step(tail);
    }
    uint64_t h = rotl(a0, 1) + rotl(a1, 7) + rotl(a2, 12) + rotl(a3, 18);
    h = (h ^ lane(0, a0)) * P1 + P4;
    h = (h ^ lane(0, a1)) * P1 + P4;
    h = (h ^ lane(0, a2)) * P1 + P4;
    h = (h ^ lane(0, a3)) * P1 + P4;
    return avalanche(h + total * P5);
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
  size_t nt = nthreads > 0 ? (size_t)nthreads : 1;
  if (nt > nchunks) nt = nchunks;
  if (nt > 64) nt = 64;
  if (nt <= 1) {
    work(0, nchunks);
  } else {
    std::vector<std::thread> pool;
    bool failed = false;
    size_t done = 0;
    for (size_t t = 0; t < nt; ++t) {
      const size_t c0 = nchunks * t / nt, c1 = nchunks * (t + 1) / nt;
      try {
        pool.emplace_back(work, c0, c1);
        done = c1;
      } catch (...) {   // no more threads: finish the rest on this one
        failed = true;
        break;
      }
    }
    if (failed) work(done, nchunks);
    for (auto &th : pool) th.join();
  }
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
  auto work = [&](size_t f0, size_t f1) {
    for (size_t f = f0; f < f1; ++f) {
      const float *s = (const float *)(src + (long long)f * frame_stride);
      float *d = dst_host + f * n_idx * 3;
      for (size_t k = 0; k < n_idx; ++k) {
        const float *p = s + (size_t)idx_host[k] * 3;
        d[3 * k + 0] = p[0];
        d[3 * k + 1] = p[1];
        d[3 * k + 2] = p[2];
      }
    }
  };
  size_t nt = nthreads > 0 ? (size_t)nthreads : 1;
  if (nt > n_frames) nt = n_frames;
  if (nt > 64) nt = 64;
  if (nt <= 1 || n_frames * n_idx < (1u << 16)) {
    work(0, n_frames);
    return RDFK_OK;
  }
  std::vector<std::thread> pool;
  size_t done = 0;
  for (size_t t = 0; t < nt; ++t) {
    const size_t f0 = n_frames * t / nt, f1 = n_frames * (t + 1) / nt;
    try {
      pool.emplace_back(work, f0, f1);
      done = f1;
    } catch (...) {
      break;
    }
  }
  if (done < n_frames) work(done, n_frames);
  for (auto &th : pool) th.join();
  return RDFK_OK;
}
