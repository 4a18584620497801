// rdfk.cu -- hand-written sm_100a kernels + C ABI (include/rdfk.h) for the
// pair-distance histogram behind pmda.rdf.InterRDF / InterRDF_s.
//
// Reference path replaced (per frame): pmda/rdf.py:123-134 and :298-305, i.e.
//   capped_distance(g1.positions, g2.positions, rmax, box) -> np.histogram(dist)
// whose arithmetic is MDAnalysis' distance_array (float32 subtraction, double
// minimum image, double sqrt) followed by numpy's uniform-bin rule
// edges[k] <= d < edges[k+1] (last bin closed).  DESIGN.md section 3 has the
// derivation of the error band used below.
//
// Kernel plan (one rdfk_rdf_hist call == one block of F frames):
//   init_ws_kernel    reset work counter / stats / extents
//   stage_kernel      gather AtomGroup indices, AoS xyz -> SoA x|y|z per frame,
//                     NaN-pad to the tile size, triclinic wrap, per-frame extents
//   tables_kernel     per-frame fp32 decision tables: for every bin k the
//                     interval [lo_k, hi_k) of *fp32 squared distance* for which
//                     the fp64 reference provably lands in bin k
//   pair_tile_kernel  persistent CTAs pull (frame, i-tile, j-tile) work items
//                     from an atomic counter; the j-tile is staged into shared
//                     memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier,
//                     double buffered); each thread keeps RI i-atoms in
//                     registers; fp32 minimum image, no sqrt unless the pair is
//                     inside the cut-off; pairs that fall outside every
//                     [lo_k, hi_k) are re-evaluated in fp64 with the reference's
//                     exact operation order; counts go to warp-private shared
//                     histograms, merged per CTA, then 64-bit global atomics
//   excl_kernel       subtracts the exclusion_block pairs (exact, tiny)
//   sites_kernel      InterRDF_s: one thread per (frame, site cell), fp64 only
//
// No tensor cores: this is not a contraction.  No CPU fallback anywhere.

#include <cuda_runtime.h>
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "rdfk.h"

// ---------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

// number of kernels this library has launched (bench.py's gpu_launches)
static unsigned long long g_launches = 0;
#define COUNT_LAUNCH(n) __atomic_fetch_add(&g_launches, (unsigned long long)(n), __ATOMIC_RELAXED)

static int set_err(int code, const char *fmt, const char *a = "",
                   const char *b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

#define CUDA_TRY(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess)                                                    \
      return set_err(RDFK_ECUDA, "%s failed: %s", #expr,                     \
                     cudaGetErrorString(_e));                                 \
  } while (0)

// ---------------------------------------------------------------------------
// compile-time shape of the tile kernel
// ---------------------------------------------------------------------------
constexpr int TPB = 256;         // threads per CTA (8 warps)
constexpr int NWARP = TPB / 32;
constexpr int RI = 4;            // i-atoms held in registers per thread
constexpr int TILE = TPB * RI;   // atoms per i-tile == atoms per j-tile
constexpr int HIST_SMEM_BUDGET = 96 * 1024;  // bytes for warp-private copies
constexpr float MAGIC = 12582912.0f;         // 1.5 * 2^23: rint via add/sub
constexpr double EPS32 = 5.9604644775390625e-8;  // 2^-24

struct FrameParams {
  float L[3];     // ortho: raw box lengths (reference `box[k]`)
  float inv[3];   // ortho: (float)(1.0 / L[k])  (reference `inverse_box[k]`)
  float fl[3];    // fast path length, 0 when the axis has no PBC
  float finv[3];  // fast path inverse, 0 when the axis has no PBC
  float tri[9];   // triclinic vectors, row-major (a; b; c)
  float tinv[3];  // fast path 1/ax, 1/by, 1/cz
  float r2_cut;   // fp32 r2 >= r2_cut  => reference distance > r1 for sure
  float r2_low;   // fp32 r2 <  r2_low  => reference distance < r0 for sure
  float inv_dr;   // nbins / (r1 - r0)
  float koff;     // -r0 * inv_dr
  int all_slow;   // fast path not provably safe for this frame
};

struct HistParams {
  double r0, r1;        // np.histogram range; r1 is also max_cutoff
  const double *edges;  // [nbins + 1]
  int nbins;
};

struct WsHeader {
  unsigned long long stats[4];  // slow pairs, all-slow frames, items, pairs
  int work_counter;
  int pad[7];
};

struct WsLayout {
  size_t header, extents, fparams, tables, g1, g2, total;
  int np1, np2;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static WsLayout ws_layout(int F, int n1, int n2, int nbins, int same_group) {
  WsLayout w;
  w.np1 = (int)align_up((size_t)(n1 > 0 ? n1 : 1), TILE);
  w.np2 = same_group ? w.np1 : (int)align_up((size_t)(n2 > 0 ? n2 : 1), TILE);
  size_t off = 0;
  w.header = off;
  off = align_up(off + sizeof(WsHeader), 256);
  w.extents = off;
  off = align_up(off + sizeof(int) * 6 * (size_t)F, 256);
  w.fparams = off;
  off = align_up(off + sizeof(FrameParams) * (size_t)F, 256);
  w.tables = off;
  off = align_up(off + sizeof(float2) * (size_t)nbins * (size_t)F, 256);
  w.g1 = off;
  off = align_up(off + sizeof(float) * 3 * (size_t)w.np1 * (size_t)F, 256);
  w.g2 = same_group ? w.g1 : off;
  if (!same_group)
    off = align_up(off + sizeof(float) * 3 * (size_t)w.np2 * (size_t)F, 256);
  w.total = off;
  return w;
}

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ int float_to_ordered(float x) {
  int i = __float_as_int(x);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src,
                                         uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0u;
}
// The wait loop is written in C++ (not as branches inside the asm) so that the
// compiler sees the divergence and re-converges the warp afterwards: with the
// loop hidden in inline PTX, lanes leaving try_wait at different times stayed
// split for the whole tile and every instruction issued twice (ncu: 16.8
// active threads per instruction, profiles/r1_notes.md).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
  __syncwarp();
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---------------------------------------------------------------------------
// fp64 re-check: the reference arithmetic, operation for operation.
//   ortho  : MDAnalysis calc_distances.h _calc_distance_array_ortho + minimum_image
//   tri    : _calc_distance_array_triclinic + minimum_image_triclinic (27 images)
//   none   : _calc_distance_array
// then `d <= max_cutoff` (lib/distances.py _bruteforce_capped) and numpy's
// uniform-bin index with its +-1 edge correction (_histograms_impl.py:836-863).
// Explicit _rn intrinsics: no FMA contraction, same bits as the C oracle built
// with -ffp-contract=off.  Returns the bin or -1.
// ---------------------------------------------------------------------------
template <int BOXKIND>
__device__ __noinline__ int exact_bin(float xi, float yi, float zi, float xj,
                                      float yj, float zj,
                                      const FrameParams *__restrict__ fp,
                                      const HistParams *__restrict__ hp) {
  double dx[3];
  dx[0] = (double)__fsub_rn(xj, xi);
  dx[1] = (double)__fsub_rn(yj, yi);
  dx[2] = (double)__fsub_rn(zj, zi);
  if (BOXKIND == RDFK_BOX_ORTHO) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float L = fp->L[k];
      if (L > FLT_EPSILON) {
        const double s = __dmul_rn((double)fp->inv[k], dx[k]);
        dx[k] = __dmul_rn((double)L, __dsub_rn(s, round(s)));
      }
    }
  } else if (BOXKIND == RDFK_BOX_TRI) {
    const float *b = fp->tri;
    double mn0 = 0.0, mn1 = 0.0, mn2 = 0.0;
    double dsq_min = (double)FLT_MAX;
    for (int ix = -1; ix < 2; ++ix) {
      const double rx = __dadd_rn(dx[0], (double)__fmul_rn(b[0], (float)ix));
      for (int iy = -1; iy < 2; ++iy) {
        const double ry0 = __dadd_rn(rx, (double)__fmul_rn(b[3], (float)iy));
        const double ry1 =
            __dadd_rn(dx[1], (double)__fmul_rn(b[4], (float)iy));
        for (int iz = -1; iz < 2; ++iz) {
          const double rz0 =
              __dadd_rn(ry0, (double)__fmul_rn(b[6], (float)iz));
          const double rz1 =
              __dadd_rn(ry1, (double)__fmul_rn(b[7], (float)iz));
          const double rz2 =
              __dadd_rn(dx[2], (double)__fmul_rn(b[8], (float)iz));
          const double dsq =
              __dadd_rn(__dadd_rn(__dmul_rn(rz0, rz0), __dmul_rn(rz1, rz1)),
                        __dmul_rn(rz2, rz2));
          if (dsq < dsq_min) {
            dsq_min = dsq;
            mn0 = rz0;
            mn1 = rz1;
            mn2 = rz2;
          }
        }
      }
    }
    dx[0] = mn0;
    dx[1] = mn1;
    dx[2] = mn2;
  }
  const double rsq =
      __dadd_rn(__dadd_rn(__dmul_rn(dx[0], dx[0]), __dmul_rn(dx[1], dx[1])),
                __dmul_rn(dx[2], dx[2]));
  const double d = __dsqrt_rn(rsq);
  if (!(d <= hp->r1)) return -1;  // capped_distance: d <= max_cutoff
  // np.histogram keep = (a >= first_edge) & (a <= last_edge)
  if (!(d >= hp->r0)) return -1;
  const int nbins = hp->nbins;
  const double f = __dmul_rn(
      __ddiv_rn(__dsub_rn(d, hp->r0), __dsub_rn(hp->r1, hp->r0)), (double)nbins);
  long long idx = (long long)f;  // astype(np.intp): truncation
  if (idx == nbins) idx -= 1;
  if (d < hp->edges[idx]) idx -= 1;
  if (d >= hp->edges[idx + 1] && idx != nbins - 1) idx += 1;
  return (int)idx;
}

// ---------------------------------------------------------------------------
// fp32 fast path: squared minimum-image distance.  Only has to agree with the
// reference to within the band folded into the decision tables.
// ---------------------------------------------------------------------------
struct FastBox {
  float l0, l1, l2, i0, i1, i2;                  // ortho
  float ax, bx, by, cx, cy, cz, iax, iby, icz;   // tri
};

template <int BOXKIND>
__device__ __forceinline__ float r2_fast(float dx, float dy, float dz,
                                         const FastBox &b) {
  if (BOXKIND == RDFK_BOX_ORTHO) {
    // n = rint(dx / L) through the 1.5*2^23 trick (FRND is a slow-pipe op)
    const float nx = __fadd_rn(__fmaf_rn(dx, b.i0, MAGIC), -MAGIC);
    const float ny = __fadd_rn(__fmaf_rn(dy, b.i1, MAGIC), -MAGIC);
    const float nz = __fadd_rn(__fmaf_rn(dz, b.i2, MAGIC), -MAGIC);
    dx = __fmaf_rn(-b.l0, nx, dx);
    dy = __fmaf_rn(-b.l1, ny, dy);
    dz = __fmaf_rn(-b.l2, nz, dz);
  } else if (BOXKIND == RDFK_BOX_TRI) {
    // sequential reduction into the brick |z|<=cz/2, |y|<=by/2, |x|<=ax/2
    const float nz = __fadd_rn(__fmaf_rn(dz, b.icz, MAGIC), -MAGIC);
    dz = __fmaf_rn(-b.cz, nz, dz);
    dy = __fmaf_rn(-b.cy, nz, dy);
    dx = __fmaf_rn(-b.cx, nz, dx);
    const float ny = __fadd_rn(__fmaf_rn(dy, b.iby, MAGIC), -MAGIC);
    dy = __fmaf_rn(-b.by, ny, dy);
    dx = __fmaf_rn(-b.bx, ny, dx);
    const float nx = __fadd_rn(__fmaf_rn(dx, b.iax, MAGIC), -MAGIC);
    dx = __fmaf_rn(-b.ax, nx, dx);
  }
  return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

__device__ __forceinline__ FastBox load_fastbox(const FrameParams *fp) {
  FastBox b;
  b.l0 = fp->fl[0]; b.l1 = fp->fl[1]; b.l2 = fp->fl[2];
  b.i0 = fp->finv[0]; b.i1 = fp->finv[1]; b.i2 = fp->finv[2];
  b.ax = fp->tri[0]; b.bx = fp->tri[3]; b.by = fp->tri[4];
  b.cx = fp->tri[6]; b.cy = fp->tri[7]; b.cz = fp->tri[8];
  b.iax = fp->tinv[0]; b.iby = fp->tinv[1]; b.icz = fp->tinv[2];
  return b;
}

// ---------------------------------------------------------------------------
// init / stage / tables
// ---------------------------------------------------------------------------
__global__ void init_ws_kernel(WsHeader *h, int *extents, int F) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) {
    h->stats[0] = h->stats[1] = h->stats[2] = h->stats[3] = 0ull;
    h->work_counter = 0;
  }
  if (t < F * 6) extents[t] = (t % 6 < 3) ? INT_MAX : INT_MIN;
}

// One thread per (frame, padded atom slot).  Replaces `ag.positions`
// (gather by AtomGroup.indices) and, for triclinic boxes, the in-place
// _triclinic_pbc wrap of distance_array's float32 copies.
template <int BOXKIND>
__global__ __launch_bounds__(256) void stage_kernel(
    const float *__restrict__ xyz, int natoms, const int *__restrict__ idx,
    int n, int npad, const float *__restrict__ box, float *__restrict__ out,
    int *__restrict__ extents) {
  const int f = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  float x = __int_as_float(0x7fc00000), y = x, z = x;  // NaN padding
  const bool valid = a < n;
  if (valid) {
    const int src = idx ? idx[a] : a;
    const float *p = xyz + ((size_t)f * natoms + src) * 3;
    x = p[0]; y = p[1]; z = p[2];
    if (BOXKIND == RDFK_BOX_TRI) {
      // same operation order as oracle_triclinic_wrap (rdf_oracle.c)
      const float *b = box + (size_t)f * 9;
      const double b0 = b[0], b3 = b[3], b4 = b[4], b6 = b[6], b7 = b[7],
                   b8 = b[8];
      const double bx_by = __ddiv_rn(b3, b4);
      const double cy_cz = __ddiv_rn(b7, b8);
      const double a_zf =
          __ddiv_rn(__dsub_rn(b6, __ddiv_rn(__dmul_rn(b3, b7), b4)), b8);
      double c0 = x, c1 = y, c2 = z;
      bool moved = false;
      if (c2 < 0.0 || c2 >= b8) {
        const double s = floor(__ddiv_rn(c2, b8));
        c2 = __dsub_rn(c2, __dmul_rn(s, b8));
        c1 = __dsub_rn(c1, __dmul_rn(s, b7));
        c0 = __dsub_rn(c0, __dmul_rn(s, b6));
        moved = true;
      }
      double lb = __dmul_rn(c2, cy_cz);
      if (c1 < lb || c1 >= __dadd_rn(lb, b4)) {
        const double s = floor(__ddiv_rn(__dsub_rn(c1, lb), b4));
        c1 = __dsub_rn(c1, __dmul_rn(s, b4));
        c0 = __dsub_rn(c0, __dmul_rn(s, b3));
        moved = true;
      }
      lb = __dadd_rn(__dmul_rn(c1, bx_by), __dmul_rn(c2, a_zf));
      if (c0 < lb || c0 >= __dadd_rn(lb, b0)) {
        const double s = floor(__ddiv_rn(__dsub_rn(c0, lb), b0));
        c0 = __dsub_rn(c0, __dmul_rn(s, b0));
        moved = true;
      }
      if (moved) { x = (float)c0; y = (float)c1; z = (float)c2; }
    }
  }
  if (a < npad) {
    float *o = out + (size_t)f * 3 * npad;
    o[a] = x;
    o[(size_t)npad + a] = y;
    o[2 * (size_t)npad + a] = z;
  }
  // per-frame extents (fminf/fmaxf drop NaN)
  const float big = FLT_MAX;
  float mn[3] = {valid ? x : big, valid ? y : big, valid ? z : big};
  float mx[3] = {valid ? x : -big, valid ? y : -big, valid ? z : -big};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // a NaN / Inf coordinate must poison the extents (-> all_slow frame)
    if (valid && !isfinite(k == 0 ? x : (k == 1 ? y : z))) {
      mn[k] = -big; mx[k] = big;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
    int *e = extents + f * 6;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (mn[k] < big) atomicMin(e + k, float_to_ordered(mn[k]));
      if (mx[k] > -big) atomicMax(e + 3 + k, float_to_ordered(mx[k]));
    }
  }
}

// One CTA per frame.  Builds FrameParams and the [lo_k, hi_k) table.
// Band derivation: DESIGN.md section 3.
__global__ void tables_kernel(int boxkind, const float *__restrict__ box,
                              const int *__restrict__ extents, HistParams hp,
                              FrameParams *__restrict__ fparams,
                              float2 *__restrict__ tables) {
  const int f = blockIdx.x;
  __shared__ double s_delta;
  __shared__ int s_slow;
  FrameParams *fp = fparams + f;
  if (threadIdx.x == 0) {
    const int *e = extents + f * 6;
    double X[3];
    bool slow = false;
    for (int k = 0; k < 3; ++k) {
      const double lo = ordered_to_float(e[k]), hi = ordered_to_float(e[3 + k]);
      X[k] = (hi >= lo) ? (hi - lo) * (1.0 + 1e-6) : 0.0;
      if (!(X[k] < 1e30)) slow = true;
    }
    double dk[3] = {0.0, 0.0, 0.0};
    const float *b = box + (size_t)f * 9;
    for (int k = 0; k < 3; ++k) {
      fp->L[k] = fp->inv[k] = fp->fl[k] = fp->finv[k] = 0.f;
      fp->tinv[k] = 0.f;
    }
    for (int k = 0; k < 9; ++k) fp->tri[k] = 0.f;
    if (boxkind == RDFK_BOX_ORTHO) {
      for (int k = 0; k < 3; ++k) {
        const float L = b[k];
        fp->L[k] = L;
        fp->inv[k] = (float)(1.0 / (double)L);
        if (L > FLT_EPSILON) {
          fp->fl[k] = L;
          fp->finv[k] = fp->inv[k];
          dk[k] = EPS32 * (X[k] + 0.5 * (double)L) * 1.01;
          if (!(X[k] * (double)fp->inv[k] < 2097152.0)) slow = true;
        } else if (!(L <= FLT_EPSILON)) {
          slow = true;  // NaN box length
        }
      }
    } else if (boxkind == RDFK_BOX_TRI) {
      for (int k = 0; k < 9; ++k) fp->tri[k] = b[k];
      const double ax = b[0], bx = b[3], by = b[4], cx = b[6], cy = b[7],
                   cz = b[8];
      const double rmax = hp.r1;
      // fast path validity: the in-range image must be unique, inside the
      // brick, and among the reference's 27 images
      bool ok = ax > 0 && by > 0 && cz > 0 && b[1] == 0.f && b[2] == 0.f &&
                b[5] == 0.f;
      if (ok) {
        const double mn = fmin(ax, fmin(by, cz));
        ok = rmax * 1.001 < 0.5 * mn;
        const double fy = (rmax + 0.5 * fabs(cy)) / by;
        const double fx = (rmax + fy * fabs(bx) + 0.5 * fabs(cx)) / ax;
        ok = ok && fy <= 0.999 && fx <= 0.999;
      }
      if (!ok) slow = true;
      fp->tinv[0] = (float)(1.0 / ax);
      fp->tinv[1] = (float)(1.0 / by);
      fp->tinv[2] = (float)(1.0 / cz);
      dk[0] = 3.0 * EPS32 * (X[0] + fabs(ax) + fabs(bx) + fabs(cx)) * 1.01;
      dk[1] = 3.0 * EPS32 * (X[1] + fabs(by) + fabs(cy)) * 1.01;
      dk[2] = 3.0 * EPS32 * (X[2] + fabs(cz)) * 1.01;
      if (ok) {
        // magic-number rint needs |d / len| < 2^22 at every stage
        if (!((X[2] + cz) / cz < 2097152.0) ||
            !((X[1] + fabs(cy) * (X[2] / cz + 1.0) + by) / by < 2097152.0) ||
            !((X[0] + (fabs(cx) + fabs(bx)) * 1e3 + ax) / ax < 2097152.0))
          slow = true;
      }
    }
    const double delta = sqrt(dk[0] * dk[0] + dk[1] * dk[1] + dk[2] * dk[2]);
    if (!(delta < 1e30)) slow = true;
    s_delta = delta;
    s_slow = slow ? 1 : 0;
    fp->all_slow = slow ? 1 : 0;
    fp->inv_dr = (float)((double)hp.nbins / (hp.r1 - hp.r0));
    fp->koff = (float)(-hp.r0 * ((double)hp.nbins / (hp.r1 - hp.r0)));
  }
  __syncthreads();
  const double delta = s_delta;
  const double gam = 3.5 * EPS32;
  // T_hi(e): fp32 r2 >= T_hi  =>  reference d >= e (and > e)
  // T_lo(e): fp32 r2 <  T_lo  =>  reference d <  e
  auto t_hi = [&](double e) -> float {
    const double dp = delta * 1.001 + 1e-15 * fabs(e) + 1e-300;
    const double v = (e + dp) * (e + dp) * (1.0 + gam);
    return nextafterf(__double2float_ru(v), INFINITY);
  };
  auto t_lo = [&](double e) -> float {
    const double dp = delta * 1.001 + 1e-15 * fabs(e) + 1e-300;
    const double m = e - dp;
    if (!(m > 0.0)) return 0.f;
    const double v = m * m * (1.0 - gam);
    const float r = nextafterf(__double2float_rd(v), 0.f);
    return r > 0.f ? r : 0.f;
  };
  const int nbins = hp.nbins;
  for (int k = threadIdx.x; k < nbins; k += blockDim.x) {
    const double e0 = hp.edges[k], e1 = hp.edges[k + 1];
    float lo = (k == 0 && e0 <= 0.0) ? -INFINITY : t_hi(e0);
    float hi = t_lo(e1);
    if (s_slow) { lo = INFINITY; hi = 0.f; }
    tables[(size_t)f * nbins + k] = make_float2(lo, hi);
  }
  if (threadIdx.x == 0) {
    fp->r2_cut = s_slow ? INFINITY : t_hi(hp.edges[nbins]);
    fp->r2_low = (hp.edges[0] > 0.0 && !s_slow) ? t_lo(hp.edges[0]) : 0.f;
  }
}

// ---------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------
struct PairArgs {
  const float *g1;  // staged [F][3][np1]
  const float *g2;  // staged [F][3][np2]
  int np1, np2, n1, n2, F;
  int nt1, nt2;          // tiles per group
  int symmetric;         // circulant half-matrix schedule over nt1 == nt2
  int items_per_frame;
  int total_items;
  const FrameParams *fparams;
  const float2 *tables;
  HistParams hp;
  int hist_copies;       // warp-private shared copies (0 = global atomics)
  unsigned int flush_every;  // items between 32-bit shared-counter flushes
  unsigned long long *hist;
  WsHeader *header;
};

// ---------------------------------------------------------------------------
// In-range pairs are a minority (0.1-16 % of evaluated pairs) scattered over
// the lanes of a warp, so binning them where they are found wastes the warp on
// divergence (measured: 2.8x slower, tools/ubench2.cu).  Instead every thread
// appends its hits to a private queue in shared memory with three predicated
// instructions (no branch), and the warp drains all queues at warp-uniform
// points.  A queue entry is (fp32 r2, j-step); the i-atom is implied by the
// owning thread.  The rare band failure recovers the exact (u, r) pair by
// re-evaluating the 4*RI candidates of that j-step (resolve_band).
// ---------------------------------------------------------------------------
constexpr int QS = 16;                 // queue slots per thread
constexpr int Q_STRIDE = TPB * 4;      // bytes between slots of one thread
constexpr int Q_JOFF = QS * TPB * 4;   // byte offset of the j-step array
constexpr int Q_CHECK = 2;             // drain check every Q_CHECK u-steps
static_assert(QS >= 2 * Q_CHECK * RI, "queue must absorb one check interval");

__device__ __forceinline__ void q_push(uint32_t &qaddr, float r2, float cut,
                                       uint32_t jb) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.lt.f32 p, %1, %2;\n"
      "@p st.shared.f32 [%0], %1;\n"
      "@p st.shared.u32 [%0+%4], %3;\n"
      "@p add.u32 %0, %0, %5;\n"
      "}\n"
      : "+r"(qaddr)
      : "f"(r2), "f"(cut), "r"(jb), "n"(Q_JOFF), "n"(Q_STRIDE)
      : "memory");
}

// Rare path: the `rank`-th (0-based, in (u, r) scan order) pair of j-step `jb`
// whose fp32 r2 equals `r2` bit for bit; returns its exact bin (or -1).
template <int BOXKIND>
__device__ __noinline__ int resolve_band(
    float r2, uint32_t jb, int rank, const float *xs, float x0, float x1,
    float x2, float x3, float y0, float y1, float y2, float y3, float z0,
    float z1, float z2, float z3, const FrameParams *__restrict__ fp,
    const HistParams *__restrict__ hp) {
  const FastBox fb = load_fastbox(fp);
  const float xi[4] = {x0, x1, x2, x3}, yi[4] = {y0, y1, y2, y3},
              zi[4] = {z0, z1, z2, z3};
  const float *ys = xs + TILE, *zs = xs + 2 * TILE;
  int seen = 0;
  for (int u = 0; u < 4; ++u) {
    const float xj = xs[jb + u], yj = ys[jb + u], zj = zs[jb + u];
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      const float v = r2_fast<BOXKIND>(__fsub_rn(xj, xi[r]), __fsub_rn(yj, yi[r]),
                                       __fsub_rn(zj, zi[r]), fb);
      if (__float_as_uint(v) == __float_as_uint(r2)) {
        if (seen == rank)
          return exact_bin<BOXKIND>(xi[r], yi[r], zi[r], xj, yj, zj, fp, hp);
        ++seen;
      }
    }
  }
  return -1;  // unreachable: the entry was produced by one of these pairs
}

template <int BOXKIND>
__global__ __launch_bounds__(TPB, 2) void pair_tile_kernel(const PairArgs a) {
  static_assert(RI == 4, "resolve_band takes four i-atoms");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *jbuf = reinterpret_cast<float *>(smem_raw);            // [2][3][TILE]
  float *q_r2 = jbuf + 2 * 3 * TILE;                            // [QS][TPB]
  uint32_t *q_j = reinterpret_cast<uint32_t *>(q_r2 + QS * TPB);  // [QS][TPB]
  unsigned int *hist_s = reinterpret_cast<unsigned int *>(q_j + QS * TPB);
  __shared__ uint64_t bars[2];
  __shared__ int s_next[2];
  __shared__ HistParams s_hp;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int nbins = a.hp.nbins;
  const int copies = a.hist_copies;
  unsigned int *wh = copies ? hist_s + (size_t)(warp % copies) * nbins : nullptr;
  const uint32_t qbase = smem_u32(q_r2 + tid);
  uint32_t qaddr = qbase;

  for (int b = tid; b < copies * nbins; b += TPB) hist_s[b] = 0u;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_hp = a.hp;
  }
  __syncthreads();

  auto decode = [&](int item, int &f, int &it, int &jt) {
    f = item / a.items_per_frame;
    const int w = item - f * a.items_per_frame;
    if (a.symmetric) {
      const int nt = a.nt1, B = (nt - 1) / 2 + 1;
      if (w < nt * B) { it = w / B; jt = it + (w - it * B); }
      else { it = w - nt * B; jt = it + nt / 2; }
      if (jt >= nt) jt -= nt;
    } else {
      it = w / a.nt2; jt = w - it * a.nt2;
    }
  };
  auto issue_load = [&](int item, int buf) {
    // thread 0 only: j-tile of `item` -> jbuf[buf] by three 1-D TMA copies
    int f, it, jt;
    decode(item, f, it, jt);
    const float *src = a.g2 + (size_t)f * 3 * a.np2 + (size_t)jt * TILE;
    float *dst = jbuf + (size_t)buf * 3 * TILE;
    mbar_expect_tx(&bars[buf], 3u * TILE * sizeof(float));
    bulk_g2s(dst, src, TILE * sizeof(float), &bars[buf]);
    bulk_g2s(dst + TILE, src + a.np2, TILE * sizeof(float), &bars[buf]);
    bulk_g2s(dst + 2 * TILE, src + 2 * (size_t)a.np2, TILE * sizeof(float),
             &bars[buf]);
  };

  // prologue: grab the first item and start its load
  if (tid == 0) {
    const int first = atomicAdd(&a.header->work_counter, 1);
    s_next[0] = first;
    if (first < a.total_items) issue_load(first, 0);
  }
  __syncthreads();

  unsigned int n_slow = 0;
  unsigned int n_items = 0;
  uint32_t phase0 = 0u, phase1 = 0u;
  int buf = 0;
  int cur = s_next[0];

  while (cur < a.total_items) {
    // fetch the next item and prefetch its j-tile into the other buffer (all
    // threads left that buffer at the __syncthreads closing the last round)
    if (tid == 0) {
      const int nxt = atomicAdd(&a.header->work_counter, 1);
      s_next[buf ^ 1] = nxt;
      if (nxt < a.total_items) issue_load(nxt, buf ^ 1);
    }

    int f, it, jt;
    decode(cur, f, it, jt);
    const unsigned int weight = (a.symmetric && it != jt) ? 2u : 1u;
    const FrameParams *fp = a.fparams + f;
    const float2 *tab = a.tables + (size_t)f * nbins;
    const FastBox fb = load_fastbox(fp);
    const float r2_cut = fp->r2_cut;
    const int all_slow = fp->all_slow;

    // my i-atoms (coalesced; NaN in the padded tail)
    float xi[RI], yi[RI], zi[RI];
    {
      const float *g = a.g1 + (size_t)f * 3 * a.np1 + (size_t)it * TILE;
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        xi[r] = g[r * TPB + tid];
        yi[r] = g[(size_t)a.np1 + r * TPB + tid];
        zi[r] = g[2 * (size_t)a.np1 + r * TPB + tid];
      }
    }
    const int jcount = min(TILE, a.n2 - jt * TILE);
    const float *xs = jbuf + (size_t)buf * 3 * TILE;
    const float *ys = xs + TILE;
    const float *zs = xs + 2 * TILE;

    if (buf == 0) { mbar_wait(&bars[0], phase0); phase0 ^= 1u; }
    else { mbar_wait(&bars[1], phase1); phase1 ^= 1u; }

    // bin everything queued so far (warp-uniform call sites only).  The loop
    // bound is warp-uniform (REDUX max) and the per-lane work is a structured
    // `if`: a per-lane trip count left the lanes split into sub-warps for the
    // rest of the tile (BSYNC did not merge them; ncu showed 16.8 active
    // threads per instruction in the hot loop).
    auto drain = [&]() {
      const unsigned int n = (qaddr - qbase) / Q_STRIDE;
      const unsigned int nmax = __reduce_max_sync(0xffffffffu, n);
      const float r2_low = fp->r2_low, inv_dr = fp->inv_dr, koff = fp->koff;
      for (unsigned int s = 0; s < nmax; ++s) {
        if (s < n) {
          const float r2 = q_r2[s * TPB + tid];
          int k = __float2int_rd(__fmaf_rn(sqrt_approx(r2), inv_dr, koff));
          k = max(0, min(k, nbins - 1));
          const float2 t = __ldg(tab + k);
          bool ok = r2 >= r2_low;
          if (ok && !(r2 >= t.x && r2 < t.y)) {
            const uint32_t jb = q_j[s * TPB + tid];
            int rank = 0;
            for (unsigned int e = 0; e < s; ++e)
              rank += (q_j[e * TPB + tid] == jb &&
                       __float_as_uint(q_r2[e * TPB + tid]) ==
                           __float_as_uint(r2));
            k = resolve_band<BOXKIND>(r2, jb, rank, xs, xi[0], xi[1], xi[2],
                                      xi[3], yi[0], yi[1], yi[2], yi[3], zi[0],
                                      zi[1], zi[2], zi[3], fp, &s_hp);
            ++n_slow;
            ok = k >= 0;
          }
          if (ok) {
            if (wh) atomicAdd(wh + k, weight);
            else atomicAdd(a.hist + k, (unsigned long long)weight);
          }
        }
      }
      qaddr = qbase;
    };

    if (!all_slow) {
#pragma unroll 1
      for (int j = 0; j < jcount; j += 4) {
        const float4 X = *reinterpret_cast<const float4 *>(xs + j);
        const float4 Y = *reinterpret_cast<const float4 *>(ys + j);
        const float4 Z = *reinterpret_cast<const float4 *>(zs + j);
        const float xj4[4] = {X.x, X.y, X.z, X.w};
        const float yj4[4] = {Y.x, Y.y, Y.z, Y.w};
        const float zj4[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int r = 0; r < RI; ++r) {
            const float r2 = r2_fast<BOXKIND>(__fsub_rn(xj4[u], xi[r]),
                                              __fsub_rn(yj4[u], yi[r]),
                                              __fsub_rn(zj4[u], zi[r]), fb);
            q_push(qaddr, r2, r2_cut, (uint32_t)j);
          }
          if ((u % Q_CHECK) == Q_CHECK - 1) {
            // room for the next Q_CHECK*RI pushes in every lane?
            if (__any_sync(0xffffffffu,
                           qaddr > qbase + (QS - Q_CHECK * RI) * Q_STRIDE)) {
              drain();
            }
          }
        }
      }
      drain();
    } else {
      // frame not provably safe for fp32: every pair through the fp64 path
      for (int j = 0; j < jcount; ++j) {
        const float xj = xs[j], yj = ys[j], zj = zs[j];
#pragma unroll 1
        for (int r = 0; r < RI; ++r) {
          if (it * TILE + r * TPB + tid >= a.n1) continue;
          const int k = exact_bin<BOXKIND>(xi[r], yi[r], zi[r], xj, yj, zj, fp,
                                           &s_hp);
          ++n_slow;
          if (k >= 0) {
            if (wh) atomicAdd(wh + k, weight);
            else atomicAdd(a.hist + k, (unsigned long long)weight);
          }
        }
      }
    }
    ++n_items;
    // shared counters are 32-bit: flush long before they can wrap
    if (copies && (n_items % a.flush_every) == 0u) {
      __syncthreads();
      for (int b = tid; b < nbins; b += TPB) {
        unsigned long long s = 0;
        for (int c = 0; c < copies; ++c) {
          s += hist_s[(size_t)c * nbins + b];
          hist_s[(size_t)c * nbins + b] = 0u;
        }
        if (s) atomicAdd(a.hist + b, s);
      }
    }
    __syncthreads();  // everyone is done with jbuf[buf] and saw s_next
    buf ^= 1;
    cur = s_next[buf];
  }

  __syncthreads();
  if (copies) {
    for (int b = tid; b < nbins; b += TPB) {
      unsigned long long s = 0;
      for (int c = 0; c < copies; ++c) s += hist_s[(size_t)c * nbins + b];
      if (s) atomicAdd(a.hist + b, s);
    }
  }
  // stats (diagnostics only)
  unsigned int tot = n_slow;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  if ((tid & 31) == 0 && tot) atomicAdd(&a.header->stats[0], (unsigned long long)tot);
  if (tid == 0 && n_items) {
    atomicAdd(&a.header->stats[2], (unsigned long long)n_items);
    atomicAdd(&a.header->stats[3],
              (unsigned long long)n_items * TILE * (unsigned long long)TILE);
  }
}

// ---------------------------------------------------------------------------
// exclusion_block: subtract the pairs with i / xa == j / xb (pmda/rdf.py:129-133)
// one thread per (frame, i); exact arithmetic only (N1 * xb pairs per frame).
// ---------------------------------------------------------------------------
template <int BOXKIND>
__global__ __launch_bounds__(256) void excl_kernel(
    const float *__restrict__ g1, const float *__restrict__ g2, int np1,
    int np2, int n1, int n2, int xa, int xb,
    const FrameParams *__restrict__ fparams, HistParams hp,
    unsigned long long *__restrict__ hist) {
  extern __shared__ unsigned int sh[];
  __shared__ HistParams s_hp;
  const int f = blockIdx.y;
  const int nbins = hp.nbins;
  for (int b = threadIdx.x; b < nbins; b += blockDim.x) sh[b] = 0u;
  if (threadIdx.x == 0) s_hp = hp;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n1) {
    const float *p1 = g1 + (size_t)f * 3 * np1;
    const float *p2 = g2 + (size_t)f * 3 * np2;
    const float xi = p1[i], yi = p1[(size_t)np1 + i], zi = p1[2 * (size_t)np1 + i];
    const long long blk = i / xa;
    const long long j0 = blk * xb;
    const long long j1 = min((long long)n2, j0 + xb);
    for (long long j = j0; j < j1; ++j) {
      const int k = exact_bin<BOXKIND>(xi, yi, zi, p2[j], p2[(size_t)np2 + j],
                                       p2[2 * (size_t)np2 + j], fparams + f,
                                       &s_hp);
      if (k >= 0) atomicAdd(sh + k, 1u);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += blockDim.x)
    if (sh[b]) atomicAdd(hist + b, 0ull - (unsigned long long)sh[b]);
}

// ---------------------------------------------------------------------------
// InterRDF_s: one thread per (frame, cell).  16k pairs per frame at the
// BASELINE config: latency/output bound, so fp64 reference arithmetic only.
// ---------------------------------------------------------------------------
template <int BOXKIND>
__global__ __launch_bounds__(256) void sites_kernel(
    const float *__restrict__ xyz, int natoms, const int *__restrict__ idxA,
    const int *__restrict__ idxB, const int *__restrict__ offA,
    const int *__restrict__ offB, const long long *__restrict__ cell_off,
    int n_sites, long long total_cells, const float *__restrict__ box,
    HistParams hp, unsigned int *__restrict__ count) {
  __shared__ HistParams s_hp;
  __shared__ FrameParams s_fp;
  const int f = blockIdx.y;
  if (threadIdx.x == 0) {
    s_hp = hp;
    FrameParams p;
    memset(&p, 0, sizeof(p));
    if (BOXKIND == RDFK_BOX_ORTHO) {
      for (int k = 0; k < 3; ++k) {
        p.L[k] = box[(size_t)f * 9 + k];
        p.inv[k] = (float)(1.0 / (double)p.L[k]);
      }
    } else if (BOXKIND == RDFK_BOX_TRI) {
      for (int k = 0; k < 9; ++k) p.tri[k] = box[(size_t)f * 9 + k];
    }
    s_fp = p;
  }
  __syncthreads();
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= total_cells) return;
  // site pair of this cell: binary search in the prefix sums
  int lo = 0, hi = n_sites;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cell_off[mid] <= cell) lo = mid; else hi = mid;
  }
  const int s = lo;
  const long long local = cell - cell_off[s];
  const int n2 = offB[s + 1] - offB[s];
  const int ia = (int)(local / n2), ib = (int)(local - (long long)ia * n2);
  const float *pa = xyz + ((size_t)f * natoms + idxA[offA[s] + ia]) * 3;
  const float *pb = xyz + ((size_t)f * natoms + idxB[offB[s] + ib]) * 3;
  float xa = pa[0], ya = pa[1], za = pa[2];
  float xb = pb[0], yb = pb[1], zb = pb[2];
  if (BOXKIND == RDFK_BOX_TRI) {
    // per-atom wrap, same operation order as stage_kernel / the oracle
    const float *b = s_fp.tri;
    const double b0 = b[0], b3 = b[3], b4 = b[4], b6 = b[6], b7 = b[7], b8 = b[8];
    const double bx_by = __ddiv_rn(b3, b4);
    const double cy_cz = __ddiv_rn(b7, b8);
    const double a_zf =
        __ddiv_rn(__dsub_rn(b6, __ddiv_rn(__dmul_rn(b3, b7), b4)), b8);
    for (int which = 0; which < 2; ++which) {
      double c0 = which ? xb : xa, c1 = which ? yb : ya, c2 = which ? zb : za;
      bool moved = false;
      if (c2 < 0.0 || c2 >= b8) {
        const double sft = floor(__ddiv_rn(c2, b8));
        c2 = __dsub_rn(c2, __dmul_rn(sft, b8));
        c1 = __dsub_rn(c1, __dmul_rn(sft, b7));
        c0 = __dsub_rn(c0, __dmul_rn(sft, b6));
        moved = true;
      }
      double lb = __dmul_rn(c2, cy_cz);
      if (c1 < lb || c1 >= __dadd_rn(lb, b4)) {
        const double sft = floor(__ddiv_rn(__dsub_rn(c1, lb), b4));
        c1 = __dsub_rn(c1, __dmul_rn(sft, b4));
        c0 = __dsub_rn(c0, __dmul_rn(sft, b3));
        moved = true;
      }
      lb = __dadd_rn(__dmul_rn(c1, bx_by), __dmul_rn(c2, a_zf));
      if (c0 < lb || c0 >= __dadd_rn(lb, b0)) {
        const double sft = floor(__ddiv_rn(__dsub_rn(c0, lb), b0));
        c0 = __dsub_rn(c0, __dmul_rn(sft, b0));
        moved = true;
      }
      if (moved) {
        if (which) { xb = (float)c0; yb = (float)c1; zb = (float)c2; }
        else { xa = (float)c0; ya = (float)c1; za = (float)c2; }
      }
    }
  }
  const int k = exact_bin<BOXKIND>(xa, ya, za, xb, yb, zb, &s_fp, &s_hp);
  if (k >= 0) atomicAdd(count + (size_t)cell * hp.nbins + k, 1u);
}

// ---------------------------------------------------------------------------
// FP32 peak probe: 8 independent FFMA chains per thread, no memory traffic.
// ---------------------------------------------------------------------------
__global__ __launch_bounds__(256) void ffma_peak_kernel(float *out, int iters,
                                                        float a, float b) {
  float v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4,
        v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      v0 = __fmaf_rn(v0, a, b); v1 = __fmaf_rn(v1, a, b);
      v2 = __fmaf_rn(v2, a, b); v3 = __fmaf_rn(v3, a, b);
      v4 = __fmaf_rn(v4, a, b); v5 = __fmaf_rn(v5, a, b);
      v6 = __fmaf_rn(v6, a, b); v7 = __fmaf_rn(v7, a, b);
    }
  }
  const float s = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
  if (s == 123.456f) out[0] = s;
}

// ---------------------------------------------------------------------------
// host side of the C ABI
// ---------------------------------------------------------------------------
static int check_device(int *sm_count) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  int major = 0, sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (major != 10)
    return set_err(RDFK_EARCH, "librdfk is built for sm_100a only%s%s");
  *sm_count = sms;
  return RDFK_OK;
}

extern "C" int rdfk_version(void) { return RDFK_VERSION; }
extern "C" const char *rdfk_last_error(void) { return g_err; }

extern "C" int rdfk_device_info(int device, int *sm_count, int *cc_major,
                                int *cc_minor, size_t *smem_optin_bytes) {
  int v = 0;
  if (sm_count) {
    CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    *sm_count = v;
  }
  if (cc_major) {
    CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
    *cc_major = v;
  }
  if (cc_minor) {
    CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
    *cc_minor = v;
  }
  if (smem_optin_bytes) {
    CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    *smem_optin_bytes = (size_t)v;
  }
  return RDFK_OK;
}

extern "C" size_t rdfk_rdf_workspace_bytes(int F, int n1, int n2, int nbins,
                                           int same_group) {
  if (F < 0 || n1 < 0 || n2 < 0 || nbins < 1) return 0;
  return ws_layout(F, n1, n2, nbins, same_group).total;
}

template <int BOXKIND>
static int launch_rdf(const float *xyz, int F, int natoms, const int *idx1,
                      int n1, const int *idx2, int n2, int same_group,
                      const float *box, const HistParams &hp, int excl_a,
                      int excl_b, unsigned long long *hist, char *ws,
                      const WsLayout &L, int sms, cudaStream_t st) {
  WsHeader *header = reinterpret_cast<WsHeader *>(ws + L.header);
  int *extents = reinterpret_cast<int *>(ws + L.extents);
  FrameParams *fparams = reinterpret_cast<FrameParams *>(ws + L.fparams);
  float2 *tables = reinterpret_cast<float2 *>(ws + L.tables);
  float *g1 = reinterpret_cast<float *>(ws + L.g1);
  float *g2 = reinterpret_cast<float *>(ws + L.g2);

  init_ws_kernel<<<(F * 6 + 255) / 256 + 1, 256, 0, st>>>(header, extents, F);
  COUNT_LAUNCH(same_group ? 3 : 4);  // init + stage(s) + tables
  stage_kernel<BOXKIND><<<dim3(L.np1 / 256, F), 256, 0, st>>>(
      xyz, natoms, idx1, n1, L.np1, box, g1, extents);
  if (!same_group)
    stage_kernel<BOXKIND><<<dim3(L.np2 / 256, F), 256, 0, st>>>(
        xyz, natoms, idx2, n2, L.np2, box, g2, extents);
  tables_kernel<<<F, 256, 0, st>>>(BOXKIND, box, extents, hp, fparams, tables);

  PairArgs a;
  a.g1 = g1; a.g2 = g2;
  a.np1 = L.np1; a.np2 = L.np2; a.n1 = n1; a.n2 = n2; a.F = F;
  a.nt1 = L.np1 / TILE; a.nt2 = L.np2 / TILE;
  a.symmetric = same_group ? 1 : 0;
  if (a.symmetric) {
    const long long nt = a.nt1, B = (nt - 1) / 2 + 1;
    a.items_per_frame = (int)(nt * B + ((nt % 2 == 0) ? nt / 2 : 0));
  } else {
    a.items_per_frame = a.nt1 * a.nt2;
  }
  const long long total = (long long)a.items_per_frame * F;
  if (total > 0x7fff0000LL)
    return set_err(RDFK_EINVAL, "too many tile work items in one call; "
                                "pass fewer frames%s%s");
  a.total_items = (int)total;
  a.fparams = fparams; a.tables = tables; a.hp = hp; a.hist = hist;
  a.header = header;
  const size_t per_copy = (size_t)hp.nbins * sizeof(unsigned int);
  int copies = (int)(HIST_SMEM_BUDGET / per_copy);
  if (copies > NWARP) copies = NWARP;
  a.hist_copies = copies;
  {
    // one item adds at most 32*RI*TILE*2 to a warp's copy; several warps may
    // share a copy when nbins is large
    const unsigned int sharers = copies ? (NWARP + copies - 1) / copies : 1;
    a.flush_every = (1u << 31) / (32u * RI * TILE * 2u * sharers);
    if (a.flush_every < 1u) a.flush_every = 1u;
  }
  const size_t smem = 2 * 3 * TILE * sizeof(float) + 2 * (size_t)QS * TPB * 4 +
                      (size_t)copies * per_copy;
  CUDA_TRY(cudaFuncSetAttribute(pair_tile_kernel<BOXKIND>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  int occ = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &occ, pair_tile_kernel<BOXKIND>, TPB, smem));
  if (occ < 1) return set_err(RDFK_ECUDA, "pair_tile_kernel does not fit%s%s");
  long long grid = (long long)sms * occ;
  if (grid > total) grid = total;
  if (grid > 0) {
    pair_tile_kernel<BOXKIND><<<(int)grid, TPB, smem, st>>>(a);
    COUNT_LAUNCH(1);
  }

  if (excl_a > 0 && excl_b > 0) {
    const size_t sm = (size_t)hp.nbins * sizeof(unsigned int);
    if (sm > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(excl_kernel<BOXKIND>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sm));
    excl_kernel<BOXKIND><<<dim3((n1 + 255) / 256, F), 256, sm, st>>>(
        g1, g2, L.np1, L.np2, n1, n2, excl_a, excl_b, fparams, hp, hist);
    COUNT_LAUNCH(1);
  }
  CUDA_TRY(cudaGetLastError());
  return RDFK_OK;
}

extern "C" int rdfk_rdf_hist(const float *xyz, int F, int natoms,
                             const int *idx1, int n1, const int *idx2, int n2,
                             int same_group, int boxkind, const float *box,
                             double r0, double r1, int nbins,
                             const double *edges, int excl_a, int excl_b,
                             unsigned long long *hist, void *workspace,
                             size_t workspace_bytes, void *stream) {
  if (F < 0 || natoms < 0 || n1 < 0 || n2 < 0 || nbins < 1)
    return set_err(RDFK_EINVAL, "negative size or nbins < 1%s%s");
  if (!(r1 > r0)) return set_err(RDFK_EINVAL, "need r1 > r0%s%s");
  if (boxkind < RDFK_BOX_NONE || boxkind > RDFK_BOX_TRI)
    return set_err(RDFK_EINVAL, "bad boxkind%s%s");
  if ((excl_a > 0) != (excl_b > 0) || excl_a < 0 || excl_b < 0)
    return set_err(RDFK_EINVAL, "exclusion block needs two positive sizes%s%s");
  if ((size_t)nbins * sizeof(unsigned int) > 200 * 1024)
    return set_err(RDFK_EINVAL, "nbins too large (max 51200)%s%s");
  if (F == 0 || n1 == 0 || n2 == 0) return RDFK_OK;
  if (!xyz || !edges || !hist || !workspace ||
      (boxkind != RDFK_BOX_NONE && !box))
    return set_err(RDFK_EINVAL, "null pointer argument%s%s");
  if (same_group && n1 != n2)
    return set_err(RDFK_EINVAL, "same_group requires n1 == n2%s%s");
  const WsLayout L = ws_layout(F, n1, n2, nbins, same_group);
  if (workspace_bytes < L.total || ((uintptr_t)workspace & 255u))
    return set_err(RDFK_EWORKSPACE, "workspace too small or not 256-byte "
                                    "aligned%s%s");
  int sms = 0;
  int rc = check_device(&sms);
  if (rc) return rc;
  HistParams hp;
  hp.r0 = r0; hp.r1 = r1; hp.edges = edges; hp.nbins = nbins;
  cudaStream_t st = (cudaStream_t)stream;
  char *ws = (char *)workspace;
  switch (boxkind) {
    case RDFK_BOX_NONE:
      return launch_rdf<RDFK_BOX_NONE>(xyz, F, natoms, idx1, n1, idx2, n2,
                                       same_group, box, hp, excl_a, excl_b,
                                       hist, ws, L, sms, st);
    case RDFK_BOX_ORTHO:
      return launch_rdf<RDFK_BOX_ORTHO>(xyz, F, natoms, idx1, n1, idx2, n2,
                                        same_group, box, hp, excl_a, excl_b,
                                        hist, ws, L, sms, st);
    default:
      return launch_rdf<RDFK_BOX_TRI>(xyz, F, natoms, idx1, n1, idx2, n2,
                                      same_group, box, hp, excl_a, excl_b, hist,
                                      ws, L, sms, st);
  }
}

extern "C" int rdfk_rdf_sites(const float *xyz, int F, int natoms,
                              const int *idxA, const int *idxB, const int *offA,
                              const int *offB, const long long *cell_off,
                              int n_sites, long long total_cells, int boxkind,
                              const float *box, double r0, double r1, int nbins,
                              const double *edges, unsigned int *count,
                              void *stream) {
  if (F < 0 || natoms < 0 || n_sites < 0 || total_cells < 0 || nbins < 1)
    return set_err(RDFK_EINVAL, "negative size or nbins < 1%s%s");
  if (!(r1 > r0)) return set_err(RDFK_EINVAL, "need r1 > r0%s%s");
  if (boxkind < RDFK_BOX_NONE || boxkind > RDFK_BOX_TRI)
    return set_err(RDFK_EINVAL, "bad boxkind%s%s");
  if (F == 0 || n_sites == 0 || total_cells == 0) return RDFK_OK;
  if (!xyz || !idxA || !idxB || !offA || !offB || !cell_off || !edges ||
      !count || (boxkind != RDFK_BOX_NONE && !box))
    return set_err(RDFK_EINVAL, "null pointer argument%s%s");
  int sms = 0;
  int rc = check_device(&sms);
  if (rc) return rc;
  if (F > 65535) return set_err(RDFK_EINVAL, "at most 65535 frames per call%s%s");
  HistParams hp;
  hp.r0 = r0; hp.r1 = r1; hp.edges = edges; hp.nbins = nbins;
  cudaStream_t st = (cudaStream_t)stream;
  const long long blocks = (total_cells + 255) / 256;
  if (blocks > 0x7fffffffLL) return set_err(RDFK_EINVAL, "too many cells%s%s");
  dim3 grid((unsigned)blocks, (unsigned)F);
  switch (boxkind) {
    case RDFK_BOX_NONE:
      sites_kernel<RDFK_BOX_NONE><<<grid, 256, 0, st>>>(
          xyz, natoms, idxA, idxB, offA, offB, cell_off, n_sites, total_cells,
          box, hp, count);
      break;
    case RDFK_BOX_ORTHO:
      sites_kernel<RDFK_BOX_ORTHO><<<grid, 256, 0, st>>>(
          xyz, natoms, idxA, idxB, offA, offB, cell_off, n_sites, total_cells,
          box, hp, count);
      break;
    default:
      sites_kernel<RDFK_BOX_TRI><<<grid, 256, 0, st>>>(
          xyz, natoms, idxA, idxB, offA, offB, cell_off, n_sites, total_cells,
          box, hp, count);
  }
  COUNT_LAUNCH(1);
  CUDA_TRY(cudaGetLastError());
  return RDFK_OK;
}

extern "C" unsigned long long rdfk_launch_count(void) {
  return __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
}

extern "C" int rdfk_rdf_stats(const void *workspace,
                              unsigned long long *out_host, void *stream) {
  if (!workspace || !out_host)
    return set_err(RDFK_EINVAL, "null pointer argument%s%s");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemcpyAsync(out_host, workspace, 4 * sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return RDFK_OK;
}

extern "C" int rdfk_measure_fp32_peak(double *tflops_host, void *stream) {
  if (!tflops_host) return set_err(RDFK_EINVAL, "null pointer argument%s%s");
  int sms = 0;
  int rc = check_device(&sms);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  float *out = nullptr;
  CUDA_TRY(cudaMalloc(&out, sizeof(float)));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int iters = 4096, grid = sms * 8;
  ffma_peak_kernel<<<grid, 256, 0, st>>>(out, 64, 1.0001f, 0.5f);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_TRY(cudaEventRecord(e0, st));
    ffma_peak_kernel<<<grid, 256, 0, st>>>(out, iters, 1.0001f, 0.5f);
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * 8 * 16 * (double)iters * 256.0 * grid;
    const double tf = flop / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops_host = best;
  return RDFK_OK;
}
