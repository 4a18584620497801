// hbond.cuh -- donor..acceptor search + D-H..A angle filter with a ragged
// (compacted) result: the hot path of pmda.hbond_analysis.HydrogenBondAnalysis
// (pmda/hbond_analysis.py:415-470) and, without hydrogens, a stand-alone
// MDAnalysis.lib.distances.capped_distance(ref, conf, max_cutoff, min_cutoff,
// box).  Included at the end of rdfk.cu (one translation unit: it reuses
// stage_kernel / params_kernel / tables_kernel and their error bounds).
//
// Per call (F frames):
//   stage donors, hydrogens, acceptors (gather, AoS -> SoA, triclinic wrap)
//   params + tables          -> r2_cut: fp32 r2 >= r2_cut  =>  reference d > cut
//   hb_cell_count / scan / hb_cell_scatter
//                            acceptors counting-sorted into a cell grid whose
//                            cells are at least the cut-off wide
//   hb_cells_kernel<false>   every (frame, D-H row) walks the 27 cells around
//                            its donor: fp32 screen, survivors remembered
//                            (HB_K per row), then decided in fp64 -> row counts
//   hb_scan_kernel           exclusive scan of the per-CTA totals (+ grand total)
//   hb_cells_kernel<true>    rows written at their offsets: replays the
//                            remembered candidates (a row with more walks again)
//   (rdfk_set_debug_flags bit 0: hb_kernel instead -- no grid, every acceptor, tiled
//   through shared memory; same results)
// The result is ordered by (frame, D-H row, acceptor): the row-major order of
// the reference's brute-force `np.where(mask)`.
//
// Arithmetic: every pair is screened with the fp32 per-pair minimum image of
// the RDF kernel; survivors (a fraction of a percent) are decided in fp64 with
// the reference's operation order -- d <= max_cutoff, d > min_cutoff, and
// rad2deg(atan2(|rji x rjk|, rji . rjk)) > angle_cutoff -- so the pair set and
// the distances are the oracle's bit for bit; angles go through CUDA's atan2
// (<= 2 ulp from glibc's).

constexpr int HB_TJ = 1024;  // acceptors staged in shared memory per round
constexpr int HB_K = 12;     // screened candidates remembered per row
constexpr int HB_NKMAX = 32;                                   // cells per axis
constexpr int HB_NCS = HB_NKMAX * HB_NKMAX * HB_NKMAX + 1;     // + end marker

struct HbLayout {
  size_t header, extents, fparams, tables, edges, gd, gh, ga, rowcount, ncand,
      cand, blocksum, cells, cid, crank, sorted, total;
  int npd, npa, nblk_x;
};

static HbLayout hb_layout(int F, int n_dh, int n_acc) {
  HbLayout w;
  w.npd = (int)align_up((size_t)(n_dh > 0 ? n_dh : 1), TILE);
  w.npa = (int)align_up((size_t)(n_acc > 0 ? n_acc : 1), TILE);
  w.nblk_x = (n_dh + 255) / 256;
  size_t off = 0;
  w.header = off;
  off = align_up(off + sizeof(WsHeader), 256);
  w.extents = off;
  off = align_up(off + sizeof(int) * EXT * (size_t)F, 256);
  w.fparams = off;
  off = align_up(off + sizeof(FrameParams) * (size_t)F, 256);
  w.tables = off;
  off = align_up(off + sizeof(float2) * 2 * 3, 256);   // [2][nbins + 2], nbins = 1
  w.edges = off;
  off = align_up(off + sizeof(double) * 2, 256);
  w.gd = off;
  off = align_up(off + sizeof(float) * 3 * (size_t)w.npd * (size_t)F, 256);
  w.gh = off;
  off = align_up(off + sizeof(float) * 3 * (size_t)w.npd * (size_t)F, 256);
  w.ga = off;
  off = align_up(off + sizeof(float) * 3 * (size_t)w.npa * (size_t)F, 256);
  w.rowcount = off;
  off = align_up(off + sizeof(uint32_t) * (size_t)w.nblk_x * 256 * (size_t)F, 256);
  w.ncand = off;
  off = align_up(off + sizeof(uint32_t) * (size_t)w.nblk_x * 256 * (size_t)F, 256);
  w.cand = off;
  off = align_up(off + sizeof(int) * HB_K * (size_t)w.nblk_x * 256 * (size_t)F, 256);
  w.blocksum = off;
  off = align_up(off + sizeof(long long) * ((size_t)w.nblk_x * (size_t)F + 1), 256);
  // acceptor cell grid: per-frame cell starts, per-atom cell and arrival rank,
  // acceptors in cell order as (x, y, z, index)
  w.cells = off;
  off = align_up(off + sizeof(uint32_t) * (size_t)HB_NCS * (size_t)F, 256);
  w.cid = off;
  off = align_up(off + sizeof(uint32_t) * (size_t)w.npa * (size_t)F, 256);
  w.crank = off;
  off = align_up(off + sizeof(uint32_t) * (size_t)w.npa * (size_t)F, 256);
  w.sorted = off;
  off = align_up(off + sizeof(float4) * (size_t)w.npa * (size_t)F, 256);
  w.total = off;
  return w;
}

// minimum-image vector r_j - r_i with the reference's arithmetic (see
// exact_bin): float difference widened to double, then minimum_image /
// minimum_image_triclinic.  Triclinic inputs are the wrapped coordinates.
template <int BOXKIND>
__device__ __forceinline__ void exact_vec(float xi, float yi, float zi, float xj,
                                          float yj, float zj,
                                          const FrameParams *__restrict__ fp,
                                          double v[3]) {
  v[0] = (double)__fsub_rn(xj, xi);
  v[1] = (double)__fsub_rn(yj, yi);
  v[2] = (double)__fsub_rn(zj, zi);
  if (BOXKIND == RDFK_BOX_ORTHO) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float L = fp->L[k];
      if (L > FLT_EPSILON) {
        const double s = __dmul_rn((double)fp->inv[k], v[k]);
        v[k] = __dmul_rn((double)L, __dsub_rn(s, round(s)));
      }
    }
  } else if (BOXKIND == RDFK_BOX_TRI) {
    const float *b = fp->tri;
    double mn0 = 0.0, mn1 = 0.0, mn2 = 0.0;
    double dsq_min = (double)FLT_MAX;
    for (int ix = -1; ix < 2; ++ix) {
      const double rx = __dadd_rn(v[0], (double)__fmul_rn(b[0], (float)ix));
      for (int iy = -1; iy < 2; ++iy) {
        const double ry0 = __dadd_rn(rx, (double)__fmul_rn(b[3], (float)iy));
        const double ry1 = __dadd_rn(v[1], (double)__fmul_rn(b[4], (float)iy));
        for (int iz = -1; iz < 2; ++iz) {
          const double rz0 = __dadd_rn(ry0, (double)__fmul_rn(b[6], (float)iz));
          const double rz1 = __dadd_rn(ry1, (double)__fmul_rn(b[7], (float)iz));
          const double rz2 = __dadd_rn(v[2], (double)__fmul_rn(b[8], (float)iz));
          const double dsq =
              __dadd_rn(__dadd_rn(__dmul_rn(rz0, rz0), __dmul_rn(rz1, rz1)),
                        __dmul_rn(rz2, rz2));
          if (dsq < dsq_min) {
            dsq_min = dsq;
            mn0 = rz0; mn1 = rz1; mn2 = rz2;
          }
        }
      }
    }
    v[0] = mn0; v[1] = mn1; v[2] = mn2;
  }
}

struct HbArgs {
  const float *gd, *gh, *ga;  // staged donors / hydrogens / acceptors (SoA)
  int npd, npa, n_dh, n_acc;
  const FrameParams *fparams;
  const WsHeader *header;
  double max_cutoff, min_cutoff, angle_cutoff;  // min_cutoff < 0: none
  int has_h;
  uint32_t *rowcount;     // [F][nblk_x * 256] hydrogen bonds of the row
  uint32_t *ncand;        // [F][nblk_x * 256] pairs that passed the fp32 screen
  int *cand;              // [F][nblk_x * 256][HB_K] the first HB_K of them
  long long *blocksum;    // [F * nblk_x + 1]: totals, then exclusive offsets
  const uint32_t *cells;  // [F][HB_NCS] cell starts (exclusive scan of counts)
  const float4 *sorted;   // [F][npa] acceptors in cell order: x, y, z, index
  long long capacity;
  int *out_row, *out_acc;
  double *out_dist, *out_angle;
};

// One candidate of the fp32 screen, decided with the reference arithmetic.
// Returns true for a hydrogen bond (or a capped pair when there is no
// hydrogen); *d_out / *ang_out are the reference's float64 results.
template <int BOXKIND>
__device__ __noinline__ bool hb_decide(float xd, float yd, float zd, float xh,
                                       float yh, float zh, float xa, float ya,
                                       float za,
                                       const FrameParams *__restrict__ fp,
                                       double max_cutoff, double min_cutoff,
                                       double angle_cutoff, int has_h,
                                       double *d_out, double *ang_out) {
  double v[3];
  exact_vec<BOXKIND>(xd, yd, zd, xa, ya, za, fp, v);
  const double d = __dsqrt_rn(__dadd_rn(
      __dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])),
      __dmul_rn(v[2], v[2])));
  if (!(d <= max_cutoff)) return false;
  if (min_cutoff >= 0.0 && !(d > min_cutoff)) return false;
  *d_out = d;
  *ang_out = 0.0;
  if (!has_h) return true;
  // calc_angles(D, H, A): rji = D - H, rjk = A - H, both minimum images
  double rji[3], rjk[3];
  exact_vec<BOXKIND>(xh, yh, zh, xd, yd, zd, fp, rji);
  exact_vec<BOXKIND>(xh, yh, zh, xa, ya, za, fp, rjk);
  const double x = __dadd_rn(
      __dadd_rn(__dmul_rn(rji[0], rjk[0]), __dmul_rn(rji[1], rjk[1])),
      __dmul_rn(rji[2], rjk[2]));
  const double xp0 = __dsub_rn(__dmul_rn(rji[1], rjk[2]), __dmul_rn(rji[2], rjk[1]));
  const double xp1 = __dadd_rn(__dmul_rn(-rji[0], rjk[2]), __dmul_rn(rji[2], rjk[0]));
  const double xp2 = __dsub_rn(__dmul_rn(rji[0], rjk[1]), __dmul_rn(rji[1], rjk[0]));
  const double y = __dsqrt_rn(__dadd_rn(
      __dadd_rn(__dmul_rn(xp0, xp0), __dmul_rn(xp1, xp1)), __dmul_rn(xp2, xp2)));
  const double ang = __dmul_rn(atan2(y, x), 57.29577951308232);  // np.rad2deg
  *ang_out = ang;
  return ang > angle_cutoff;
}

template <int BOXKIND, bool FILL>
__global__ __launch_bounds__(256) void hb_kernel(const HbArgs a) {
  __shared__ __align__(16) float sx[HB_TJ], sy[HB_TJ], sz[HB_TJ];
  __shared__ long long s_warp[8];
  const int f = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.x * 256 + tid;
  const bool valid = k < a.n_dh;
  const FrameParams *fp = a.fparams + f;
  const FastBox fb = load_fastbox(fp);
  const float cut = a.header->r2_cut[1];
  const bool slow = fp->all_slow != 0;
  const float *gd = a.gd + (size_t)f * 3 * a.npd;
  const float *ga = a.ga + (size_t)f * 3 * a.npa;
  const float nan = __int_as_float(0x7fc00000);
  // pads of the staged arrays are NaN: never a candidate
  const float xd = valid ? gd[k] : nan, yd = valid ? gd[a.npd + k] : nan,
              zd = valid ? gd[2 * (size_t)a.npd + k] : nan;
  const float nxd = xd, nyd = yd, nzd = zd;   // (pair_r2 subtracts)
  float xh = 0.f, yh = 0.f, zh = 0.f;
  if (a.has_h && valid) {
    const float *gh = a.gh + (size_t)f * 3 * a.npd;
    xh = gh[k]; yh = gh[a.npd + k]; zh = gh[2 * (size_t)a.npd + k];
  }
  const size_t row = (size_t)f * gridDim.x * 256 + (size_t)blockIdx.x * 256 + tid;
  const size_t blk = (size_t)f * gridDim.x + blockIdx.x;

  long long pos = 0;
  if (FILL) {
    // my offset = CTA offset + exclusive prefix of the row counts of the CTA
    const uint32_t mine = a.rowcount[row];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    long long before = a.blocksum[blk];
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    pos = before + (long long)(incl - mine);
    __syncthreads();
  }

  uint32_t cnt = 0, nc = 0;
  // a pair that passed the fp32 screen (or any pair of an fp64-only frame):
  // the reference decision, and in the fill pass the output row
  auto decide = [&](int j, float xa, float ya, float za) {
    double d, ang;
    if (!hb_decide<BOXKIND>(xd, yd, zd, xh, yh, zh, xa, ya, za, fp, a.max_cutoff,
                            a.min_cutoff, a.angle_cutoff, a.has_h, &d, &ang))
      return;
    if (FILL) {
      const long long p = pos + cnt;
      if (p < a.capacity) {
        a.out_row[p] = f * a.n_dh + k;
        a.out_acc[p] = j;
        a.out_dist[p] = d;
        if (a.out_angle) a.out_angle[p] = ang;
      }
    }
    ++cnt;
  };
  int *my_cand = a.cand + row * HB_K;

  // Phase A (count pass): the fp32 screen.  Candidates are only remembered
  // here -- deciding them inside the loop would run the fp64 code with one
  // active lane per warp.
  if (!FILL) {
    if (slow) {
      nc = valid ? HB_K + 1 : 0;      // no screen: straight to phase C
    } else {
      auto push = [&](int j) {
        if (nc < (uint32_t)HB_K) my_cand[nc] = j;
        ++nc;
      };
      for (int j0 = 0; j0 < a.npa; j0 += HB_TJ) {
        const int tj = min(HB_TJ, a.npa - j0);   // multiple of 128
        for (int t = tid; t < tj; t += 256) {
          sx[t] = ga[j0 + t];
          sy[t] = ga[(size_t)a.npa + j0 + t];
          sz[t] = ga[2 * (size_t)a.npa + j0 + t];
        }
        __syncthreads();
#pragma unroll 2
        for (int t = 0; t < tj; t += 4) {
          const float4 X = *reinterpret_cast<const float4 *>(sx + t);
          const float4 Y = *reinterpret_cast<const float4 *>(sy + t);
          const float4 Z = *reinterpret_cast<const float4 *>(sz + t);
          const float r0 = pair_r2<BOXKIND, false>(X.x, Y.x, Z.x, nxd, nyd, nzd, fb);
          const float r1 = pair_r2<BOXKIND, false>(X.y, Y.y, Z.y, nxd, nyd, nzd, fb);
          const float r2 = pair_r2<BOXKIND, false>(X.z, Y.z, Z.z, nxd, nyd, nzd, fb);
          const float r3 = pair_r2<BOXKIND, false>(X.w, Y.w, Z.w, nxd, nyd, nzd, fb);
          if (fminf(fminf(r0, r1), fminf(r2, r3)) < cut) {
            if (r0 < cut) push(j0 + t);
            if (r1 < cut) push(j0 + t + 1);
            if (r2 < cut) push(j0 + t + 2);
            if (r3 < cut) push(j0 + t + 3);
          }
        }
        __syncthreads();
      }
    }
  } else {
    // fill pass: rows without a hydrogen bond have nothing to write
    nc = a.rowcount[row] > 0 ? a.ncand[row] : 0;
  }

  // Phase B: decide the remembered candidates, lanes side by side.
  const bool over = nc > (uint32_t)HB_K;
  if (!over) {
    for (uint32_t c = 0; c < nc; ++c) {
      const int j = my_cand[c];
      decide(j, ga[j], ga[(size_t)a.npa + j], ga[2 * (size_t)a.npa + j]);
    }
  }

  // Phase C: a row with more than HB_K candidates (or an fp64-only frame)
  // makes its CTA walk the acceptors again, deciding in place.
  if (__syncthreads_or(over ? 1 : 0)) {
    for (int j0 = 0; j0 < a.npa; j0 += HB_TJ) {
      const int tj = min(HB_TJ, a.npa - j0);
      for (int t = tid; t < tj; t += 256) {
        sx[t] = ga[j0 + t];
        sy[t] = ga[(size_t)a.npa + j0 + t];
        sz[t] = ga[2 * (size_t)a.npa + j0 + t];
      }
      __syncthreads();
      if (over) {
        const int lim = min(tj, a.n_acc - j0);
        for (int t = 0; t < lim; ++t) {
          if (!slow &&
              !(pair_r2<BOXKIND, false>(sx[t], sy[t], sz[t], nxd, nyd, nzd, fb) < cut))
            continue;
          decide(j0 + t, sx[t], sy[t], sz[t]);
        }
      }
      __syncthreads();
    }
  }

  if (!FILL) {
    a.rowcount[row] = cnt;
    a.ncand[row] = nc;
    long long s = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_warp[warp] = s;
    __syncthreads();
    if (tid == 0) {
      long long tot = 0;
      for (int w = 0; w < 8; ++w) tot += s_warp[w];
      a.blocksum[blk] = tot;
    }
  }
}

// One CTA: exclusive scan of the per-CTA totals in place; grand total to
// blocksum[n] and *total.
__global__ __launch_bounds__(1024) void hb_scan_kernel(long long *blocksum,
                                                       long long n,
                                                       long long *total) {
  __shared__ long long s_w[32];
  __shared__ long long s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (long long base = 0; base < n; base += 1024) {
    const long long i = base + tid;
    const long long v = i < n ? blocksum[i] : 0;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    long long before = s_carry;
    for (int w = 0; w < warp; ++w) before += s_w[w];
    if (i < n) blocksum[i] = before + incl - v;
    __syncthreads();
    if (tid == 1023) s_carry = before + incl;
    __syncthreads();
  }
  if (tid == 0) {
    blocksum[n] = s_carry;
    *total = s_carry;
  }
}

// ---------------------------------------------------------------------------
// Cell grid over the acceptors.  Cells live in the RDF path's "prune space"
// (FrameParams: Cartesian offset wrapped into the box, or triclinic fractional
// coordinates), at least `rprune` (cut-off + fp32 slack) wide in every
// direction, so every acceptor within the cut-off of a donor sits in one of the
// 27 cells around the donor's.  A periodic axis with fewer than 3 cells, and
// every axis of a frame without usable bounds, gets a single cell.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void hb_grid(const FrameParams &p, int nk[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int n = 1;
    if (p.prune_ok && !p.all_slow && p.gscale[k] > 0.f) {
      const float cells = __fdiv_rd(p.wk[k], __fmul_ru(p.gscale[k], p.rprune));
      n = cells >= (float)HB_NKMAX ? HB_NKMAX : (cells >= 1.f ? (int)cells : 1);
      if (p.per[k] < INFINITY && n < 3) n = 1;
    }
    nk[k] = n;
  }
}

template <int BOXKIND>
__device__ __forceinline__ void hb_cell_of(float x, float y, float z,
                                           const FrameParams &p, const int nk[3],
                                           int c[3]) {
  float u[3];
  prune_coords<BOXKIND>(x, y, z, p, u[0], u[1], u[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // NaN -> 0; a coordinate that rounds just outside the grid is clamped into
    // the edge cell, a neighbour of the cell it belongs to
    const int v = __float2int_rd(__fmul_rn(__fmul_rn(u[k], p.gscale[k]), (float)nk[k]));
    c[k] = max(0, min(nk[k] - 1, v));
  }
}

template <int BOXKIND>
__global__ __launch_bounds__(256) void hb_cell_count_kernel(
    const float *__restrict__ ga, int n, int np,
    const FrameParams *__restrict__ fparams, uint32_t *__restrict__ cells,
    uint32_t *__restrict__ cid, uint32_t *__restrict__ crank) {
  const int f = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const FrameParams &p = fparams[f];
  const float *g = ga + (size_t)f * 3 * np;
  int nk[3], c[3];
  hb_grid(p, nk);
  hb_cell_of<BOXKIND>(g[a], g[(size_t)np + a], g[2 * (size_t)np + a], p, nk, c);
  const uint32_t id = (uint32_t)((c[2] * nk[1] + c[1]) * nk[0] + c[0]);
  cid[(size_t)f * np + a] = id;
  crank[(size_t)f * np + a] = atomicAdd(cells + (size_t)f * HB_NCS + id, 1u);
}

// exclusive scan of one frame's cell counts (the cells the frame's grid uses,
// plus the end marker), one CTA per frame
__global__ __launch_bounds__(1024) void hb_cell_scan_kernel(
    uint32_t *__restrict__ cells, const FrameParams *__restrict__ fparams) {
  int nk[3];
  hb_grid(fparams[blockIdx.x], nk);
  const int ncell = nk[0] * nk[1] * nk[2] + 1;
  uint32_t *c = cells + (size_t)blockIdx.x * HB_NCS;
  const int tid = threadIdx.x;
  const int strip = (ncell + 1023) / 1024;
  const int lo = min(ncell, tid * strip), hi = min(ncell, lo + strip);
  uint32_t s = 0;
  for (int i = lo; i < hi; ++i) s += c[i];
  __shared__ uint32_t wsum[32];
  uint32_t incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += v;
  }
  if ((tid & 31) == 31) wsum[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    uint32_t w = wsum[tid];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
      if (tid >= o) w += v;
    }
    wsum[tid] = w;
  }
  __syncthreads();
  uint32_t run = incl - s + ((tid >> 5) ? wsum[(tid >> 5) - 1] : 0u);
  for (int i = lo; i < hi; ++i) {
    const uint32_t v = c[i];
    c[i] = run;
    run += v;
  }
}

__global__ __launch_bounds__(256) void hb_cell_scatter_kernel(
    const float *__restrict__ ga, int n, int np,
    const uint32_t *__restrict__ cells, const uint32_t *__restrict__ cid,
    const uint32_t *__restrict__ crank, float4 *__restrict__ sorted) {
  const int f = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  const float *g = ga + (size_t)f * 3 * np;
  const uint32_t pos = cells[(size_t)f * HB_NCS + cid[(size_t)f * np + a]] +
                       crank[(size_t)f * np + a];
  sorted[(size_t)f * np + pos] =
      make_float4(g[a], g[(size_t)np + a], g[2 * (size_t)np + a], __int_as_float(a));
}

// One thread per (frame, D-H row): walk the 27 cells around the donor.  Same
// two passes and the same candidate replay as hb_kernel; a row's bonds come
// out in cell order and are sorted by acceptor index at the end of the fill.
template <int BOXKIND, bool FILL>
__global__ __launch_bounds__(256) void hb_cells_kernel(const HbArgs a) {
  __shared__ long long s_warp[8];
  const int f = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.x * 256 + tid;
  const bool valid = k < a.n_dh;
  const FrameParams *fp = a.fparams + f;
  const FastBox fb = load_fastbox(fp);
  const float cut = a.header->r2_cut[1];
  const bool slow = fp->all_slow != 0;
  const float *gd = a.gd + (size_t)f * 3 * a.npd;
  const float4 *sa = a.sorted + (size_t)f * a.npa;
  const uint32_t *cs = a.cells + (size_t)f * HB_NCS;
  const float nan = __int_as_float(0x7fc00000);
  const float xd = valid ? gd[k] : nan, yd = valid ? gd[a.npd + k] : nan,
              zd = valid ? gd[2 * (size_t)a.npd + k] : nan;
  const float nxd = xd, nyd = yd, nzd = zd;   // (pair_r2 subtracts)
  float xh = 0.f, yh = 0.f, zh = 0.f;
  if (a.has_h && valid) {
    const float *gh = a.gh + (size_t)f * 3 * a.npd;
    xh = gh[k]; yh = gh[a.npd + k]; zh = gh[2 * (size_t)a.npd + k];
  }
  const size_t row = (size_t)f * gridDim.x * 256 + (size_t)blockIdx.x * 256 + tid;
  const size_t blk = (size_t)f * gridDim.x + blockIdx.x;

  long long pos = 0;
  if (FILL) {
    const uint32_t mine = a.rowcount[row];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    long long before = a.blocksum[blk];
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    pos = before + (long long)(incl - mine);
  }

  uint32_t cnt = 0, nc = 0;
  auto decide = [&](const float4 A) {
    double d, ang;
    if (!hb_decide<BOXKIND>(xd, yd, zd, xh, yh, zh, A.x, A.y, A.z, fp, a.max_cutoff,
                            a.min_cutoff, a.angle_cutoff, a.has_h, &d, &ang))
      return;
    if (FILL) {
      const long long p = pos + cnt;
      if (p < a.capacity) {
        a.out_row[p] = f * a.n_dh + k;
        a.out_acc[p] = __float_as_int(A.w);
        a.out_dist[p] = d;
        if (a.out_angle) a.out_angle[p] = ang;
      }
    }
    ++cnt;
  };
  int *my_cand = a.cand + row * HB_K;

  // the cells around mine; `each(lo, hi)` gets their ranges of `sorted`
  int nk[3], c[3];
  hb_grid(*fp, nk);
  hb_cell_of<BOXKIND>(xd, yd, zd, *fp, nk, c);
  auto walk = [&](auto &&each) {
    for (int dz = -1; dz <= 1; ++dz) {
      int z = c[2] + dz;
      if (nk[2] == 1 ? dz != 0 : false) continue;
      if (z < 0 || z >= nk[2]) {
        if (!(fp->per[2] < INFINITY)) continue;
        z += z < 0 ? nk[2] : -nk[2];
      }
      for (int dy = -1; dy <= 1; ++dy) {
        int y = c[1] + dy;
        if (nk[1] == 1 ? dy != 0 : false) continue;
        if (y < 0 || y >= nk[1]) {
          if (!(fp->per[1] < INFINITY)) continue;
          y += y < 0 ? nk[1] : -nk[1];
        }
        // the (up to) three x cells are contiguous unless they wrap
        for (int dx = -1; dx <= 1; ++dx) {
          int x = c[0] + dx;
          if (nk[0] == 1 ? dx != 0 : false) continue;
          if (x < 0 || x >= nk[0]) {
            if (!(fp->per[0] < INFINITY)) continue;
            x += x < 0 ? nk[0] : -nk[0];
          }
          const int id = (z * nk[1] + y) * nk[0] + x;
          each((int)cs[id], (int)cs[id + 1]);
        }
      }
    }
  };

  const bool want = FILL ? a.rowcount[row] > 0 : valid;
  if (!FILL) {
    // count pass: fp32 screen, candidates (positions in `sorted`) remembered
    if (want) {
      walk([&](int lo, int hi) {
        for (int t = lo; t < hi; ++t) {
          const float4 A = sa[t];
          if (slow || pair_r2<BOXKIND, false>(A.x, A.y, A.z, nxd, nyd, nzd, fb) < cut) {
            if (nc < (uint32_t)HB_K) my_cand[nc] = t;
            ++nc;
          }
        }
      });
    }
  } else {
    nc = want ? a.ncand[row] : 0;
  }
  if (nc <= (uint32_t)HB_K) {
    for (uint32_t t = 0; t < nc; ++t) decide(sa[my_cand[t]]);
  } else {
    // more candidates than remembered: walk again, deciding in place
    walk([&](int lo, int hi) {
      for (int t = lo; t < hi; ++t) {
        const float4 A = sa[t];
        if (slow || pair_r2<BOXKIND, false>(A.x, A.y, A.z, nxd, nyd, nzd, fb) < cut)
          decide(A);
      }
    });
  }

  if (!FILL) {
    a.rowcount[row] = cnt;
    a.ncand[row] = nc;
    long long s = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_warp[warp] = s;
    __syncthreads();
    if (tid == 0) {
      long long tot = 0;
      for (int w = 0; w < 8; ++w) tot += s_warp[w];
      a.blocksum[blk] = tot;
    }
  } else if (cnt > 1) {
    // my bonds by ascending acceptor index (the reference's order); a handful
    // per row: insertion sort in place
    const long long end = min(pos + (long long)cnt, a.capacity);
    for (long long i = pos + 1; i < end; ++i) {
      const int ka = a.out_acc[i];
      const double kd = a.out_dist[i];
      const double kg = a.out_angle ? a.out_angle[i] : 0.0;
      long long j = i - 1;
      while (j >= pos && a.out_acc[j] > ka) {
        a.out_acc[j + 1] = a.out_acc[j];
        a.out_dist[j + 1] = a.out_dist[j];
        if (a.out_angle) a.out_angle[j + 1] = a.out_angle[j];
        --j;
      }
      a.out_acc[j + 1] = ka;
      a.out_dist[j + 1] = kd;
      if (a.out_angle) a.out_angle[j + 1] = kg;
    }
  }
}

__global__ void hb_edges_kernel(double *edges, double r1) {
  edges[0] = 0.0;
  edges[1] = r1;
}

template <int BOXKIND>
static int launch_hbonds(const float *xyz, int F, int natoms, const int *donors,
                         const int *hydrogens, int n_dh, const int *acceptors,
                         int n_acc, const float *box, double max_cutoff,
                         double min_cutoff, double angle_cutoff,
                         long long capacity, int *out_row, int *out_acc,
                         double *out_dist, double *out_angle, long long *total,
                         char *ws, const HbLayout &L, cudaStream_t st) {
  WsHeader *header = reinterpret_cast<WsHeader *>(ws + L.header);
  int *extents = reinterpret_cast<int *>(ws + L.extents);
  FrameParams *fparams = reinterpret_cast<FrameParams *>(ws + L.fparams);
  float2 *tables = reinterpret_cast<float2 *>(ws + L.tables);
  double *edges = reinterpret_cast<double *>(ws + L.edges);
  float *gd = reinterpret_cast<float *>(ws + L.gd);
  float *gh = reinterpret_cast<float *>(ws + L.gh);
  float *ga = reinterpret_cast<float *>(ws + L.ga);

  init_ws_kernel<<<(F * EXT + 255) / 256 + 1, 256, 0, st>>>(header, extents, F, 0);
  hb_edges_kernel<<<1, 1, 0, st>>>(edges, max_cutoff);
  stage_kernel<BOXKIND><<<dim3((L.npd + 255) / 256, F), 256, 0, st>>>(
      xyz, natoms, donors, n_dh, L.npd, box, gd, extents);
  stage_kernel<BOXKIND><<<dim3((L.npa + 255) / 256, F), 256, 0, st>>>(
      xyz, natoms, acceptors, n_acc, L.npa, box, ga, extents);
  if (hydrogens)
    stage_kernel<BOXKIND><<<dim3((L.npd + 255) / 256, F), 256, 0, st>>>(
        xyz, natoms, hydrogens, n_dh, L.npd, box, gh, extents);
  HistParams hp;
  hp.r0 = 0.0; hp.r1 = max_cutoff; hp.edges = edges; hp.nbins = 1;
  // debug flag 2: bounds of the per-pair minimum image only (no shifted variant)
  params_kernel<<<(F + 127) / 128, 128, 0, st>>>(BOXKIND, box, extents, hp, 2, F,
                                                 fparams);
  tables_kernel<<<1, 256, 0, st>>>(hp, F, fparams, tables, header, extents);
  COUNT_LAUNCH(hydrogens ? 7 : 6);

  HbArgs a;
  a.gd = gd; a.gh = gh; a.ga = ga;
  a.npd = L.npd; a.npa = L.npa; a.n_dh = n_dh; a.n_acc = n_acc;
  a.fparams = fparams; a.header = header;
  a.max_cutoff = max_cutoff; a.min_cutoff = min_cutoff;
  a.angle_cutoff = angle_cutoff;
  a.has_h = hydrogens ? 1 : 0;
  a.rowcount = reinterpret_cast<uint32_t *>(ws + L.rowcount);
  a.ncand = reinterpret_cast<uint32_t *>(ws + L.ncand);
  a.cand = reinterpret_cast<int *>(ws + L.cand);
  a.blocksum = reinterpret_cast<long long *>(ws + L.blocksum);
  a.cells = nullptr; a.sorted = nullptr;
  a.capacity = capacity;
  a.out_row = out_row; a.out_acc = out_acc;
  a.out_dist = out_dist; a.out_angle = out_angle;
  const dim3 grid((unsigned)L.nblk_x, (unsigned)F);
  // rdfk_set_debug_flags bit 0: no cell grid, every row against every acceptor (tiled
  // through shared memory); the results must not depend on it (tests/)
  const int debug_flags = __atomic_load_n(&g_debug_flags, __ATOMIC_RELAXED);
  if (debug_flags & 1) {
    hb_kernel<BOXKIND, false><<<grid, 256, 0, st>>>(a);
    hb_scan_kernel<<<1, 1024, 0, st>>>(a.blocksum, (long long)L.nblk_x * F, total);
    hb_kernel<BOXKIND, true><<<grid, 256, 0, st>>>(a);
    COUNT_LAUNCH(3);
  } else {
    uint32_t *cells = reinterpret_cast<uint32_t *>(ws + L.cells);
    uint32_t *cid = reinterpret_cast<uint32_t *>(ws + L.cid);
    uint32_t *crank = reinterpret_cast<uint32_t *>(ws + L.crank);
    float4 *sorted = reinterpret_cast<float4 *>(ws + L.sorted);
    CUDA_TRY(cudaMemsetAsync(cells, 0, sizeof(uint32_t) * (size_t)HB_NCS * F, st));
    const dim3 agrid((unsigned)((n_acc + 255) / 256), (unsigned)F);
    hb_cell_count_kernel<BOXKIND><<<agrid, 256, 0, st>>>(ga, n_acc, L.npa, fparams,
                                                         cells, cid, crank);
    hb_cell_scan_kernel<<<F, 1024, 0, st>>>(cells, fparams);
    hb_cell_scatter_kernel<<<agrid, 256, 0, st>>>(ga, n_acc, L.npa, cells, cid,
                                                  crank, sorted);
    a.cells = cells;
    a.sorted = sorted;
    hb_cells_kernel<BOXKIND, false><<<grid, 256, 0, st>>>(a);
    hb_scan_kernel<<<1, 1024, 0, st>>>(a.blocksum, (long long)L.nblk_x * F, total);
    hb_cells_kernel<BOXKIND, true><<<grid, 256, 0, st>>>(a);
    COUNT_LAUNCH(6);
  }
  CUDA_TRY(cudaGetLastError());
  return RDFK_OK;
}

extern "C" size_t rdfk_hbonds_workspace_bytes(int F, int n_dh, int n_acc) {
  if (F < 0 || n_dh < 0 || n_acc < 0) return 0;
  return hb_layout(F, n_dh, n_acc).total;
}

extern "C" int rdfk_hbonds(const float *xyz, int F, int natoms,
                           const int *donors, const int *hydrogens, int n_dh,
                           const int *acceptors, int n_acc, int boxkind,
                           const float *box, double max_cutoff,
                           double min_cutoff, double angle_cutoff_deg,
                           long long capacity, int *out_row, int *out_acc,
                           double *out_dist, double *out_angle,
                           long long *total, void *workspace,
                           size_t workspace_bytes, void *stream) {
  if (F < 0 || natoms < 0 || n_dh < 0 || n_acc < 0 || capacity < 0)
    return set_err(RDFK_EINVAL, "negative size%s%s");
  if (!(max_cutoff > 0.0)) return set_err(RDFK_EINVAL, "need max_cutoff > 0%s%s");
  if (boxkind < RDFK_BOX_NONE || boxkind > RDFK_BOX_TRI)
    return set_err(RDFK_EINVAL, "bad boxkind%s%s");
  if (!total) return set_err(RDFK_EINVAL, "null pointer argument%s%s");
  cudaStream_t st = (cudaStream_t)stream;
  if (F == 0 || n_dh == 0 || n_acc == 0) {
    CUDA_TRY(cudaMemsetAsync(total, 0, sizeof(long long), st));
    return RDFK_OK;
  }
  if (!xyz || !workspace || (boxkind != RDFK_BOX_NONE && !box) ||
      (capacity > 0 && (!out_row || !out_acc || !out_dist)))
    return set_err(RDFK_EINVAL, "null pointer argument%s%s");
  if (F > 65535) return set_err(RDFK_EINVAL, "at most 65535 frames per call%s%s");
  if ((long long)F * n_dh > 0x7fffffffLL)
    return set_err(RDFK_EINVAL, "too many (frame, donor) rows in one call; "
                                "pass fewer frames%s%s");
  const HbLayout L = hb_layout(F, n_dh, n_acc);
  if (workspace_bytes < L.total || ((uintptr_t)workspace & 255u))
    return set_err(RDFK_EWORKSPACE, "workspace too small or not 256-byte "
                                    "aligned%s%s");
  int sms = 0;
  int rc = check_device(&sms);
  if (rc) return rc;
  char *ws = (char *)workspace;
  switch (boxkind) {
    case RDFK_BOX_NONE:
      return launch_hbonds<RDFK_BOX_NONE>(
          xyz, F, natoms, donors, hydrogens, n_dh, acceptors, n_acc, box,
          max_cutoff, min_cutoff, angle_cutoff_deg, capacity, out_row, out_acc,
          out_dist, out_angle, total, ws, L, st);
    case RDFK_BOX_ORTHO:
      return launch_hbonds<RDFK_BOX_ORTHO>(
          xyz, F, natoms, donors, hydrogens, n_dh, acceptors, n_acc, box,
          max_cutoff, min_cutoff, angle_cutoff_deg, capacity, out_row, out_acc,
          out_dist, out_angle, total, ws, L, st);
    default:
      return launch_hbonds<RDFK_BOX_TRI>(
          xyz, F, natoms, donors, hydrogens, n_dh, acceptors, n_acc, box,
          max_cutoff, min_cutoff, angle_cutoff_deg, capacity, out_row, out_acc,
          out_dist, out_angle, total, ws, L, st);
  }
}
