/*
 * rdf_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic that sits under pmda's
 * InterRDF / InterRDF_s hot path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (pmda_b200/) never does.
 *
 * PARITY STATUS: *partially pinned*.
 *   - pinned   : the binning rule (oracle_np_bin) is checked against the real
 *                numpy.histogram that the reference calls (pmda/rdf.py:134,
 *                :304-305) in tests/test_oracle.py; normalisation and the block
 *                partition are checked against the reference's own code
 *                (tests/golden/make_golden.py imports /root/reference/pmda).
 *   - UNPINNED : the per-pair distance arithmetic.  It lives in MDAnalysis
 *                (>=1.0.0, setup.py:51), which is neither vendored under
 *                /root/reference nor installable offline.  The functions below
 *                restate, from memory of the published source, MDAnalysis
 *                `lib/include/calc_distances.h` (_calc_distance_array,
 *                _calc_distance_array_ortho, _calc_distance_array_triclinic,
 *                minimum_image, minimum_image_triclinic, _triclinic_pbc) and
 *                `lib/distances.py` (_bruteforce_capped).  No golden vector in
 *                /root/reference fixes a distance value independent of
 *                MDAnalysis, so these are anchored on
 *                known-answer tests we wrote ourselves (tests/test_kat.py).
 *
 * Upstream release each restated function follows (recollection of the
 * published sources; nothing of MDAnalysis is available offline to check):
 *   minimum_image, _calc_distance_array, _calc_distance_array_ortho,
 *   _calc_distance_array_triclinic, minimum_image_triclinic (27-image search,
 *   strict '<', FLT_MAX start)
 *       calc_distances.h as shipped from MDAnalysis 0.20 through 2.x (the
 *       1.0.x line pmda's setup.py:51 asks for is inside that range; the
 *       functions did not change across it).
 *   _triclinic_pbc (oracle_triclinic_wrap)
 *       the double-precision "lower bound / upper bound" variant of 1.0.x-2.x
 *       (wrap along c, then b, then a; coordinates already inside the cell
 *       untouched).  Releases up to 0.19 wrapped in float with
 *       floor(coord / box) per axis; the two differ for out-of-cell input, and
 *       the exact operation order of the newer one is NOT trusted either.
 *   _bruteforce_capped (d <= max_cutoff on the distance_array result)
 *       lib/distances.py 0.19 through 2.x.
 *   numpy.histogram's uniform-bin rule: numpy >= 1.15 (checked live).
 * tools/pin_against_mdanalysis.py is the one-command way to turn "recollection"
 * into "pinned": where MDAnalysis imports it recomputes every golden case and
 * the out-of-cell / on-edge / angle cases with the real library, diffs them
 * against this file and writes tests/golden/mdanalysis_pinned.npz, which
 * tests/test_pinning.py then enforces for the oracle and the CUDA path.  The
 * product counts frames that took the wrap branch (rdfk_rdf_stats [4], [5])
 * and the engine warns about them.
 *
 * Call sites in the reference that this file follows:
 *   pmda/rdf.py:120-137  InterRDF._single_frame   -> oracle_rdf_frame
 *   pmda/rdf.py:292-309  InterRDF_s._single_frame -> oracle_rdf_s_frame
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC
 * (strict IEEE: no FMA contraction, no reassociation).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_BOX_NONE 0
#define ORACLE_BOX_ORTHO 1
#define ORACLE_BOX_TRI 2

/* ------------------------------------------------------------------------ */
/* minimum_image  [recalled-upstream calc_distances.h]                       */
/*   s = inverse_box[i] * x[i];  x[i] = box[i] * (s - round(s))              */
/*   skipped when box[i] <= FLT_EPSILON.  inverse_box is FLOAT.              */
/* ------------------------------------------------------------------------ */
static inline void minimum_image(double *x, const float *box,
                                 const float *inverse_box) {
  for (int i = 0; i < 3; i++) {
    if (box[i] > FLT_EPSILON) {
      double s = inverse_box[i] * x[i];
      x[i] = box[i] * (s - round(s));
    }
  }
}

/* ------------------------------------------------------------------------ */
/* minimum_image_triclinic  [recalled-upstream]                              */
/*   exhaustive search over the 27 neighbouring images; strict '<' keeps the */
/*   first minimum; starts from FLT_MAX.  box is the float 3x3 lower-        */
/*   triangular matrix, row-major (a, b, c rows).                            */
/* ------------------------------------------------------------------------ */
static inline void minimum_image_triclinic(double *dx, const float *box) {
  double dx_min[3] = {0.0, 0.0, 0.0};
  double dsq_min = FLT_MAX;
  double dsq, rx, ry[2], rz[3];
  for (int ix = -1; ix < 2; ++ix) {
    rx = dx[0] + box[0] * ix;
    for (int iy = -1; iy < 2; ++iy) {
      ry[0] = rx + box[3] * iy;
      ry[1] = dx[1] + box[4] * iy;
      for (int iz = -1; iz < 2; ++iz) {
        rz[0] = ry[0] + box[6] * iz;
        rz[1] = ry[1] + box[7] * iz;
        rz[2] = dx[2] + box[8] * iz;
        dsq = rz[0] * rz[0] + rz[1] * rz[1] + rz[2] * rz[2];
        if (dsq < dsq_min) {
          dsq_min = dsq;
          dx_min[0] = rz[0];
          dx_min[1] = rz[1];
          dx_min[2] = rz[2];
        }
      }
    }
  }
  dx[0] = dx_min[0];
  dx[1] = dx_min[1];
  dx[2] = dx_min[2];
}

/* ------------------------------------------------------------------------ */
/* _triclinic_pbc  [recalled-upstream, exact op order NOT trusted]           */
/*   moves every coordinate into the primary cell: first along c (z), then   */
/*   along b (y bounded by a z-dependent lower bound), then along a.         */
/*   In-cell coordinates are left bit-identical (no branch taken), which is  */
/*   what every parity fixture relies on; out-of-cell behaviour is the       */
/*   unpinned part (SURVEY.md 8c "residual risk").                           */
/* ------------------------------------------------------------------------ */
void oracle_triclinic_wrap(float *coords, int64_t n, const float *box) {
  const double bx_by = (double)box[3] / (double)box[4];
  const double cy_cz = (double)box[7] / (double)box[8];
  const double a_zf =
      ((double)box[6] - (double)box[3] * (double)box[7] / (double)box[4]) /
      (double)box[8];
  for (int64_t i = 0; i < n; i++) {
    double c0 = coords[3 * i + 0], c1 = coords[3 * i + 1],
           c2 = coords[3 * i + 2];
    int moved = 0;
    if (c2 < 0.0 || c2 >= (double)box[8]) {
      double s = floor(c2 / (double)box[8]);
      c2 = c2 - s * (double)box[8];
      c1 = c1 - s * (double)box[7];
      c0 = c0 - s * (double)box[6];
      moved = 1;
    }
    double lb = c2 * cy_cz;
    if (c1 < lb || c1 >= lb + (double)box[4]) {
      double s = floor((c1 - lb) / (double)box[4]);
      c1 = c1 - s * (double)box[4];
      c0 = c0 - s * (double)box[3];
      moved = 1;
    }
    lb = c1 * bx_by + c2 * a_zf;
    if (c0 < lb || c0 >= lb + (double)box[0]) {
      double s = floor((c0 - lb) / (double)box[0]);
      c0 = c0 - s * (double)box[0];
      moved = 1;
    }
    if (moved) {
      coords[3 * i + 0] = (float)c0;
      coords[3 * i + 1] = (float)c1;
      coords[3 * i + 2] = (float)c2;
    }
  }
}

/* ------------------------------------------------------------------------ */
/* one pair, exactly as the loop body of _calc_distance_array{,_ortho,       */
/* _triclinic}:  FLOAT subtraction conf[j]-ref[i], widened to double,        */
/* minimum image in double, rsq left-to-right, sqrt.  For ORACLE_BOX_TRI the */
/* coordinates must already have gone through oracle_triclinic_wrap.         */
/* ------------------------------------------------------------------------ */
static inline double pair_distance(const float *ri, const float *cj,
                                   int boxkind, const float *box,
                                   const float *inverse_box) {
  double dx[3];
  float d0 = cj[0] - ri[0];
  float d1 = cj[1] - ri[1];
  float d2 = cj[2] - ri[2];
  dx[0] = d0;
  dx[1] = d1;
  dx[2] = d2;
  if (boxkind == ORACLE_BOX_ORTHO)
    minimum_image(dx, box, inverse_box);
  else if (boxkind == ORACLE_BOX_TRI)
    minimum_image_triclinic(dx, box);
  double rsq = (dx[0] * dx[0]) + (dx[1] * dx[1]) + (dx[2] * dx[2]);
  return sqrt(rsq);
}

static inline void make_inverse_box(const float *box, float *inverse_box) {
  /* inverse_box[k] = 1.0 / box[k]  (double quotient stored to float) */
  inverse_box[0] = 1.0 / box[0];
  inverse_box[1] = 1.0 / box[1];
  inverse_box[2] = 1.0 / box[2];
}

double oracle_pair_distance(const float *ri, const float *cj, int boxkind,
                            const float *box) {
  float inv[3] = {0, 0, 0};
  if (boxkind == ORACLE_BOX_ORTHO) make_inverse_box(box, inv);
  return pair_distance(ri, cj, boxkind, box, inv);
}

/* distance_array(ref, conf, box) -> double [n1][n2]; ref/conf are copied and
 * wrapped first when the box is triclinic (the reference works on the
 * contiguous float32 copies made by @check_coords).                        */
void oracle_distance_array(const float *ref, int64_t n1, const float *conf,
                           int64_t n2, int boxkind, const float *box,
                           double *out) {
  float inv[3] = {0, 0, 0};
  float *r = (float *)ref, *c = (float *)conf;
  if (boxkind == ORACLE_BOX_ORTHO) make_inverse_box(box, inv);
  if (boxkind == ORACLE_BOX_TRI) {
    r = (float *)malloc(sizeof(float) * 3 * (size_t)n1);
    c = (float *)malloc(sizeof(float) * 3 * (size_t)n2);
    memcpy(r, ref, sizeof(float) * 3 * (size_t)n1);
    memcpy(c, conf, sizeof(float) * 3 * (size_t)n2);
    oracle_triclinic_wrap(r, n1, box);
    oracle_triclinic_wrap(c, n2, box);
  }
  for (int64_t i = 0; i < n1; i++)
    for (int64_t j = 0; j < n2; j++)
      out[i * n2 + j] = pair_distance(r + 3 * i, c + 3 * j, boxkind, box, inv);
  if (boxkind == ORACLE_BOX_TRI) {
    free(r);
    free(c);
  }
}

/* ------------------------------------------------------------------------ */
/* numpy.histogram, uniform-bin fast path, one value.                        */
/* Follows numpy/lib/_histograms_impl.py:836-863 op for op:                  */
/*   keep = (a >= first_edge) & (a <= last_edge)                             */
/*   f = ((a - first_edge) / (last_edge - first_edge)) * n_equal_bins        */
/*   idx = (intp) f ; idx == n -> n-1                                        */
/*   a < edges[idx] -> idx-1 ; a >= edges[idx+1] and idx != n-1 -> idx+1     */
/* returns -1 when the value is dropped.                                     */
/* ------------------------------------------------------------------------ */
int oracle_np_bin(double a, double first_edge, double last_edge, int nbins,
                  const double *edges) {
  if (!(a >= first_edge)) return -1;
  if (!(a <= last_edge)) return -1;
  double norm_denom = last_edge - first_edge;
  double f = ((a - first_edge) / norm_denom) * (double)nbins;
  int64_t idx = (int64_t)f;
  if (idx == nbins) idx -= 1;
  if (a < edges[idx]) idx -= 1;
  if (a >= edges[idx + 1] && idx != nbins - 1) idx += 1;
  return (int)idx;
}

/* ------------------------------------------------------------------------ */
/* InterRDF._single_frame  (pmda/rdf.py:120-137)                             */
/*   capped_distance -> _bruteforce_capped: distance_array, keep d<=max_cut; */
/*   exclusion: drop pairs with i/xA == j/xB (integer floor division);       */
/*   np.histogram(dist, bins=nbins, range=(r0,r1)).                          */
/* `count` is ACCUMULATED into (the caller plays _reduce).                   */
/* g1/g2: float [n][3].  box: 3 floats (ORTHO) or 9 floats (TRI).            */
/* ------------------------------------------------------------------------ */
void oracle_rdf_frame(const float *g1, int64_t n1, const float *g2,
                      int64_t n2, int boxkind, const float *box,
                      double max_cutoff, double r0, double r1, int nbins,
                      const double *edges, int excl_a, int excl_b,
                      int64_t *count, int nthreads) {
  float inv[3] = {0, 0, 0};
  float *r = (float *)g1, *c = (float *)g2;
  if (boxkind == ORACLE_BOX_ORTHO) make_inverse_box(box, inv);
  if (boxkind == ORACLE_BOX_TRI) {
    r = (float *)malloc(sizeof(float) * 3 * (size_t)n1);
    c = (float *)malloc(sizeof(float) * 3 * (size_t)n2);
    memcpy(r, g1, sizeof(float) * 3 * (size_t)n1);
    memcpy(c, g2, sizeof(float) * 3 * (size_t)n2);
    oracle_triclinic_wrap(r, n1, box);
    oracle_triclinic_wrap(c, n2, box);
  }
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
  {
    int64_t *local = (int64_t *)calloc((size_t)nbins, sizeof(int64_t));
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n1; i++) {
      for (int64_t j = 0; j < n2; j++) {
        double d = pair_distance(r + 3 * i, c + 3 * j, boxkind, box, inv);
        if (!(d <= max_cutoff)) continue;
        if (excl_a > 0 && excl_b > 0 && (i / excl_a) == (j / excl_b)) continue;
        int k = oracle_np_bin(d, r0, r1, nbins, edges);
        if (k >= 0) local[k] += 1;
      }
    }
#pragma omp critical
    for (int k = 0; k < nbins; k++) count[k] += local[k];
    free(local);
  }
  if (boxkind == ORACLE_BOX_TRI) {
    free(r);
    free(c);
  }
}

/* ------------------------------------------------------------------------ */
/* The same frame through a cell list (orthorhombic boxes, all axes          */
/* periodic): what MDAnalysis' capped_distance does on the CPU when it       */
/* dispatches to its grid search instead of distance_array                   */
/* (lib/distances.py _determine_method: N1*N2 >= 1e8, call site              */
/* pmda/rdf.py:123).  Every pair within reach is evaluated with exactly the  */
/* arithmetic above, so the counts equal oracle_rdf_frame's; only pairs in   */
/* cells further than one cell (edge >= 1.001 * max_cutoff) apart are        */
/* skipped.  Used as the faster CPU baseline of bench.py and as a cross      */
/* check of the pruning idea.  Returns 0, or -1 when the box does not allow  */
/* a cell list (the caller then uses oracle_rdf_frame).                      */
/* ------------------------------------------------------------------------ */
int oracle_rdf_frame_cells(const float *g1, int64_t n1, const float *g2,
                           int64_t n2, const float *box, double max_cutoff,
                           double r0, double r1, int nbins,
                           const double *edges, int excl_a, int excl_b,
                           int64_t *count, int nthreads) {
  int nc[3];
  for (int k = 0; k < 3; k++) {
    if (!(box[k] > FLT_EPSILON)) return -1;
    double m = floor((double)box[k] / (max_cutoff * 1.001 + 1e-6));
    if (m < 1.0) m = 1.0;
    if (m > 512.0) m = 512.0;
    nc[k] = (int)m;
  }
  float inv[3];
  make_inverse_box(box, inv);
  const int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
  int64_t *start = (int64_t *)calloc((size_t)ncell + 1, sizeof(int64_t));
  int64_t *cell_of = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n2 > 0 ? n2 : 1));
  int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n2 > 0 ? n2 : 1));
  int bad = 0;
  for (int64_t j = 0; j < n2; j++) {
    int c[3];
    for (int k = 0; k < 3; k++) {
      double f = (double)g2[3 * j + k] / (double)box[k];
      f -= floor(f);
      if (!(f >= 0.0 && f <= 1.0)) { bad = 1; f = 0.0; }   /* NaN / Inf */
      int ci = (int)(f * nc[k]);
      c[k] = ci >= nc[k] ? nc[k] - 1 : ci;
    }
    cell_of[j] = ((int64_t)c[0] * nc[1] + c[1]) * nc[2] + c[2];
    start[cell_of[j] + 1]++;
  }
  if (bad) { free(start); free(cell_of); free(order); return -1; }
  for (int64_t c = 0; c < ncell; c++) start[c + 1] += start[c];
  {
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)ncell);
    memcpy(fill, start, sizeof(int64_t) * (size_t)ncell);
    for (int64_t j = 0; j < n2; j++) order[fill[cell_of[j]]++] = j;
    free(fill);
  }
  if (nthreads < 1) nthreads = 1;
  int fail = 0;
#pragma omp parallel num_threads(nthreads)
  {
    int64_t *local = (int64_t *)calloc((size_t)nbins, sizeof(int64_t));
#pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < n1; i++) {
      int c[3], ok = 1;
      for (int k = 0; k < 3; k++) {
        double f = (double)g1[3 * i + k] / (double)box[k];
        f -= floor(f);
        if (!(f >= 0.0 && f <= 1.0)) { ok = 0; f = 0.0; }
        int ci = (int)(f * nc[k]);
        c[k] = ci >= nc[k] ? nc[k] - 1 : ci;
      }
      if (!ok) {
#pragma omp atomic write
        fail = 1;
        continue;
      }
      /* distinct neighbour cells along each axis (fewer than 3 cells: all) */
      int nb[3][3], nn[3];
      for (int k = 0; k < 3; k++) {
        if (nc[k] < 3) {
          nn[k] = nc[k];
          for (int t = 0; t < nc[k]; t++) nb[k][t] = t;
        } else {
          nn[k] = 3;
          for (int t = 0; t < 3; t++) nb[k][t] = (c[k] + t - 1 + nc[k]) % nc[k];
        }
      }
      for (int a = 0; a < nn[0]; a++)
        for (int b = 0; b < nn[1]; b++)
          for (int e = 0; e < nn[2]; e++) {
            const int64_t cell = ((int64_t)nb[0][a] * nc[1] + nb[1][b]) * nc[2] + nb[2][e];
            for (int64_t t = start[cell]; t < start[cell + 1]; t++) {
              const int64_t j = order[t];
              double d = pair_distance(g1 + 3 * i, g2 + 3 * j, ORACLE_BOX_ORTHO,
                                       box, inv);
              if (!(d <= max_cutoff)) continue;
              if (excl_a > 0 && excl_b > 0 && (i / excl_a) == (j / excl_b)) continue;
              int k = oracle_np_bin(d, r0, r1, nbins, edges);
              if (k >= 0) local[k] += 1;
            }
          }
    }
#pragma omp critical
    if (!fail)
      for (int k = 0; k < nbins; k++) count[k] += local[k];
    free(local);
  }
  free(start);
  free(cell_of);
  free(order);
  return fail ? -1 : 0;
}

/* ------------------------------------------------------------------------ */
/* InterRDF_s._single_frame for ONE site pair (pmda/rdf.py:296-305):         */
/*   count[idx1, idx2, :] = one-hot histogram of dist for every in-range     */
/*   (idx1, idx2).  `count` is double [n1][n2][nbins]; this ADDS the one-hot */
/*   (the caller plays _reduce's `res += r`; a frame's own tensor starts at  */
/*   zero so assignment == addition within a frame).                         */
/* ------------------------------------------------------------------------ */
void oracle_rdf_s_frame(const float *g1, int64_t n1, const float *g2,
                        int64_t n2, int boxkind, const float *box,
                        double max_cutoff, double r0, double r1, int nbins,
                        const double *edges, double *count) {
  float inv[3] = {0, 0, 0};
  float *r = (float *)g1, *c = (float *)g2;
  if (boxkind == ORACLE_BOX_ORTHO) make_inverse_box(box, inv);
  if (boxkind == ORACLE_BOX_TRI) {
    r = (float *)malloc(sizeof(float) * 3 * (size_t)n1);
    c = (float *)malloc(sizeof(float) * 3 * (size_t)n2);
    memcpy(r, g1, sizeof(float) * 3 * (size_t)n1);
    memcpy(c, g2, sizeof(float) * 3 * (size_t)n2);
    oracle_triclinic_wrap(r, n1, box);
    oracle_triclinic_wrap(c, n2, box);
  }
  for (int64_t i = 0; i < n1; i++)
    for (int64_t j = 0; j < n2; j++) {
      double d = pair_distance(r + 3 * i, c + 3 * j, boxkind, box, inv);
      if (!(d <= max_cutoff)) continue;
      int k = oracle_np_bin(d, r0, r1, nbins, edges);
      if (k >= 0) count[(i * n2 + j) * nbins + k] += 1.0;
    }
  if (boxkind == ORACLE_BOX_TRI) {
    free(r);
    free(c);
  }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
