/*
 * rdfk.h -- C ABI of librdfk.so: the B200 (sm_100a) pair-distance histogram
 * kernels behind pmda.rdf.InterRDF / InterRDF_s.
 *
 * The reference (MDAnalysis/pmda) is pure Python and has no FFI for this path;
 * the boundary it crosses into native code is the pair of calls
 *
 *     pairs, dist = MDAnalysis.lib.distances.capped_distance(
 *                       g1.positions, g2.positions, maxrange, box=u.dimensions)
 *     count = np.histogram(dist, bins=nbins, range=range)[0]
 *
 * at pmda/rdf.py:123-134 (InterRDF._single_frame) and pmda/rdf.py:298-305
 * (InterRDF_s._single_frame).  Each entry point below replaces that pair of
 * calls for a whole *block of frames* (the unit pmda/parallel.py:429-439 loops
 * over), so one call == one `_dask_helper` block body.  INTEGRATION.md shows
 * the ctypes stub a pmda maintainer would add.
 *
 * Conventions
 *   - plain C types only; every pointer is a CUDA *device* pointer unless the
 *     name ends in `_host`.  `stream` is a cudaStream_t passed as void*.
 *   - the caller owns every buffer, including the scratch `workspace`
 *     (size from the matching *_workspace_bytes query, 256-byte aligned).
 *   - launches are asynchronous on `stream`; nothing synchronises.
 *   - every function returns 0 on success, a negative RDFK_E* code on error;
 *     rdfk_last_error() returns a thread-local message.  Nothing throws,
 *     nothing calls exit().
 *   - re-entrant from several host threads on different streams/devices; the
 *     library keeps no mutable global state besides per-device function
 *     attributes and the thread-local error string.
 */
#ifndef RDFK_H
#define RDFK_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDFK_VERSION 130

#define RDFK_OK 0
#define RDFK_EINVAL -1    /* bad argument                                   */
#define RDFK_ECUDA -2     /* a CUDA runtime call failed (see last_error)    */
#define RDFK_EWORKSPACE -3 /* workspace too small / misaligned              */
#define RDFK_EARCH -4     /* device is not sm_100                           */

/* box kinds: what MDAnalysis.lib.util.check_box returns for ts.dimensions   */
#define RDFK_BOX_NONE 0  /* box=None: plain distances                        */
#define RDFK_BOX_ORTHO 1 /* box[f][0..2] = lx,ly,lz (float32)                */
#define RDFK_BOX_TRI 2   /* box[f][0..8] = float32 triclinic_vectors, row-major */

int rdfk_version(void);
const char *rdfk_last_error(void);

/* Cross-check switches for the test suite; 0 (the default) in production.
 * bit 0: no bounding-box pruning (every group pair is evaluated; hydrogen
 * bonds: no acceptor cell grid), bit 1: no one-shift-per-group arithmetic,
 * bit 2: no Euclidean refinement of the triclinic box test, bit 4: the general
 * candidate-bin arithmetic of the pair kernel's drain although the range
 * starts at 0, bit 5: no
 * CUDA-graph replay of repeated small calls, bits 8-11: number
 * of slices the last work items of a launch are cut into (0 = the library's
 * own rule).  Results never
 * depend on the flags, throughput does.  An explicit call on purpose -- not an
 * environment variable a production job could inherit.  Process-wide; returns
 * the previous flags.                                                        */
int rdfk_set_debug_flags(int flags);

/* 1 when the library was built with -DRDFK_CHECK (device-side guard rails on
 * the pair kernel's per-warp queue / log bounds; see rdfk_rdf_stats), else 0 */
int rdfk_build_has_checks(void);

/* sm count, compute capability and opt-in shared memory of `device`.       */
int rdfk_device_info(int device, int *sm_count, int *cc_major, int *cc_minor,
                     size_t *smem_optin_bytes);

/* ------------------------------------------------------------------------
 * InterRDF block kernel.  Replaces, for F frames at once,
 *   capped_distance(g1.positions, g2.positions, range[1], box) + optional
 *   exclusion_block mask + np.histogram(dist, bins=nbins, range=(r0, r1))
 * (pmda/rdf.py:120-137) and the `res += r` of _reduce (pmda/rdf.py:180-190):
 * counts are ADDED into hist[0..nbins).
 *
 *   xyz        [F][natoms][3] float32  trajectory frames (ts.positions)
 *   idx1/idx2  [n1] / [n2] int32       AtomGroup.indices (NULL = 0..n-1)
 *   same_group non-zero when idx1 and idx2 are the same selection (enables
 *              the symmetric half-matrix schedule; counts are unchanged)
 *   boxkind    RDFK_BOX_*, uniform over the call
 *   box        [F][9] float32 (see RDFK_BOX_*), ignored for RDFK_BOX_NONE
 *   r0, r1     histogram range; r1 is also capped_distance's max_cutoff
 *   edges      [nbins+1] float64 = np.histogram([-1], nbins, (r0,r1))[1]
 *   excl_a/b   exclusion_block (0,0 = None): pairs with i/excl_a == j/excl_b
 *              are dropped (i, j are positions inside g1, g2)
 *   hist       [nbins] uint64, accumulated into
 * ------------------------------------------------------------------------ */
size_t rdfk_rdf_workspace_bytes(int F, int n1, int n2, int nbins,
                                int same_group);

int rdfk_rdf_hist(const float *xyz, int F, int natoms, const int *idx1, int n1,
                  const int *idx2, int n2, int same_group, int boxkind,
                  const float *box, double r0, double r1, int nbins,
                  const double *edges, int excl_a, int excl_b,
                  unsigned long long *hist, void *workspace,
                  size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------
 * InterRDF_s block kernel (pmda/rdf.py:292-309 + _reduce :362-372), batched
 * over F frames and n_sites site pairs in one launch.  For site pair s with
 * groups A_s (n1_s atoms) and B_s (n2_s atoms), cell (a, b) of the dense
 * tensor count[cell_off[s] + a*n2_s + b][0..nbins) receives +1 in the bin of
 * the distance A_s[a]..B_s[b] for every frame in which that distance is
 * within [r0, r1] (the per-pair one-hot np.histogram of :303-305).
 *
 *   idxA/idxB  concatenated AtomGroup.indices of all A_s / B_s
 *   offA/offB  [n_sites+1] CSR offsets into idxA / idxB
 *   cell_off   [n_sites+1] int64 prefix sums of n1_s*n2_s
 *   count      [cell_off[n_sites]][nbins] uint32, accumulated into
 * ------------------------------------------------------------------------ */
int rdfk_rdf_sites(const float *xyz, int F, int natoms, const int *idxA,
                   const int *idxB, const int *offA, const int *offB,
                   const long long *cell_off, int n_sites,
                   long long total_cells, int boxkind, const float *box,
                   double r0, double r1, int nbins, const double *edges,
                   unsigned int *count, void *stream);

/* ------------------------------------------------------------------------
 * Contacts frame kernel (the sibling analysis on the same
 * distance arithmetic).  Replaces, for F frames at once, the body of
 * pmda/contacts.py:270-289 (Contacts._single_frame) for ONE reference pair:
 *   d = distance_array(grA.positions, grB.positions)      (no box)
 *   r = d[initial_contacts]
 *   y = fraction_contacts(r, r0[initial_contacts], **kwargs)
 * The caller passes the contact pairs as two index lists into the frame
 * (atom idx_a[p] of grA with atom idx_b[p] of grB, row-major order of the
 * boolean mask) and their reference distances r0[p].
 *
 *   method RDFK_CONTACT_HARD    out[f*out_stride] += #{p : d_p <= r0[p]}
 *          RDFK_CONTACT_RADIUS  out[f*out_stride] += #{p : d_p <= radius}
 *          RDFK_CONTACT_SOFT    out[f*out_stride] +=
 *                                   sum_p 1 / (1 + exp(beta (d_p - lambda r0[p])))
 *          RDFK_CONTACT_DIST    out[f*out_stride + p] = d_p  (for Python
 *                               `method` callables; out_stride >= n_pairs)
 * The host divides by the number of pairs like hard_cut_q / soft_cut_q /
 * radius_cut_q (MDAnalysis.analysis.contacts) do.  Counts are exact integers
 * in double; the soft sum is accumulated in double in unspecified order.
 * ------------------------------------------------------------------------ */
#define RDFK_CONTACT_HARD 0
#define RDFK_CONTACT_RADIUS 1
#define RDFK_CONTACT_SOFT 2
#define RDFK_CONTACT_DIST 3

int rdfk_contacts(const float *xyz, int F, int natoms, const int *idx_a,
                  const int *idx_b, const double *r0, int n_pairs, int method,
                  double radius, double beta, double lambda, double *out,
                  long long out_stride, void *stream);

/* ------------------------------------------------------------------------
 * Trajectory ingest (what the reference does frame by
 * frame in pmda/parallel.py:432-434 through MDAnalysis' DCD reader).  `raw`
 * holds F frames exactly as they lie in a DCD file, `frame_stride` bytes
 * apart; inside a frame the float32 planes x[natoms], y[natoms], z[natoms]
 * start at byte offsets off_x / off_y / off_z (all multiples of 4).  Writes
 * xyz[F][natoms][3], the layout of ts.positions the kernels above take.
 * ------------------------------------------------------------------------ */
int rdfk_unpack_planes(const void *raw, long long frame_stride,
                       long long off_x, long long off_y, long long off_z,
                       int F, int natoms, float *xyz, void *stream);

/* ------------------------------------------------------------------------
 * Hydrogen-bond search.  Replaces, for F frames at
 * once, the body of pmda/hbond_analysis.py:415-470
 * (HydrogenBondAnalysis._single_frame):
 *   pairs, d = capped_distance(donors.positions, acceptors.positions,
 *                              max_cutoff, min_cutoff=1.0, box, True)
 *   ang = rad2deg(calc_angles(D[pairs[:,0]], H[pairs[:,0]], A[pairs[:,1]], box))
 *   keep = ang > angle_cutoff_deg
 * The result is ragged; it is written compacted, ordered by (frame, D-H row,
 * acceptor) -- the row-major order of the reference's brute-force search.
 *
 *   donors/hydrogens [n_dh] int32  paired atom indices (row k = donor k with
 *                                  its hydrogen k; a donor with two hydrogens
 *                                  appears twice)
 *   hydrogens == NULL              plain capped_distance(donors, acceptors):
 *                                  no angle, out_angle may be NULL
 *                                  (pmda/hbond_analysis.py:372-378 uses this
 *                                  to pair donors with hydrogens)
 *   acceptors [n_acc] int32
 *   min_cutoff < 0                 no lower bound (the reference's None)
 *   capacity                       rows the out_* arrays can hold
 *   out_row  [capacity] int32      frame * n_dh + k
 *   out_acc  [capacity] int32      position in `acceptors`
 *   out_dist [capacity] float64    D..A distance, the reference's bits
 *   out_angle[capacity] float64    D-H..A angle in degrees
 *   total    device int64          number of rows found.  When it exceeds
 *                                  `capacity` only the first `capacity` rows
 *                                  were written: call again with larger arrays.
 * ------------------------------------------------------------------------ */
size_t rdfk_hbonds_workspace_bytes(int F, int n_dh, int n_acc);

int rdfk_hbonds(const float *xyz, int F, int natoms, const int *donors,
                const int *hydrogens, int n_dh, const int *acceptors,
                int n_acc, int boxkind, const float *box, double max_cutoff,
                double min_cutoff, double angle_cutoff_deg, long long capacity,
                int *out_row, int *out_acc, double *out_dist,
                double *out_angle, long long *total, void *workspace,
                size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------
 * Frame-cache support (host memory, no CUDA).  The reference re-reads every
 * frame on every run (pmda/parallel.py:434); a frame block kept resident in
 * HBM may therefore only be reused if the host data it came from is unchanged.
 * Returns a 64-bit content hash of EVERY byte of `n_rows` rows of `row_bytes`
 * bytes lying `row_stride` bytes apart at `base_host` (a strided selection of
 * frames is hashed in place).  `nthreads` host threads share the work; the
 * result does not depend on it.  Not cryptographic.
 * ------------------------------------------------------------------------ */
unsigned long long rdfk_host_hash64_rows(const void *base_host, size_t n_rows,
                                         size_t row_bytes,
                                         long long row_stride,
                                         unsigned long long seed,
                                         int nthreads);

/* Host-side gather of the atoms an analysis references (all pointers HOST
 * memory, no CUDA): dst[f][k][0..2] = src[f][idx[k]][0..2] for f < n_frames,
 * k < n_idx, float32 triples, source frames `frame_stride` bytes apart.  This
 * is the `ag.positions` gather of pmda/rdf.py:296-301, done before the upload
 * so that only referenced atoms cross PCIe; `nthreads` host threads share the
 * frames.                                                                    */
int rdfk_host_gather_rows(const void *src_host, size_t n_frames,
                          long long frame_stride, size_t natoms,
                          const int *idx_host, size_t n_idx, float *dst_host,
                          int nthreads);

/* InterRDF_s epilogue (pmda/rdf.py:311-334) on the HOST, threaded, for the
 * counts of n_blocks frame blocks after their D2H copy, count[b][i],
 * i < n = cells * nbins:
 *   count_f64[b][i] = (double) count[b][i]        the float64 tensors the
 *                                                  reference returns per block
 *   total[i]        = sum_b count[b][i]           np.sum(self._results[:, 0])
 *   rdf[i]          = total[i] / den[i % nbins]   den = density*vol*n_frames or
 *                                                  vol*n_frames, computed by the
 *                                                  caller in the reference's
 *                                                  operation order
 * Any of the three outputs may be NULL.  One IEEE double division per
 * element, as in numpy: same bits.  All pointers are host memory.          */
int rdfk_host_sites_finish(const unsigned int *count_host, int n_blocks,
                           long long n, int nbins, const double *den_host,
                           double *count_f64_host, double *total_host,
                           double *rdf_host, int nthreads);

/* ------------------------------------------------------------------------
 * Diagnostics / measurement helpers (not on the reference path).
 * ------------------------------------------------------------------------ */

/* counters kept in `workspace` by rdfk_rdf_hist; out_host holds 8 values.
 * Of the last call: [0] pairs sent to the fp64 re-check, [1] frames that ran
 * entirely on the fp64 path, [2] (frame, 128-atom cluster) work items
 * executed, [3] pairs evaluated by the pair kernel (after the bounding-box
 * pruning).  Cumulative since the first call on this workspace: [4] atoms
 * that lay outside the primary triclinic cell and were wrapped into it (the
 * _triclinic_pbc step of distance_array, whose operation order is the one
 * part of the arithmetic restated without a pinned reference: see DESIGN.md),
 * [5] frames that held such atoms, [6] internal consistency violations (a
 * queue entry whose pair could not be found again -- always counted -- and,
 * in RDFK_CHECK builds, queue / log / staging / table-index bound violations;
 * must be 0), [7] code of
 * the first violation.  Synchronises `stream`.                               */
int rdfk_rdf_stats(const void *workspace, unsigned long long *out_host,
                   void *stream);

/* The same 8 counters copied to DEVICE memory `out_device`, asynchronously on
 * `stream` (for callers that fetch them together with the histograms).       */
int rdfk_rdf_stats_async(const void *workspace, unsigned long long *out_device,
                         void *stream);

/* cudaMemsetAsync(dst, 0, nbytes) on `stream`: zero a result vector before
 * the first accumulating call.                                               */
int rdfk_zero_async(void *dst, size_t nbytes, void *stream);

/* number of CUDA kernels this library has launched since it was loaded
 * (all streams, all devices); bench.py reports the difference over its timed
 * region as `gpu_launches`.                                                   */
unsigned long long rdfk_launch_count(void);

/* number of rdfk_rdf_hist calls answered by replaying a captured CUDA graph
 * (small calls repeated with identical arguments; see rdfk_set_debug_flags
 * bit 5).  rdfk_launch_count counts the kernels of a replay like those of a
 * plain call.                                                                 */
unsigned long long rdfk_graph_replay_count(void);

/* FP32 roofline denominator: runs a dependent-chain-free FFMA kernel on every
 * SM and returns the achieved TFLOP/s (2 flop per FFMA lane).  Synchronises. */
int rdfk_measure_fp32_peak(double *tflops_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RDFK_H */
