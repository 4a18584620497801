// ubench2.cu -- how to handle the in-range minority of pairs without paying
// warp divergence: strategies H0..H3 on realistic random data, plus shared
// atomic throughput probes.
//   H0  branch per pair, bin inline (what rdfk.cu v1 does)
//   H1  per-thread private queue, predicated append, drain when nearly full
//   H2  warp-dense queue: ballot + popc append every pair-step, dense drain
//   H3  H2 behind a warp vote per group of RI pairs
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/_build/ubench2 tools/ubench2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int TPB = 256, NW = 8, RI = 4, TJ = 1024, NB = 200;
constexpr float MAGIC = 12582912.0f;

struct P { float L, invL, cut, inv_dr; };

__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ float r2_generic(float dx, float dy, float dz, const P &p) {
  const float nx = __fadd_rn(__fmaf_rn(dx, p.invL, MAGIC), -MAGIC), ny = __fadd_rn(__fmaf_rn(dy, p.invL, MAGIC), -MAGIC), nz = __fadd_rn(__fmaf_rn(dz, p.invL, MAGIC), -MAGIC);
  dx = __fmaf_rn(-p.L, nx, dx); dy = __fmaf_rn(-p.L, ny, dy); dz = __fmaf_rn(-p.L, nz, dz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

__device__ __forceinline__ void bin_one(float r2, const P &p, const float2 *__restrict__ tab, unsigned *wh, unsigned &nslow) {
  const float d = sqrt_approx(r2);
  int k = __float2int_rd(d * p.inv_dr);
  k = max(0, min(k, NB - 1));
  const float2 t = tab[k];
  if (r2 >= t.x && r2 < t.y) atomicAdd(wh + k, 1u); else ++nslow;
}

template <int H>
__global__ __launch_bounds__(TPB, 2) void hkernel(const float *__restrict__ gi, const float *__restrict__ gj, int ntiles, P p, const float2 *__restrict__ tab_g, unsigned long long *out) {
  __shared__ __align__(16) float xs[TJ], ys[TJ], zs[TJ];
  __shared__ unsigned hist[NW][NB];
  __shared__ float2 tab[NB];
  __shared__ float q1[16 * TPB];           // H1: 16 slots per thread
  __shared__ float q2[NW][160];            // H2/H3: per-warp dense queue
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int b = tid; b < NW * NB; b += TPB) (&hist[0][0])[b] = 0;
  for (int b = tid; b < NB; b += TPB) tab[b] = tab_g[b];
  float xi[RI], yi[RI], zi[RI];
#pragma unroll
  for (int r = 0; r < RI; ++r) {
    const size_t a = (size_t)blockIdx.x * TPB * RI + r * TPB + tid;
    xi[r] = gi[3 * a]; yi[r] = gi[3 * a + 1]; zi[r] = gi[3 * a + 2];
  }
  unsigned *wh = hist[warp];
  unsigned nslow = 0, nq = 0;      // H1: my queue fill;  H2: warp queue fill (uniform)
  const unsigned lt = (1u << lane) - 1u;
  for (int t = 0; t < ntiles; ++t) {
    __syncthreads();
    for (int a = tid; a < TJ; a += TPB) {
      const float *s = gj + 3 * ((size_t)t * TJ + a);
      xs[a] = s[0]; ys[a] = s[1]; zs[a] = s[2];
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < TJ; j += 4) {
      const float4 X = *reinterpret_cast<const float4 *>(xs + j);
      const float4 Y = *reinterpret_cast<const float4 *>(ys + j);
      const float4 Z = *reinterpret_cast<const float4 *>(zs + j);
      const float xj[4] = {X.x, X.y, X.z, X.w}, yj[4] = {Y.x, Y.y, Y.z, Y.w}, zj[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float r2v[RI];
#pragma unroll
        for (int r = 0; r < RI; ++r)
          r2v[r] = r2_generic(__fsub_rn(xj[u], xi[r]), __fsub_rn(yj[u], yi[r]), __fsub_rn(zj[u], zi[r]), p);
        if (H == 0) {
          const float m = fminf(fminf(r2v[0], r2v[1]), fminf(r2v[2], r2v[3]));
          if (m < p.cut) {
#pragma unroll
            for (int r = 0; r < RI; ++r) if (r2v[r] < p.cut) bin_one(r2v[r], p, tab, wh, nslow);
          }
        } else if (H == 1) {
#pragma unroll
          for (int r = 0; r < RI; ++r) {
            if (r2v[r] < p.cut) { q1[nq * TPB + tid] = r2v[r]; ++nq; }
          }
        } else if (H == 2) {
#pragma unroll
          for (int r = 0; r < RI; ++r) {
            const bool in = r2v[r] < p.cut;
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) q2[warp][nq + __popc(m & lt)] = r2v[r];
            nq += __popc(m);
          }
        } else {
          const float mm = fminf(fminf(r2v[0], r2v[1]), fminf(r2v[2], r2v[3]));
          if (__any_sync(0xffffffffu, mm < p.cut)) {
#pragma unroll
            for (int r = 0; r < RI; ++r) {
              const bool in = r2v[r] < p.cut;
              const unsigned m = __ballot_sync(0xffffffffu, in);
              if (in) q2[warp][nq + __popc(m & lt)] = r2v[r];
              nq += __popc(m);
            }
          }
        }
        if (H == 2 || H == 3) {
          if (nq >= 32) {           // uniform: <= 128 new entries per u-step, capacity 160
            __syncwarp();
            unsigned done = 0;
            while (nq - done >= 32) { bin_one(q2[warp][done + lane], p, tab, wh, nslow); done += 32; }
            const unsigned rem = nq - done;
            float v = 0.f;
            if (lane < rem) v = q2[warp][done + lane];
            __syncwarp();
            if (lane < rem) q2[warp][lane] = v;
            nq = rem;
            __syncwarp();
          }
        }
      }
    }
    if (H == 2 || H == 3) {
      __syncwarp();
      if (lane < nq) bin_one(q2[warp][lane], p, tab, wh, nslow);
      nq = 0;
      __syncwarp();
    }
  }
  __syncthreads();
  unsigned long long tot = 0;
  for (int b = tid; b < NB; b += TPB) { unsigned long long s = 0; for (int w = 0; w < NW; ++w) s += hist[w][b]; tot += s; }
  atomicAdd(out, tot);
  atomicAdd(out + 1, (unsigned long long)nslow);
}

// H1 needs a real drain threshold; separate kernel to keep the template readable
__global__ __launch_bounds__(TPB, 2) void h1kernel(const float *__restrict__ gi, const float *__restrict__ gj, int ntiles, P p, const float2 *__restrict__ tab_g, unsigned long long *out) {
  constexpr int QS = 24;                   // slots per thread
  __shared__ __align__(16) float xs[TJ], ys[TJ], zs[TJ];
  __shared__ unsigned hist[NW][NB];
  __shared__ float2 tab[NB];
  __shared__ float q1[QS * TPB];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int b = tid; b < NW * NB; b += TPB) (&hist[0][0])[b] = 0;
  for (int b = tid; b < NB; b += TPB) tab[b] = tab_g[b];
  float xi[RI], yi[RI], zi[RI];
#pragma unroll
  for (int r = 0; r < RI; ++r) {
    const size_t a = (size_t)blockIdx.x * TPB * RI + r * TPB + tid;
    xi[r] = gi[3 * a]; yi[r] = gi[3 * a + 1]; zi[r] = gi[3 * a + 2];
  }
  unsigned *wh = hist[warp];
  unsigned nslow = 0, nq = 0;
  for (int t = 0; t < ntiles; ++t) {
    __syncthreads();
    for (int a = tid; a < TJ; a += TPB) {
      const float *s = gj + 3 * ((size_t)t * TJ + a);
      xs[a] = s[0]; ys[a] = s[1]; zs[a] = s[2];
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < TJ; j += 4) {
      const float4 X = *reinterpret_cast<const float4 *>(xs + j);
      const float4 Y = *reinterpret_cast<const float4 *>(ys + j);
      const float4 Z = *reinterpret_cast<const float4 *>(zs + j);
      const float xj[4] = {X.x, X.y, X.z, X.w}, yj[4] = {Y.x, Y.y, Y.z, Y.w}, zj[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int r = 0; r < RI; ++r) {
          const float r2 = r2_generic(__fsub_rn(xj[u], xi[r]), __fsub_rn(yj[u], yi[r]), __fsub_rn(zj[u], zi[r]), p);
          if (r2 < p.cut) { q1[nq * TPB + tid] = r2; ++nq; }
        }
      }
      if (__any_sync(0xffffffffu, nq > QS - 16)) {   // next j-step may add 16
        for (unsigned s = 0; s < nq; ++s) bin_one(q1[s * TPB + tid], p, tab, wh, nslow);
        nq = 0;
      }
    }
  }
  for (unsigned s = 0; s < nq; ++s) bin_one(q1[s * TPB + tid], p, tab, wh, nslow);
  __syncthreads();
  unsigned long long tot = 0;
  for (int b = tid; b < NB; b += TPB) { unsigned long long s = 0; for (int w = 0; w < NW; ++w) s += hist[w][b]; tot += s; }
  atomicAdd(out, tot);
  atomicAdd(out + 1, (unsigned long long)nslow);
}

// ------------------------------------------------ shared atomics throughput
template <int MODE>
__global__ __launch_bounds__(256) void atoms_kernel(unsigned *out, int iters, const int *__restrict__ bins) {
  __shared__ unsigned h[8][256];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int b = tid; b < 8 * 256; b += 256) (&h[0][0])[b] = 0;
  __syncthreads();
  int k = bins[tid];
  unsigned acc = 0;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      int b;
      if (MODE == 0) b = lane;                       // 32 distinct banks
      else if (MODE == 1) b = 7;                     // one address
      else b = (k + u * 37 + it) & 255;              // pseudo-random over 256 bins
      if (MODE == 3) { if ((lane & 7) == 0) atomicAdd(&h[warp][b], 1u); }   // 4 active lanes
      else atomicAdd(&h[warp][b], 1u);
    }
  }
  __syncthreads();
  for (int b = tid; b < 256; b += 256) acc += h[warp][b];
  if (acc == 0xffffffffu) out[0] = acc;
}

static double frand() { return rand() / (double)RAND_MAX; }

int main() {
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = sms * 2;
  const size_t ni = (size_t)grid * TPB * RI;
  const int ntiles = 24;
  const size_t nj = (size_t)ntiles * TJ;
  float *hi = (float *)malloc(ni * 12), *hj = (float *)malloc(nj * 12);
  unsigned long long *dout; CK(cudaMalloc(&dout, 64));
  float *gi, *gj; CK(cudaMalloc(&gi, ni * 12)); CK(cudaMalloc(&gj, nj * 12));
  float2 *dtab; CK(cudaMalloc(&dtab, NB * sizeof(float2)));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double Ls[2] = {100.0, 44.78};
  for (int c = 0; c < 2; ++c) {
    const double L = Ls[c], rc = 15.0, dr = rc / NB;
    srand(3 + c);
    for (size_t i = 0; i < ni * 3; ++i) hi[i] = (float)(L * frand());
    for (size_t i = 0; i < nj * 3; ++i) hj[i] = (float)(L * frand());
    CK(cudaMemcpy(gi, hi, ni * 12, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(gj, hj, nj * 12, cudaMemcpyHostToDevice));
    float2 tab[NB];
    for (int k = 0; k < NB; ++k) { const double a = k * dr + 4e-6, b = (k + 1) * dr - 4e-6; tab[k] = make_float2(k ? (float)(a * a) : -1.f, (float)(b * b)); }
    CK(cudaMemcpy(dtab, tab, sizeof(tab), cudaMemcpyHostToDevice));
    P p; p.L = (float)L; p.invL = (float)(1.0 / L); p.cut = (float)(rc * rc); p.inv_dr = (float)(1.0 / dr);
    const double pairs = (double)ni * nj;
    printf("--- L=%.2f  in-range fraction %.4f  pairs %.3e\n", L, 4.18879 * rc * rc * rc / (L * L * L), pairs);
    for (int h = 0; h < 4; ++h) {
      float best = 1e30f; unsigned long long res[2] = {0, 0};
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(dout, 0, 64));
        CK(cudaEventRecord(e0));
        if (h == 0) hkernel<0><<<grid, TPB>>>(gi, gj, ntiles, p, dtab, dout);
        if (h == 1) h1kernel<<<grid, TPB>>>(gi, gj, ntiles, p, dtab, dout);
        if (h == 2) hkernel<2><<<grid, TPB>>>(gi, gj, ntiles, p, dtab, dout);
        if (h == 3) hkernel<3><<<grid, TPB>>>(gi, gj, ntiles, p, dtab, dout);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        CK(cudaMemcpy(res, dout, 16, cudaMemcpyDeviceToHost));
      }
      CK(cudaGetLastError());
      printf("H%d  %8.3f ms  %.3e pairs/s  binned %llu  band %llu\n", h, best, pairs / (best * 1e-3), res[0], res[1]);
    }
  }
  // shared atomics
  int hb[256]; for (int i = 0; i < 256; ++i) hb[i] = rand() & 255;
  int *dbins; CK(cudaMalloc(&dbins, sizeof(hb))); CK(cudaMemcpy(dbins, hb, sizeof(hb), cudaMemcpyHostToDevice));
  const char *names[4] = {"32 distinct banks", "single address", "random 256 bins", "random, 4 active lanes"};
  for (int m = 0; m < 4; ++m) {
    float best = 1e30f; const int iters = 4096;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      if (m == 0) atoms_kernel<0><<<sms * 8, 256>>>((unsigned *)dout, iters, dbins);
      if (m == 1) atoms_kernel<1><<<sms * 8, 256>>>((unsigned *)dout, iters, dbins);
      if (m == 2) atoms_kernel<2><<<sms * 8, 256>>>((unsigned *)dout, iters, dbins);
      if (m == 3) atoms_kernel<3><<<sms * 8, 256>>>((unsigned *)dout, iters, dbins);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double winstr = (double)iters * 8 * 8 * sms * 8;
    printf("ATOMS %-24s %8.3f ms  %.3f warp-instr/clk/SM @1.965GHz\n", names[m], best, winstr / sms / (best * 1e-3) / 1.965e9);
  }
  return 0;
}
